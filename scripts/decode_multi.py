#!/usr/bin/env python
"""ONE 8320x40000 image (restart intervals of one MCU row) decoded by N GPUs, one process each
(strips.StripDecoder): every rank takes a byte range of the scan, finds its RSTn markers on the GPU, one all_gather of
the marker counts tells it which MCU rows it holds, it decodes them (plus a halo interval) from device memory.

    torchrun --nproc-per-node N scripts/decode_multi.py [--css 422 --quality 95 --rows 1]
The timed region starts with the JPEG in host memory (every rank has the file) and ends with the rank's rows in HBM.
"""
import argparse, hashlib, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import torch.distributed as dist
import nvjpeg_imagecompressor_b200 as P
from nvjpeg_imagecompressor_b200.strips import StripDecoder
from nvjpeg_imagecompressor_b200.synth import synth

ap = argparse.ArgumentParser()
ap.add_argument("--width", type=int, default=8320); ap.add_argument("--height", type=int, default=40000)
ap.add_argument("--css", default="422"); ap.add_argument("--quality", type=int, default=95)
ap.add_argument("--rows", type=int, default=1); ap.add_argument("--iters", type=int, default=5)
a = ap.parse_args()
rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
W, H = a.width, a.height
# every rank writes the same stream itself (deterministic synthetic image, b2j_set_restart_rows)
eng = P.Engine(W, H, a.quality, True, a.css, device=lr)
eng.set_restart_rows(a.rows)
img = synth(W, H)
optr, _ = eng.encode_device(img.data_ptr(), W * 3, W, H)
n = eng.encode_finish()
from nvjpeg_imagecompressor_b200.strips import _view
jt = torch.empty(n, dtype=torch.uint8, pin_memory=True)
jt.copy_(_view(optr, (n,), "|u1", dev)); torch.cuda.synchronize()
jpg = jt.numpy()
want_sha = None
if rank == 0:
    full = torch.empty((H, W, 3), dtype=torch.uint8, device=dev)
    eng.decode_device(jpg, full.data_ptr(), W * 3); eng.decode_finish()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    eng.decode_device(jpg, full.data_ptr(), W * 3); eng.decode_finish()
    torch.cuda.synchronize()
    one_gpu_ms = (time.perf_counter() - t0) * 1e3
eng.close(); del img
dec = StripDecoder(jt, rank, world, device=lr)
cnt = torch.zeros(world, dtype=torch.int64, device=dev)
def run():
    c = dec.count_markers()
    if world > 1:
        mine = torch.tensor([c], dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(cnt, mine)
        counts = cnt.cpu().tolist()
    else:
        counts = [c]
    return dec.decode(counts)
y0, rows = run()
torch.cuda.synchronize()
ts = []
for _ in range(a.iters):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    y0, rows = run()
    torch.cuda.synchronize()
    ts.append((time.perf_counter() - t0) * 1e3)
t = torch.tensor([min(ts)], device=dev, dtype=torch.float64)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
# correctness: rank 0 compares every rank's rows with its own single-GPU decode
ok = True
if world > 1:
    shapes = [None] * world
    dist.all_gather_object(shapes, (int(y0), int(rows.shape[0])))
    if rank == 0:
        ok = bool(torch.equal(rows, full[y0:y0 + rows.shape[0]]))
        for r in range(1, world):
            buf = torch.empty((shapes[r][1], W, 3), dtype=torch.uint8, device=dev)
            dist.recv(buf, src=r)
            ok = ok and bool(torch.equal(buf, full[shapes[r][0]:shapes[r][0] + shapes[r][1]]))
    else:
        dist.send(rows.contiguous(), dst=0)
else:
    ok = bool(torch.equal(rows, full))
if rank == 0:
    print(json.dumps({"case": "single-image decode over strips (restart intervals)", "W": W, "H": H, "css": a.css, "quality": a.quality,
                      "restart_rows": a.rows, "n_gpus": world, "jpeg_bytes": int(n), "ms": round(float(t.item()), 3),
                      "mpix_s": round(W * H / float(t.item()) / 1e3, 1), "one_gpu_whole_image_ms": round(one_gpu_ms, 3),
                      "equals_single_gpu_decode": ok}))
if world > 1:
    dist.barrier(); dist.destroy_process_group()
