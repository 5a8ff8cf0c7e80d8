#!/usr/bin/env python
"""BASELINE.json config 4: secondary compression round trip (encode -> reconstruct -> difference map -> encode(diff) +
PSNR) of the 8320x40000 image, on 1 GPU (b2j_secondary) or as MCU-row strips on N GPUs (strips.StripSecondary).

    python scripts/secondary_multi.py                                           # 1 GPU, host -> host through the C-ABI
    torchrun --nproc-per-node 8 scripts/secondary_multi.py [--check 1]          # 8 strips, device resident
--check 1: rank 0 also runs the single-GPU path on the whole image and compares both stitched streams and the PSNR.
"""
import argparse, hashlib, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import nvjpeg_imagecompressor_b200 as P
from nvjpeg_imagecompressor_b200.synth import synth, synth_rows

ap = argparse.ArgumentParser()
ap.add_argument("--width", type=int, default=8320); ap.add_argument("--height", type=int, default=40000)
ap.add_argument("--css", default="422"); ap.add_argument("--quality", type=int, default=95)
ap.add_argument("--iters", type=int, default=5); ap.add_argument("--check", type=int, default=1)
a = ap.parse_args()
rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
W, H = a.width, a.height
sha = lambda x: hashlib.sha256(np.ascontiguousarray(x).tobytes()).hexdigest()[:32]
out = {"case": "secondary compression", "W": W, "H": H, "css": a.css, "quality": a.quality, "n_gpus": world}
if world == 1:
    img = synth(W, H).cpu().numpy()
    eng = P.Engine(W, H, a.quality, True, a.css)
    j1, j2, _, ps = eng.secondary(img, 1, want_recon=False)
    ts = []
    for _ in range(a.iters):
        t0 = time.perf_counter(); j1, j2, _, ps = eng.secondary(img, 1, want_recon=False); ts.append(time.perf_counter() - t0)
    out.update(ms_host_to_host=round(min(ts) * 1e3, 2), jpeg1_bytes=int(j1.size), jpeg2_bytes=int(j2.size), psnr=round(ps, 4),
               jpeg1_sha=sha(j1), jpeg2_sha=sha(j2))
    # device resident (b2j_secondary_device): the image already in HBM, both streams and the difference map stay there
    d_img = torch.from_numpy(img).to(dev)
    st = torch.cuda.Stream(); eng.set_stream(st.cuda_stream)
    with torch.cuda.stream(st):
        eng.secondary_device(d_img.data_ptr(), W * 3, W, H, 1); eng.secondary_finish()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(a.iters):
            eng.secondary_device(d_img.data_ptr(), W * 3, W, H, 1)
        e1.record(st)
        n1, n2, ps2, ssd = eng.secondary_finish()
    out.update(ms_device_resident=round(e0.elapsed_time(e1) / a.iters, 3), device_lengths=[int(n1), int(n2)], device_psnr=round(ps2, 4))
    print(json.dumps(out))
    sys.exit(0)
import torch.distributed as dist
from nvjpeg_imagecompressor_b200.strips import StripSecondary
dist.init_process_group("nccl", device_id=dev)
sec = StripSecondary(W, H, a.quality, True, a.css, diff_mode=1, device=lr)
strip = synth_rows(W, H, sec.y0, sec.y1 - sec.y0, 0, 8, dev)
for _ in range(2):
    sec.run(strip.data_ptr(), W * 3)
torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.iters):
    n1, n2, ps = sec.run(strip.data_ptr(), W * 3)
torch.cuda.current_stream().wait_stream(sec.stream)
e1.record(); torch.cuda.synchronize()
t = torch.tensor([e0.elapsed_time(e1) / a.iters], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
with torch.cuda.stream(sec.stream):
    s1 = sec.enc1.gather_jpeg(0)
    s2 = sec.enc2.gather_jpeg(0)
    sec.stream.synchronize()
if rank == 0:
    s1, s2 = s1.cpu().numpy(), s2.cpu().numpy()
    out.update(ms_device_resident=round(float(t.item()), 3), mpix_s=round(W * H / float(t.item()) / 1e3, 1), jpeg1_bytes=int(s1.size),
               jpeg2_bytes=int(s2.size), psnr=round(ps, 4), jpeg1_sha=sha(s1), jpeg2_sha=sha(s2), exchange=sec.enc1.exchange)
    if a.check:
        img = synth(W, H).cpu().numpy()
        eng = P.Engine(W, H, a.quality, True, a.css, device=lr)
        j1, j2, _, p1 = eng.secondary(img, 1, want_recon=False)
        out["equals_single_gpu"] = bool(np.array_equal(j1, s1) and np.array_equal(j2, s2) and abs(p1 - ps) < 1e-9)
    print(json.dumps(out))
dist.barrier(); dist.destroy_process_group()
