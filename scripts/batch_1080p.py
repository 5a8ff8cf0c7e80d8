#!/usr/bin/env python
"""BASELINE.json config 5 (second half): a batch of 1920x1080 images, sharded trivially over the ranks (one process
per GPU; images are independent, no collective on the data path). Every rank encodes and decodes its share with a few
engines on separate streams so that launch latency of one image hides behind the kernels of another.

    python scripts/batch_1080p.py [--images 512] [--engines 8]            # one GPU
    torchrun --nproc-per-node 8 scripts/batch_1080p.py --images 4096      # 512 per GPU
"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import nvjpeg_imagecompressor_b200 as P
from nvjpeg_imagecompressor_b200.synth import synth_rows

ap = argparse.ArgumentParser()
ap.add_argument("--images", type=int, default=512); ap.add_argument("--engines", type=int, default=8)
ap.add_argument("--threads", type=int, default=0, help="host threads per GPU, one engine each (0: one thread drives all engines)")
ap.add_argument("--css", default="420"); ap.add_argument("--quality", type=int, default=95); ap.add_argument("--optimize", type=int, default=1)
a = ap.parse_args()
rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
if world > 1:
    import torch.distributed as dist
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
W, H = 1920, 1080
mine = list(range(rank, a.images, world))            # image ids = synth seeds
nuniq = min(len(mine), 16)                           # distinct images kept resident; the batch cycles through them
imgs = [synth_rows(W, H, 0, H, seed, 8, dev) for seed in mine[:nuniq]]
if a.threads:
    a.engines = a.threads
engs = [P.Engine(W, H, a.quality, bool(a.optimize), a.css, device=lr) for _ in range(a.engines)]
streams = [torch.cuda.Stream(device=dev) for _ in engs]
for e, s in zip(engs, streams):
    e.set_stream(s.cuda_stream)
torch.cuda.synchronize()


def per_thread(fn):
    """One host thread per engine (ctypes releases the GIL inside the library): fn(engine index, image indices)."""
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(len(engs)) as ex:
        return sum(ex.map(lambda k: fn(k, range(k, len(mine), len(engs))), range(len(engs))))


def enc_worker(k, idx):
    torch.cuda.set_device(lr)
    n = 0
    for i in idx:
        engs[k].encode_device(imgs[i % nuniq].data_ptr(), W * 3, W, H)
        n += engs[k].encode_finish()
    return n


def dec_worker(k, idx):
    torch.cuda.set_device(lr)
    for i in idx:
        engs[k].decode_device(jpgs[i % nuniq], outs[k].data_ptr(), W * 3)
        engs[k].decode_finish()
    return 0


def run_encode():
    if a.threads:
        return per_thread(enc_worker)
    n = 0
    for i in range(len(mine)):
        e = engs[i % len(engs)]
        if i >= len(engs):
            n += e.encode_finish()                    # the engine's previous image
        e.encode_device(imgs[i % nuniq].data_ptr(), W * 3, W, H)
    for e in engs[: min(len(engs), len(mine))]:
        n += e.encode_finish()
    return n


run_encode()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
nbytes = run_encode()
torch.cuda.synchronize()
dt = time.perf_counter() - t0
# decode: one JPEG per distinct image (host bytes), outputs stay on the device
jpgs = []
for k in range(nuniq):
    j = engs[0].encode(imgs[k].cpu().numpy())
    keep = torch.empty(len(j), dtype=torch.uint8, pin_memory=True)   # pinned JPEG bytes: the upload is a plain DMA
    keep.numpy()[:] = j
    jpgs.append(keep.numpy())
outs = [torch.empty((H, W, 3), dtype=torch.uint8, device=dev) for _ in engs]
for k, e in enumerate(engs):
    e.decode_device(jpgs[k % nuniq], outs[k].data_ptr(), W * 3)
for e in engs:
    e.decode_finish()
t0 = time.perf_counter()
if a.threads:
    per_thread(dec_worker)
else:
    for i in range(len(mine)):
        e = engs[i % len(engs)]
        if i >= len(engs):
            e.decode_finish()                              # the engine's previous image: validated, pixels usable
        e.decode_device(jpgs[i % nuniq], outs[i % len(engs)].data_ptr(), W * 3)
    for e in engs:
        e.decode_finish()
dtd = time.perf_counter() - t0
res = torch.tensor([dt, dtd], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(res, op=dist.ReduceOp.MAX)
if rank == 0:
    mp = a.images * W * H / 1e6
    print(json.dumps({"case": "batch 1920x1080", "images": a.images, "n_gpus": world, "engines_per_gpu": a.engines, "host_threads_per_gpu": a.threads or 1, "css": a.css,
                      "quality": a.quality, "optimize": a.optimize, "encode_mpix_s": round(mp / float(res[0]), 1),
                      "encode_images_s": round(a.images / float(res[0]), 1), "decode_mpix_s": round(mp / float(res[1]), 1),
                      "decode_images_s": round(a.images / float(res[1]), 1), "jpeg_bytes_per_image": int(nbytes / len(mine))}))
if world > 1:
    dist.barrier(); dist.destroy_process_group()
