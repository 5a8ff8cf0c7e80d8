#!/usr/bin/env python
"""Device times of the decode path for one configuration."""
import argparse, json, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import nvjpeg_imagecompressor_b200 as P
from nvjpeg_imagecompressor_b200.synth import synth
ap = argparse.ArgumentParser()
ap.add_argument("--width", type=int, default=8320); ap.add_argument("--height", type=int, default=40000)
ap.add_argument("--css", default="422"); ap.add_argument("--quality", type=int, default=95)
ap.add_argument("--optimize", type=int, default=1); ap.add_argument("--iters", type=int, default=4)
ap.add_argument("--restart-rows", type=int, default=0)
a = ap.parse_args()
img = synth(a.width, a.height)
eng = P.Engine(a.width, a.height, a.quality, bool(a.optimize), a.css)
st = torch.cuda.Stream(); eng.set_stream(st.cuda_stream)
if a.restart_rows:
    eng.set_restart_rows(a.restart_rows)
optr, lptr = eng.encode_device(img.data_ptr(), a.width * 3, a.width, a.height)
n = eng.encode_finish()
jpg_d = torch.empty(n, dtype=torch.uint8, device="cuda")
import ctypes
cudart = ctypes.CDLL("libcudart.so")
cudart.cudaMemcpy(ctypes.c_void_p(jpg_d.data_ptr()), ctypes.c_void_p(optr), ctypes.c_size_t(n), 3)
jpg = torch.empty(n, dtype=torch.uint8, pin_memory=True); jpg.copy_(jpg_d); torch.cuda.synchronize()
out = torch.empty((a.height, a.width, 3), dtype=torch.uint8, device="cuda")
eng.enable_timing(True)
for i in range(a.iters):
    t0 = time.perf_counter()
    eng.decode_device(jpg.numpy(), out.data_ptr(), a.width * 3)
    eng.decode_finish()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) * 1e3
    t = eng.timings()
    print(json.dumps(dict(iter=i, jpeg_bytes=n, wall_ms=round(dt, 2), mpix_s=round(a.width * a.height / dt / 1e3, 1),
                          **{k: round(v, 3) for k, v in t.items() if k.startswith("dec_") or k == "total"})))
print("roundtrip equal to input? psnr:", end=" ")
s = eng.diff_psnr_device(img.data_ptr(), out.data_ptr(), img.numel(), 0, 0)
torch.cuda.synchronize()
