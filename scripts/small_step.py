#!/usr/bin/env python
"""Device time per encode of a 1/8-height strip-sized image (what one rank of an 8-GPU run does), with and without
the library's per-stage CUDA events, to see what the fixed costs are."""
import argparse, json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import nvjpeg_imagecompressor_b200 as P
from nvjpeg_imagecompressor_b200.synth import synth
ap = argparse.ArgumentParser()
ap.add_argument("--width", type=int, default=8320); ap.add_argument("--height", type=int, default=5000)
ap.add_argument("--iters", type=int, default=50)
a = ap.parse_args()
img = synth(a.width, a.height)
eng = P.Engine(a.width, a.height, 95, True, "422")
st = torch.cuda.Stream(); eng.set_stream(st.cuda_stream)
for timing in (True, False, True, False):
    eng.enable_timing(timing)
    with torch.cuda.stream(st):
        for _ in range(5):
            eng.encode_device(img.data_ptr(), a.width * 3, a.width, a.height)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(a.iters):
            eng.encode_device(img.data_ptr(), a.width * 3, a.width, a.height)
        e1.record(st)
        st.synchronize()
    n = eng.encode_finish()
    out = dict(timing=timing, ms_per_encode=round(e0.elapsed_time(e1) / a.iters, 4), jpeg_bytes=n)
    if timing:
        out.update({k: round(v, 4) for k, v in eng.timings().items() if not k.startswith("dec_")})
    print(json.dumps(out))
