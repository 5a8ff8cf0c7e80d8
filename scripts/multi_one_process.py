#!/usr/bin/env python
"""b2j_multi_*: the headline image over N GPUs from ONE process and one host thread (host BGR -> host JPEG).
    python scripts/multi_one_process.py --gpus 8
"""
import argparse, hashlib, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import nvjpeg_imagecompressor_b200 as P
import oracle as O

ap = argparse.ArgumentParser()
ap.add_argument("--gpus", type=int, default=torch.cuda.device_count())
ap.add_argument("--width", type=int, default=8320); ap.add_argument("--height", type=int, default=40000)
ap.add_argument("--iters", type=int, default=5)
a = ap.parse_args()
W, H = a.width, a.height
h_img = torch.empty((H, W, 3), dtype=torch.uint8, pin_memory=True)
for y0 in range(0, H, 2000):
    r = min(2000, H - y0)
    h_img[y0:y0 + r] = torch.from_numpy(O.synth(W, H, 0, 8, y0=y0, rows=r))
h_out = torch.empty(W * H, dtype=torch.uint8, pin_memory=True)
m = P.MultiEngine(W, H, 95, True, "422", devices=list(range(a.gpus)))
n = m.encode_ptr(h_img.data_ptr(), W * 3, W, H, h_out.data_ptr(), h_out.numel())
ts = []
for _ in range(a.iters):
    t0 = time.perf_counter()
    n = m.encode_ptr(h_img.data_ptr(), W * 3, W, H, h_out.data_ptr(), h_out.numel())
    ts.append((time.perf_counter() - t0) * 1e3)
digest = hashlib.sha256(h_out[:n].numpy().tobytes()).hexdigest()
p_img = np.array(h_img.numpy(), copy=True)
p_out = np.zeros(W * H, np.uint8)
m.encode(p_img, out=p_out)
tp = []
for _ in range(3):
    t0 = time.perf_counter(); m.encode(p_img, out=p_out); tp.append((time.perf_counter() - t0) * 1e3)
exact = None
try:
    with open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "golden.json")) as f:
        for c in json.load(f)["headline"]["encodes"]:
            if (c["css"], c["quality"], c["optimize"]) == (1, 95, 1) and (W, H) == (8320, 40000):
                exact = bool(c["jpeg_len"] == n and digest[:32] == c["jpeg_sha256_128"])
except Exception:
    pass
print(json.dumps({"case": "b2j_multi_encode, one process", "n_gpus": a.gpus, "W": W, "H": H, "jpeg_bytes": int(n),
                  "pinned_ms": round(min(ts), 2), "pageable_ms": round(min(tp), 2), "mpix_s_pinned": round(W * H / min(ts) / 1e3, 1),
                  "bit_exact_vs_libjpeg_turbo": exact}))
m.close()
