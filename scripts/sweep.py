#!/usr/bin/env python
"""BASELINE.json configs 1, 3 and 4 on one GPU: for every chroma subsampling x quality (x Huffman mode) of the
8320x40000 synthetic image: JPEG size, PSNR of the reconstruction, device-resident encode and decode Mpix/s; plus
the secondary-compression round trip (encode -> decode -> difference map -> encode(diff) + PSNR) at the headline
setting. Prints one JSON line per case (and writes them to --out).

    python scripts/sweep.py [--quick] [--out gpurun_out/sweep.jsonl]
"""
import argparse, ctypes, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import nvjpeg_imagecompressor_b200 as P
from nvjpeg_imagecompressor_b200.synth import synth

ap = argparse.ArgumentParser()
ap.add_argument("--width", type=int, default=8320); ap.add_argument("--height", type=int, default=40000)
ap.add_argument("--quick", action="store_true"); ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--out", default="")
a = ap.parse_args()
W, H = a.width, a.height
img = synth(W, H)
torch.cuda.synchronize()
cudart = ctypes.CDLL("libcudart.so")
out_f = open(a.out, "w") if a.out else None
rec = torch.empty((H, W, 3), dtype=torch.uint8, device="cuda")


def emit(d):
    s = json.dumps(d)
    print(s, flush=True)
    if out_f:
        out_f.write(s + "\n"); out_f.flush()


cases = [("420", 95, 0)]                                                     # config 1
qs = (95,) if a.quick else (75, 85, 95, 100)
cases += [(c, q, 1) for c in ("444", "422", "440", "420", "411") for q in qs]   # config 3
st = torch.cuda.Stream()
for css, q, opt in cases:
    eng = P.Engine(W, H, q, bool(opt), css)
    eng.set_stream(st.cuda_stream)
    eng.enable_timing(True)
    enc_ms = []
    with torch.cuda.stream(st):
        for i in range(a.iters + 1):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            optr, _ = eng.encode_device(img.data_ptr(), W * 3, W, H)
            e1.record(st)
            n = eng.encode_finish()
            if i:
                enc_ms.append(e0.elapsed_time(e1))
        tm = eng.timings()
        jpg = torch.empty(n, dtype=torch.uint8, pin_memory=True)
        cudart.cudaMemcpy(ctypes.c_void_p(jpg.data_ptr()), ctypes.c_void_p(optr), ctypes.c_size_t(n), 2)
        dec_ms = []
        for i in range(a.iters + 1):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            eng.decode_device(jpg.numpy(), rec.data_ptr(), W * 3)
            e1.record(st)
            eng.decode_finish()
            if i:
                dec_ms.append(e0.elapsed_time(e1))
        dt = eng.timings()
        sptr = eng.diff_psnr_device(img.data_ptr(), rec.data_ptr(), img.numel(), 0, 0)
        st.synchronize()
        ssd_h = ctypes.c_uint64(0)
        cudart.cudaMemcpy(ctypes.byref(ssd_h), ctypes.c_void_p(sptr), ctypes.c_size_t(8), 2)
        ssd = int(ssd_h.value)
    psnr = float("inf") if ssd == 0 else 20 * np.log10(255.0 / np.sqrt(ssd / img.numel()))
    e, d = float(np.median(enc_ms)), float(np.median(dec_ms))
    emit({"case": "encode/decode", "css": css, "quality": q, "optimize": opt, "jpeg_bytes": int(n),
          "ratio_pct": round(100.0 * n / img.numel(), 2), "psnr_db": round(psnr, 3),
          "encode_ms": round(e, 3), "encode_mpix_s": round(W * H / e / 1e3, 1),
          "decode_ms": round(d, 3), "decode_mpix_s": round(W * H / d / 1e3, 1),
          "enc_stages_ms": {k: round(v, 3) for k, v in tm.items() if k in ("fdct", "tables", "pack", "scan", "stuff")},
          "dec_stages_ms": {k: round(v, 3) for k, v in dt.items() if k.startswith("dec_")}})
    eng.close()

# config 4: secondary compression round trip through the host API (host BGR in, two JPEGs + PSNR out)
eng = P.Engine(W, H, 95, True, "422")
h_img = img.cpu().numpy()
for i in range(2):
    t0 = time.perf_counter()
    j1, j2, _, ps = eng.secondary(h_img, diff_mode=1, want_recon=False)
    dt = (time.perf_counter() - t0) * 1e3
emit({"case": "secondary (host BGR -> 2 host JPEGs + PSNR, pageable buffers)", "css": "422", "quality": 95, "optimize": 1,
      "wall_ms": round(dt, 1), "mpix_s": round(W * H / dt / 1e3, 1), "jpeg1_bytes": int(j1.size), "jpeg2_bytes": int(j2.size),
      "psnr_db": round(float(ps), 3)})
eng.close()
