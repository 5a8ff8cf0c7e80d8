#!/usr/bin/env python
"""BASELINE.json config 5, first half at full size: the 8320x40000 JPEG written by the reference's nvJPEG call sequence
(baseline/_ref/libref_nvjpeg.so: baseline sequential, optimised Huffman, no restart markers) decoded by this engine:
time (host bytes -> BGR in HBM, CUDA events) and pixel equality with libjpeg-turbo (cv2.imdecode) on the same bytes."""
import ctypes as C, hashlib, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import nvjpeg_imagecompressor_b200 as P
from nvjpeg_imagecompressor_b200.synth import synth
W, H = 8320, 40000
L = C.CDLL(os.path.join(ROOT, "baseline", "_ref", "libref_nvjpeg.so"))
img = synth(W, H).cpu().numpy()
for css, name in ((1, "422"), (0, "444")):
    h = C.c_void_p()
    assert L.ref_create(W, H, 95, 1, css, 0, C.byref(h)) == 0
    assert L.ref_build_compress_env(h) == 0, "nvJPEG cannot run here"
    out = np.empty(W * H * 3 // 2, np.uint8)
    n = C.c_size_t(0)
    assert L.ref_compress(h, C.c_void_p(img.ctypes.data), C.c_size_t(W * 3), C.c_void_p(out.ctypes.data), C.c_size_t(out.size), C.byref(n)) == 0
    L.ref_destroy(h)
    pin = torch.empty(n.value, dtype=torch.uint8, pin_memory=True)
    pin.numpy()[:] = out[: n.value]
    jpg = pin.numpy()
    eng = P.Engine(W, H, 95, True, name)
    d = torch.empty((H, W, 3), dtype=torch.uint8, device="cuda")
    st = torch.cuda.Stream(); eng.set_stream(st.cuda_stream)
    with torch.cuda.stream(st):
        eng.decode_device(jpg, d.data_ptr(), W * 3); eng.decode_finish()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(5):
            eng.decode_device(jpg, d.data_ptr(), W * 3); eng.decode_finish()
        e1.record(st); st.synchronize()
    ms = e0.elapsed_time(e1) / 5
    got = d.cpu().numpy()
    import cv2
    t0 = time.perf_counter(); want = cv2.imdecode(jpg, cv2.IMREAD_COLOR); tcv = time.perf_counter() - t0
    print(json.dumps(dict(case="decode of the nvJPEG-written stream", W=W, H=H, css=name, jpeg_bytes=int(n.value), decode_ms=round(ms, 3),
                          mpix_s=round(W * H / ms / 1e3, 1), equals_libjpeg_turbo=bool(np.array_equal(got, want)),
                          cv2_imdecode_ms=round(tcv * 1e3, 1))))
    eng.close(); del d, got, want
