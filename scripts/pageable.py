#!/usr/bin/env python
"""b2j_encode / b2j_decode on PAGEABLE host buffers (what cv::Mat and std::vector are) for a few copy-thread counts
(B2J_COPY_THREADS is read when the context first needs its copy pool)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
W, H = 8320, 40000
for nt in [int(x) for x in (sys.argv[1:] or ["4", "8", "12", "16"])]:
    os.environ["B2J_COPY_THREADS"] = str(nt)
    import nvjpeg_imagecompressor_b200 as P
    from nvjpeg_imagecompressor_b200.synth import synth
    img = synth(W, H).cpu().numpy()
    eng = P.Engine(W, H, 95, True, "422")
    out = np.zeros(200 << 20, np.uint8)
    jpg = eng.encode(img, out=out)
    ts = []
    for _ in range(5):
        t0 = time.perf_counter(); jpg = eng.encode(img, out=out); ts.append(time.perf_counter() - t0)
    rec = np.zeros((H, W, 3), np.uint8)
    j = np.array(jpg, copy=True)
    eng.decode_ptr(j, rec.ctypes.data, W * 3)
    td = []
    for _ in range(3):
        t0 = time.perf_counter(); eng.decode_ptr(j, rec.ctypes.data, W * 3); td.append(time.perf_counter() - t0)
    print(json.dumps(dict(copy_threads=nt, encode_ms=[round(t * 1e3, 2) for t in ts], decode_ms=[round(t * 1e3, 2) for t in td])))
    eng.close()
