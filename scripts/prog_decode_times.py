#!/usr/bin/env python
"""Wall-clock of b2j_decode on progressive (SOF2) files written by the checker's jpeg_simple_progression encoder
(byte-identical to cv2.imencode(IMWRITE_JPEG_PROGRESSIVE)): sequential per scan without restart markers."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import nvjpeg_imagecompressor_b200 as P
import oracle as O
for (W, H, css) in ((640, 360, 3), (1920, 1080, 3), (1920, 1080, 0), (4096, 3072, 1)):
    img = O.synth(W, H, 2, 8)
    jpg = O.encode_progressive(img, css, 95)
    eng = P.Engine(W, H, 95, True, css)
    out = eng.decode(jpg)
    ts = []
    for _ in range(3):
        t0 = time.perf_counter(); out = eng.decode(jpg); ts.append((time.perf_counter() - t0) * 1e3)
    base = O.encode(img, css, 95, 1)
    t0 = time.perf_counter(); eng.decode(base); tb = (time.perf_counter() - t0) * 1e3
    print(json.dumps({"case": "progressive decode (host -> host)", "W": W, "H": H, "css": css, "jpeg_bytes": int(jpg.size), "ms": round(min(ts), 2),
                      "baseline_file_same_image_ms": round(tb, 2)}))
    eng.close()
