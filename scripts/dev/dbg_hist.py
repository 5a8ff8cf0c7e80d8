import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import nvjpeg_imagecompressor_b200 as P
from nvjpeg_imagecompressor_b200 import _native as N
import oracle as O
for (W, H, css) in ((64, 96, 0), (48, 64, 0), (50, 70, 0), (135, 121, 1)):
    q, opt = 95, 1
    eng = P.Engine(300, 160, q, True, css)
    img = O.synth(W, H, W * 31 + H, 8)
    eng.set_debug(1)
    jpg = eng.encode(img)
    g = O.geometry(W, H, css)
    ref = O.forward(img, css, q)
    hist = eng.debug_read(N.DBG_HIST, np.uint32).reshape(4, 257).astype(np.int64)
    want = O.histogram(ref, g.bpm).astype(np.int64)
    d = np.argwhere(hist != want)
    print(W, H, css, "ndiff", len(d), [(int(t), hex(int(s)), int(hist[t, s]), int(want[t, s])) for t, s in d[:12]])
    w2 = O.encode(img, css, q, opt)
    print("  bytes equal:", jpg.size == w2.size and bool(np.array_equal(jpg, w2)))
    eng.close()
