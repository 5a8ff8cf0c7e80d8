import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import nvjpeg_imagecompressor_b200 as P
from nvjpeg_imagecompressor_b200 import _native as N
import oracle as O
W, H, css, q = 48, 64, 0, 95
img = O.synth(W, H, W * 31 + H, 8)
for opt in (1, 0):
    for dbg in (0, 1):
        eng = P.Engine(300, 160, q, bool(opt), css)
        eng.set_debug(dbg)
        jpg = eng.encode(img)
        pool = eng.debug_read(N.DBG_TOKENS, np.uint32)
        recs = eng.debug_read(N.DBG_TILE_RECS, np.uint8).reshape(-1, 24)
        base, count = recs[3, :8].view(np.uint32)
        w2 = O.encode(img, css, q, opt)
        print("opt", opt, "dbg", dbg, "tok239", hex(int(pool[base + 239])), "bytes equal", jpg.size == w2.size and bool(np.array_equal(jpg, w2)),
              "recs3", recs[3, 8:].view(np.int16).tolist())
        eng.close()
# the DC values around: block index
ref = O.forward(img, css, q)
g = O.geometry(W, H, css)
for mx in range(6):
    b = (3 * 6 + mx) * 3
    print("mcu", mx, "dc", int(ref[b, 0]), int(ref[b + 1, 0]), int(ref[b + 2, 0]), "nnz", int((ref[b, 1:] != 0).sum()), int((ref[b+1, 1:] != 0).sum()), int((ref[b+2, 1:] != 0).sum()))
