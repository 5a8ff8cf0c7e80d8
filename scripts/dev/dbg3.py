import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import nvjpeg_imagecompressor_b200 as P
from nvjpeg_imagecompressor_b200 import _native as N
import oracle as O
W, H, css, q, opt = 48, 64, 0, 95, 1
img = O.synth(W, H, W * 31 + H, 8)
eng = P.Engine(300, 160, q, bool(opt), css)
jpg = eng.encode(img)
pool = eng.debug_read(N.DBG_TOKENS, np.uint32)
recs = eng.debug_read(N.DBG_TILE_RECS, np.uint8).reshape(-1, 24)
tb = eng.debug_read(N.DBG_TILE_BITS, np.uint32)
t = eng.tables()
enc = np.array([[t.enc[i][j] for j in range(256)] for i in range(4)], dtype=np.uint64)
def tok_len(tk):
    tbl = (tk >> 20) & 3; nb = (tk >> 16) & 15; run = (tk >> 22) & 15; nz = (tk >> 26) & 3
    sym = (run << 4 | nb) if tbl & 1 else nb
    l = int(enc[tbl][sym]) & 0xFF
    zl = int(enc[tbl][0xF0]) & 0xFF if tbl & 1 else 0
    return l + nb + nz * zl
for ti in range(recs.shape[0]):
    base, count = recs[ti, :8].view(np.uint32)
    want = sum(tok_len(int(x)) for x in pool[base:base + count])
    print("tile", ti, "count", count, "a", base & 3, "bits", int(tb[ti]), "want", want, "" if want == tb[ti] else "<<<<")
w2 = O.encode(img, css, q, opt)
n = min(jpg.size, w2.size)
d = np.nonzero(jpg[:n] != w2[:n])[0]
print("len", jpg.size, w2.size, "first diff", d[:5], "hdr", int(t.hdr_len))
if d.size:
    i = d[0]
    print(" got ", jpg[i-4:i+12].tolist()); print(" want", w2[i-4:i+12].tolist())
