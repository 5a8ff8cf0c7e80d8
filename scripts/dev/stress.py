import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import nvjpeg_imagecompressor_b200 as P
from nvjpeg_imagecompressor_b200 import _native as N
import oracle as O
rng = np.random.default_rng(1)
nfail = 0; ntot = 0
for css in range(5):
    for q, opt in ((95, 1), (75, 0), (100, 1)):
        eng = P.Engine(300, 160, q, bool(opt), css)
        for trial in range(12):
            W = int(rng.integers(1, 300)); H = int(rng.integers(1, 160))
            img = O.synth(W, H, W * 31 + H, 8)
            want = O.encode(img, css, q, opt)
            for rep in range(3):
                jpg = eng.encode(img)
                ntot += 1
                ok = jpg.size == want.size and bool(np.array_equal(jpg, want))
                if not ok:
                    nfail += 1
                    recs = eng.debug_read(N.DBG_TILE_RECS, np.uint8).reshape(-1, 24)
                    a = [int(x) & 3 for x in recs[:, :4].view(np.uint32).ravel()]
                    cnt = [int(x) for x in recs[:, 4:8].view(np.uint32).ravel()]
                    n = min(jpg.size, want.size)
                    d = np.nonzero(jpg[:n] != want[:n])[0]
                    if nfail <= 12:
                        print(f"FAIL css{css} q{q} opt{opt} {W}x{H} rep{rep} len {jpg.size}/{want.size} firstdiff {d[:1]} a={a[:10]} cnt={cnt[:10]}")
        eng.close()
print("total", ntot, "fail", nfail)
