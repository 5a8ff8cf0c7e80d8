"""Compare the device token pool with tokens derived from the checker's coefficients (development aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import nvjpeg_imagecompressor_b200 as P
from nvjpeg_imagecompressor_b200 import _native as N
import oracle as O

def nbits(v):
    return int(abs(int(v))).bit_length()

def vbits(v, nb):
    v = int(v)
    return (v + (-1 if v < 0 else 0)) & ((1 << nb) - 1)

def ref_tokens(coef, bpm, hv, mcux, mcuy, tm, pred0):
    """tokens per tile, final format (after k_dc_edge_hist resolved the raw DCs)"""
    tiles = []
    pred = list(pred0)
    tiles_x = (mcux + tm - 1) // tm
    for my in range(mcuy):
        for tx in range(tiles_x):
            toks = []
            for mx in range(tx * tm, min(mcux, (tx + 1) * tm)):
                for bn in range(bpm):
                    b = (my * mcux + mx) * bpm + bn
                    c = 0 if bn < hv else bn - hv + 1
                    tbl = 0 if bn < hv else 1
                    blk = coef[b]
                    d = int(blk[0]) - pred[c]
                    pred[c] = int(blk[0])
                    nb = nbits(d)
                    toks.append(((tbl * 2) << 20) | (nb << 16) | vbits(d, nb))
                    run = 0
                    for k in range(1, 64):
                        z = int(blk[k])
                        if z == 0:
                            run += 1
                            continue
                        nb = nbits(z)
                        toks.append(((run >> 4) << 26) | ((run & 15) << 22) | ((tbl * 2 + 1) << 20) | (nb << 16) | vbits(z, nb))
                        run = 0
                    if run:
                        toks.append((tbl * 2 + 1) << 20)
            tiles.append(toks)
    return tiles

for (W, H, css) in ((48, 64, 0), (64, 96, 0), (135, 121, 1)):
    q = 95
    eng = P.Engine(300, 160, q, True, css)
    img = O.synth(W, H, W * 31 + H, 8)
    eng.set_debug(1)
    jpg = eng.encode(img)
    g = O.geometry(W, H, css)
    ref = O.forward(img, css, q)
    pool = eng.debug_read(N.DBG_TOKENS, np.uint32)
    recs = eng.debug_read(N.DBG_TILE_RECS, np.uint8).reshape(-1, 24)
    hv = g.bpm - 2
    tmax = (256 // g.bpm) & ~1
    tiles_x = (g.mcux + tmax - 1) // tmax
    tm = (g.mcux + tiles_x - 1) // tiles_x
    tm = min(tmax, (tm + 1) & ~1)
    want = ref_tokens(ref, g.bpm, hv, g.mcux, g.mcuy, tm, (0, 0, 0))
    nbad = 0
    for t, toks in enumerate(want):
        base, count = recs[t, :8].view(np.uint32)
        got = pool[base:base + count].tolist()
        if got != toks:
            nbad += 1
            if nbad <= 3:
                print(f"tile {t}: count {count} want {len(toks)}")
                for i, (a, b) in enumerate(zip(got, toks)):
                    if a != b:
                        print("   got ", [f"{x:08x}" for x in got[max(0,i-6):i+12]])
                        print("   want", [f"{x:08x}" for x in toks[max(0,i-6):i+12]])
                        nd = sum(1 for a2, b2 in zip(got, toks) if a2 != b2)
                        print("   ndiff in tile", nd)
                        print(f"   first diff at {i}: got {a:08x} want {b:08x}; neighbours got {[hex(x) for x in got[max(0,i-2):i+3]]} want {[hex(x) for x in toks[max(0,i-2):i+3]]}")
                        break
    print(W, H, css, "tiles", len(want), "bad", nbad)
    eng.close()
