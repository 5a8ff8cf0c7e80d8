import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import nvjpeg_imagecompressor_b200 as P
from nvjpeg_imagecompressor_b200 import _native as N
import oracle as O
css, q, opt = 0, 95, 1
eng = P.Engine(300, 160, q, bool(opt), css)
def bits_of(pool, enc):
    out = []
    for tk in pool:
        tk = int(tk)
        tbl = (tk >> 20) & 3; nb = (tk >> 16) & 15; run = (tk >> 22) & 15; nz = (tk >> 26) & 3
        sym = (run << 4 | nb) if tbl & 1 else nb
        for _ in range(nz):
            e = int(enc[tbl][0xF0]); out.append((e >> 8, e & 0xFF))
        e = int(enc[tbl][sym]); out.append((e >> 8, e & 0xFF))
        if nb: out.append((tk & 0xFFFF, nb))
    return out
def pack_words(pairs):
    acc = 0; n = 0
    for v, l in pairs:
        acc = (acc << l) | v; n += l
    pad = (-n) % 32
    acc <<= pad
    return [(acc >> (32 * i)) & 0xFFFFFFFF for i in range((n + pad) // 32 - 1, -1, -1)], n
nbad = 0
for (W, H) in ((64, 96), (11, 23), (26, 5), (75, 50), (77, 66), (48, 64), (11, 23), (26, 5)):
    img = O.synth(W, H, W * 31 + H, 8)
    want = O.encode(img, css, q, opt)
    jpg = eng.encode(img)
    ok = jpg.size == want.size and bool(np.array_equal(jpg, want))
    pool = eng.debug_read(N.DBG_TOKENS, np.uint32)
    recs = eng.debug_read(N.DBG_TILE_RECS, np.uint8).reshape(-1, 24)
    slots = eng.debug_read(N.DBG_SLOTS, np.uint32).reshape(-1, 13312)
    tb = eng.debug_read(N.DBG_TILE_BITS, np.uint32)
    t = eng.tables()
    enc = [[t.enc[i][j] for j in range(256)] for i in range(4)]
    print(W, H, "ok" if ok else "FAIL")
    for ti in range(recs.shape[0]):
        base, count = recs[ti, :8].view(np.uint32)
        words, n = pack_words(bits_of(pool[base:base + count], enc))
        got = slots[ti, :len(words)].tolist()
        if n % 32:
            m = (0xFFFFFFFF << (32 - n % 32)) & 0xFFFFFFFF
            got[-1] &= m
        bad = [i for i, (a, b) in enumerate(zip(got, words)) if a != b]
        if bad or n != tb[ti]:
            print(f"  tile {ti}: a={base & 3} count={count} bits {tb[ti]}/{n} bad words {bad[:8]} of {len(words)}")
            for i in bad[:3]:
                print(f"     word {i}: got {got[i]:08x} want {words[i]:08x} xor {got[i]^words[i]:08x}")
eng.close()
