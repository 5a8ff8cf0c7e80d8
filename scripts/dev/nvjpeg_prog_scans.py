import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from nvjpeg_imagecompressor_b200.synth import synth
L = C.CDLL(os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "baseline", "_ref", "libref_nvjpeg.so"))
W, H = 1024, 768
img = synth(W, H, 3).cpu().numpy()
for css in (0, 1):
    h = C.c_void_p()
    assert L.ref_create(W, H, 95, 1, css, 1, C.byref(h)) == 0
    assert L.ref_build_compress_env(h) == 0
    out = np.empty(W * H * 3, np.uint8); n = C.c_size_t(0)
    assert L.ref_compress(h, C.c_void_p(img.ctypes.data), C.c_size_t(W * 3), C.c_void_p(out.ctypes.data), C.c_size_t(out.size), C.byref(n)) == 0
    b = out[: n.value]
    p = 2
    print("css", css, "bytes", n.value)
    while p + 4 <= b.size:
        if b[p] != 0xFF: break
        m = int(b[p + 1]); p += 2
        if m in (0x01,) or 0xD0 <= m <= 0xD8: continue
        if m == 0xD9: break
        Lh = (int(b[p]) << 8) | int(b[p + 1])
        if m == 0xDA:
            nc = int(b[p + 2]); Ss, Se, A = int(b[p + 3 + 2 * nc]), int(b[p + 4 + 2 * nc]), int(b[p + 5 + 2 * nc])
            # find end of segment
            e = p + Lh
            while e + 1 < b.size and not (b[e] == 0xFF and b[e + 1] != 0 and not (0xD0 <= b[e + 1] <= 0xD7)): e += 1
            print("  SOS ncomp", nc, "comps", [int(b[p + 3 + 2 * i]) for i in range(nc)], "Ss", Ss, "Se", Se, "Ah", A >> 4, "Al", A & 15, "bytes", e - (p + Lh))
            p = e; continue
        if m == 0xDD: print("  DRI", (int(b[p + 2]) << 8) | int(b[p + 3]))
        p += Lh
    L.ref_destroy(h)
