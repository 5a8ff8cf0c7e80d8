"""Small end-to-end pass over every kernel family for compute-sanitizer (memcheck / racecheck)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import nvjpeg_imagecompressor_b200 as P
import oracle as O
ok = True
for (W, H, css) in ((200, 136, 1), (129, 77, 3), (64, 48, 0), (257, 63, 4), (96, 250, 2)):
    img = O.synth(W, H, 5, 8)
    g = O.geometry(W, H, css)
    for q, opt in ((95, 1), (100, 0)):
        eng = P.Engine(W, H, q, bool(opt), css)
        jpg = eng.encode(img)
        ok &= np.array_equal(jpg, O.encode(img, css, q, opt))
        rec = eng.decode(jpg)
        ok &= np.array_equal(rec, O.decode(jpg))
        eng.set_debug(4)                       # short schedule -> checked retry -> k_dec_sync_long
        ok &= np.array_equal(eng.decode(jpg), rec)
        eng.set_debug(0)
        eng.set_restart_rows(1)
        jr = eng.encode(img)
        ok &= np.array_equal(jr, O.encode(img, css, q, opt, g.mcux))
        ok &= np.array_equal(eng.decode(jr), rec)
        eng.set_restart_rows(0)
        jp = O.encode_progressive(img, css, q)
        ok &= np.array_equal(eng.decode(jp), O.decode(jp))
        d_img = torch.from_numpy(img).cuda()
        eng.secondary_device(d_img.data_ptr(), W * 3, W, H, 1)
        n1, n2, ps, ssd = eng.secondary_finish()
        ok &= n1 == jpg.size and ssd == O.ssd(img, rec)
        eng.close()
    m = P.MultiEngine(W, H, 95, True, css, devices=[0, 0, 0])
    ok &= np.array_equal(m.encode(img), O.encode(img, css, 95, 1))
    m.close()
print("sanitize pass ok" if ok else "MISMATCH")
sys.exit(0 if ok else 1)
