"""The CPU oracle (oracle/jpeg_oracle.c) against committed golden vectors made by libjpeg-turbo 3.1.2
(tests/golden/make_golden.py) and, when cv2 is importable, against the live library."""
import hashlib
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_generator_known_answer(oracle):
    # row 0, first 8 pixels and the last pixel of the headline image
    row0 = oracle.synth(8320, 40000, 0, 8, y0=0, rows=1)
    assert row0[0, :8].ravel().tolist() == [239, 183, 156, 250, 182, 164, 245, 184, 169, 239, 185, 172, 243, 186, 164,
                                            240, 172, 167, 238, 174, 180, 223, 168, 171]
    last = oracle.synth(8320, 40000, 0, 8, y0=39999, rows=1)
    assert last[0, 8319].tolist() == [128, 76, 43]


def test_encode_matches_golden(oracle, golden):
    for c in golden["cases"]:
        if c["W"] * c["H"] > 700000:
            continue
        img = oracle.synth(c["W"], c["H"], c["seed"], c["amp"])
        jpg = oracle.encode(img, c["css"], c["quality"], c["optimize"])
        assert jpg.size == c["jpeg_len"], c
        assert sha(jpg) == c["jpeg_sha256"], c


def test_decode_matches_golden(oracle, golden):
    for f in golden["files"]:
        jpg = np.fromfile(os.path.join(HERE, "golden", f["file"]), np.uint8)
        assert sha(jpg) == f["jpeg_sha256"]
        dec = oracle.decode(jpg)
        assert dec.shape == (f["H"], f["W"], 3)
        assert sha(dec) == f["decoded_sha256"], f
        # and the same file is what the oracle's encoder makes
        img = oracle.synth(f["W"], f["H"], f["seed"], f["amp"])
        assert np.array_equal(oracle.encode(img, f["css"], f["quality"], f["optimize"]), jpg)


def test_roundtrip_decode_golden_digests(oracle, golden):
    for c in golden["cases"]:
        if c["W"] * c["H"] > 100000 or c["quality"] != 95:
            continue
        img = oracle.synth(c["W"], c["H"], c["seed"], c["amp"])
        dec = oracle.decode(oracle.encode(img, c["css"], c["quality"], c["optimize"]))
        assert sha(dec) == c["decoded_sha256"], c
        assert abs(oracle.psnr(img, dec) - c["psnr"]) < 1e-9


def test_1080p_golden(oracle, golden):
    c = [c for c in golden["cases"] if c["W"] == 1920 and c["css"] == 3 and c["optimize"] == 1][0]
    img = oracle.synth(c["W"], c["H"], c["seed"], c["amp"])
    jpg = oracle.encode(img, c["css"], c["quality"], c["optimize"])
    assert sha(jpg) == c["jpeg_sha256"]
    assert sha(oracle.decode(jpg)) == c["decoded_sha256"]


def test_stage_functions_compose(oracle):
    """forward -> histogram -> tables -> entropy -> stuff -> headers == encode (the per-stage entry points the
    GPU parity tests compare against)."""
    img = oracle.synth(100, 52, 3, 8)
    for css in range(5):
        for opt in (0, 1):
            g = oracle.geometry(100, 52, css)
            coef = oracle.forward(img, css, 90)
            if opt:
                h = oracle.histogram(coef, g.bpm)
                tb = [oracle.gen_optimal_table(h[t]) for t in range(4)]
                bits = np.stack([t[0] for t in tb])
                vals = np.stack([t[1] for t in tb])
            else:
                bits, vals = oracle.std_tables()
            raw, nbits = oracle.entropy_bits(coef, g.bpm, bits, vals)
            body = oracle.stuff(raw, nbits)
            hdr = oracle.headers(100, 52, css, oracle.quant_tables(90), bits, vals)
            whole = np.concatenate([hdr, body, np.array([0xFF, 0xD9], np.uint8)])
            assert np.array_equal(whole, oracle.encode(img, css, 90, opt))
            coef2, info = oracle.decode_coefs(whole)
            assert np.array_equal(coef2, coef)


def test_strip_predictors_compose(oracle):
    """Entropy bits of two MCU-row strips (second one seeded with the first's last DCs) concatenate to the whole."""
    img = oracle.synth(96, 64, 5, 8)
    css = 1
    g = oracle.geometry(96, 64, css)
    coef = oracle.forward(img, css, 95)
    bits, vals = oracle.std_tables()
    raw, nbits = oracle.entropy_bits(coef, g.bpm, bits, vals)
    half = (g.mcuy // 2) * g.mcux * g.bpm
    a, na = oracle.entropy_bits(coef[:half], g.bpm, bits, vals)
    last = coef[:half].reshape(-1, g.bpm, 64)[-1]
    pred = np.array([last[g.bpm - 3, 0], last[g.bpm - 2, 0], last[g.bpm - 1, 0]], np.int16)
    b, nb = oracle.entropy_bits(coef[half:], g.bpm, bits, vals, pred_in=pred)
    assert na + nb == nbits
    ab = np.unpackbits(a)[:na]
    bb = np.unpackbits(b)[:nb]
    assert np.array_equal(np.concatenate([ab, bb]), np.unpackbits(raw)[:nbits])
    h = oracle.histogram(coef, g.bpm)
    h2 = oracle.histogram(coef[:half], g.bpm) + oracle.histogram(coef[half:], g.bpm, pred_in=pred)
    assert np.array_equal(h, h2)


def test_fibonacci_length_limit(oracle):
    """Frequencies that force natural code lengths > 16 (jpeg_gen_optimal_table's limiting loop)."""
    freq = np.zeros(257, np.uint32)
    f = [1, 1]
    while len(f) < 30:
        f.append(f[-1] + f[-2])
    freq[1:31] = f
    bits, vals, n = oracle.gen_optimal_table(freq)
    assert n == 30 and bits[1:].sum() == 30
    assert bits[16] > 0 and bits[17:].sum() == 0 if len(bits) > 17 else True
    # Kraft inequality strictly below 1 (all-ones code reserved)
    assert sum(int(b) << (16 - l) for l, b in enumerate(bits) if l) < (1 << 16)


def test_live_cv2_sweep(oracle):
    cv2 = pytest.importorskip("cv2")
    cv2.setNumThreads(1)
    sf = {0: cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444, 1: cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422,
          2: cv2.IMWRITE_JPEG_SAMPLING_FACTOR_440, 3: cv2.IMWRITE_JPEG_SAMPLING_FACTOR_420,
          4: cv2.IMWRITE_JPEG_SAMPLING_FACTOR_411}
    rng = np.random.default_rng(11)
    for (W, H) in ((31, 47), (16, 16), (2, 2), (3, 5), (90, 40), (8, 9)):
        for img in (oracle.synth(W, H, 1, 8), rng.integers(0, 256, (H, W, 3), dtype=np.uint8)):
            for css in range(5):
                for q, opt in ((95, 1), (60, 0), (100, 1)):
                    ok, ref = cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_QUALITY, q, cv2.IMWRITE_JPEG_OPTIMIZE, opt,
                                                         cv2.IMWRITE_JPEG_SAMPLING_FACTOR, sf[css]])
                    ref = ref.ravel()
                    assert np.array_equal(oracle.encode(img, css, q, opt), ref), (W, H, css, q, opt)
                    assert np.array_equal(oracle.decode(ref), cv2.imdecode(ref, cv2.IMREAD_COLOR)), (W, H, css, q, opt)
    # restart-interval streams (row N2): decode only
    img = oracle.synth(70, 50, 2, 8)
    ok, ref = cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_QUALITY, 90, cv2.IMWRITE_JPEG_RST_INTERVAL, 3,
                                         cv2.IMWRITE_JPEG_SAMPLING_FACTOR, sf[3]])
    assert np.array_equal(oracle.decode(ref.ravel()), cv2.imdecode(ref, cv2.IMREAD_COLOR))


def test_restart_interval_golden(oracle):
    """Row N2's checker half: streams with DRI / RSTn (jchuff.c emit_restart) == cv2 IMWRITE_JPEG_RST_INTERVAL digests,
    and their decode == cv2.imdecode digests (tests/golden/make_golden_rst.py)."""
    import json
    with open(os.path.join(HERE, "golden", "golden_rst.json")) as f:
        cases = json.load(f)["cases"]
    assert len(cases) >= 100
    for c in cases:
        img = oracle.synth(c["W"], c["H"], c["seed"], c["amp"])
        jpg = oracle.encode(img, c["css"], c["quality"], c["optimize"], c["restart_interval"])
        assert jpg.size == c["jpeg_len"] and sha(jpg)[:32] == c["jpeg_sha256_128"], c
        if c["W"] * c["H"] < 100000:
            assert sha(oracle.decode(jpg))[:32] == c["decoded_sha256_128"], c


def test_restart_interval_live_cv2(oracle):
    cv2 = pytest.importorskip("cv2")
    cv2.setNumThreads(1)
    sf = {0: cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444, 1: cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422,
          2: cv2.IMWRITE_JPEG_SAMPLING_FACTOR_440, 3: cv2.IMWRITE_JPEG_SAMPLING_FACTOR_420,
          4: cv2.IMWRITE_JPEG_SAMPLING_FACTOR_411}
    rng = np.random.default_rng(5)
    for (W, H) in ((50, 70), (8, 8), (33, 90)):
        for img in (oracle.synth(W, H, 4, 8), rng.integers(0, 256, (H, W, 3), dtype=np.uint8)):
            for css in range(5):
                for q, opt, ri in ((95, 1, 1), (80, 0, 4), (100, 1, 7), (90, 1, 500)):
                    ok, ref = cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_QUALITY, q, cv2.IMWRITE_JPEG_OPTIMIZE, opt,
                                                         cv2.IMWRITE_JPEG_SAMPLING_FACTOR, sf[css],
                                                         cv2.IMWRITE_JPEG_RST_INTERVAL, ri])
                    ref = ref.ravel()
                    assert np.array_equal(oracle.encode(img, css, q, opt, ri), ref), (W, H, css, q, opt, ri)
                    assert np.array_equal(oracle.decode(ref), cv2.imdecode(ref, cv2.IMREAD_COLOR)), (W, H, css, q, opt, ri)


def test_progressive_decode_golden(oracle):
    """Row N4's checker half: progressive (SOF2) streams, the format the reference as shipped writes
    (ImageCompressorImpl.cu:28), decode to cv2.imdecode's pixels (fixtures: tests/golden/make_golden_prog.py)."""
    import json
    with open(os.path.join(HERE, "golden", "golden_prog.json")) as f:
        files = json.load(f)["files"]
    assert len(files) >= 5
    for c in files:
        jpg = np.fromfile(os.path.join(HERE, "golden", c["file"]), np.uint8)
        assert sha(jpg) == c["jpeg_sha256"]
        dec = oracle.decode(jpg)
        assert dec.shape == (c["H"], c["W"], 3) and sha(dec) == c["decoded_sha256"], c
        if c["restart_interval"] == 0:   # and the same file is what the checker's progressive encoder writes
            img = oracle.synth(c["W"], c["H"], 11, 8)
            assert np.array_equal(oracle.encode_progressive(img, c["css"], c["quality"]), jpg), c


def test_progressive_encode_live_cv2(oracle):
    """jcphuff.c restated (jpeg_simple_progression, per-scan optimal tables): bytes == cv2 IMWRITE_JPEG_PROGRESSIVE."""
    cv2 = pytest.importorskip("cv2")
    cv2.setNumThreads(1)
    sf = {0: cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444, 1: cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422,
          2: cv2.IMWRITE_JPEG_SAMPLING_FACTOR_440, 3: cv2.IMWRITE_JPEG_SAMPLING_FACTOR_420,
          4: cv2.IMWRITE_JPEG_SAMPLING_FACTOR_411}
    rng = np.random.default_rng(9)
    for (W, H) in ((50, 70), (8, 8), (1, 1), (135, 121), (257, 63)):
        for img in (oracle.synth(W, H, 6, 8), rng.integers(0, 256, (H, W, 3), dtype=np.uint8)):
            for css in range(5):
                for q in (50, 95, 100):
                    ok, ref = cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_QUALITY, q, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, sf[css],
                                                         cv2.IMWRITE_JPEG_PROGRESSIVE, 1])
                    assert np.array_equal(oracle.encode_progressive(img, css, q), ref.ravel()), (W, H, css, q)


def test_progressive_decode_live_cv2(oracle):
    cv2 = pytest.importorskip("cv2")
    cv2.setNumThreads(1)
    sf = {0: cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444, 1: cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422,
          2: cv2.IMWRITE_JPEG_SAMPLING_FACTOR_440, 3: cv2.IMWRITE_JPEG_SAMPLING_FACTOR_420,
          4: cv2.IMWRITE_JPEG_SAMPLING_FACTOR_411}
    rng = np.random.default_rng(8)
    for (W, H) in ((50, 70), (8, 8), (1, 1), (135, 121)):
        for img in (oracle.synth(W, H, 6, 8), rng.integers(0, 256, (H, W, 3), dtype=np.uint8)):
            for css in range(5):
                for q, opt, ri in ((95, 1, 0), (50, 0, 0), (100, 1, 7)):
                    p = [cv2.IMWRITE_JPEG_QUALITY, q, cv2.IMWRITE_JPEG_OPTIMIZE, opt, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, sf[css],
                         cv2.IMWRITE_JPEG_PROGRESSIVE, 1] + ([cv2.IMWRITE_JPEG_RST_INTERVAL, ri] if ri else [])
                    ok, ref = cv2.imencode(".jpg", img, p)
                    ref = ref.ravel()
                    assert np.array_equal(oracle.decode(ref), cv2.imdecode(ref, cv2.IMREAD_COLOR)), (W, H, css, q, opt, ri)


def test_diff_psnr(oracle):
    rng = np.random.default_rng(3)
    a = rng.integers(0, 256, (40, 30, 3), dtype=np.uint8)
    b = rng.integers(0, 256, (40, 30, 3), dtype=np.uint8)
    d = a.astype(int) - b.astype(int)
    assert np.array_equal(oracle.diff(a, b, 0), np.abs(d).astype(np.uint8))
    assert np.array_equal(oracle.diff(a, b, 1), np.clip(d + 128, 0, 255).astype(np.uint8))
    assert oracle.ssd(a, b) == int((d * d).sum())
    cv2 = pytest.importorskip("cv2")
    assert abs(oracle.psnr(a, b) - cv2.PSNR(a, b)) < 1e-12  # same formula; libm vs OpenCV log10 differ by <=1 ulp
