#!/usr/bin/env python
"""Generates tests/golden/golden.json and tests/golden/*.jpg with the LIVE bit-exactness oracle of the
north-star: cv2.imencode / cv2.imdecode (OpenCV 4.13.0 wheel, libjpeg-turbo 3.1.2, DCT_ISLOW).

Inputs come from the integer synthetic generator of SURVEY.md Appendix B (oracle.synth), so only digests
need to be committed. Run from the repo root:  python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

import cv2
import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import oracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
SF = {0: cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444, 1: cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422,
      2: cv2.IMWRITE_JPEG_SAMPLING_FACTOR_440, 3: cv2.IMWRITE_JPEG_SAMPLING_FACTOR_420,
      4: cv2.IMWRITE_JPEG_SAMPLING_FACTOR_411}


def cv_encode(img, css, q, opt, rst=0):
    p = [cv2.IMWRITE_JPEG_QUALITY, q, cv2.IMWRITE_JPEG_OPTIMIZE, int(opt), cv2.IMWRITE_JPEG_SAMPLING_FACTOR, SF[css]]
    if rst:
        p += [cv2.IMWRITE_JPEG_RST_INTERVAL, rst]
    ok, b = cv2.imencode(".jpg", img, p)
    assert ok
    return b.ravel()


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    cv2.setNumThreads(1)
    cases = []
    sizes = [(64, 96), (48, 64), (50, 70), (17, 33), (135, 121), (8, 8), (1, 1), (257, 63), (640, 360), (1920, 1080)]
    for (W, H) in sizes:
        seed = W * 31 + H
        img = O.synth(W, H, seed=seed, amp=8)
        for css in range(5):
            for q, opt in ((95, 1), (95, 0), (75, 1), (100, 0)):
                if W * H > 100000 and (q, opt) not in ((95, 1), (95, 0)):
                    continue
                jpg = cv_encode(img, css, q, opt)
                dec = cv2.imdecode(jpg, cv2.IMREAD_COLOR)
                cases.append(dict(W=W, H=H, seed=seed, amp=8, css=css, quality=q, optimize=opt, jpeg_len=int(jpg.size),
                                  jpeg_sha256=sha(jpg), decoded_sha256=sha(dec),
                                  psnr=float(cv2.PSNR(img, dec))))
    # three tiny files committed verbatim (decode fixtures that do not need the generator)
    files = []
    for (W, H, css, q, opt) in ((50, 70, 1, 95, 1), (135, 121, 3, 95, 0), (33, 17, 4, 75, 1)):
        img = O.synth(W, H, seed=7, amp=8)
        jpg = cv_encode(img, css, q, opt)
        name = f"synth_{W}x{H}_css{css}_q{q}_opt{opt}.jpg"
        jpg.tofile(os.path.join(HERE, name))
        dec = cv2.imdecode(jpg, cv2.IMREAD_COLOR)
        files.append(dict(file=name, W=W, H=H, seed=7, amp=8, css=css, quality=q, optimize=opt,
                          jpeg_sha256=sha(jpg), decoded_sha256=sha(dec)))
    # headline-size known answers measured by the survey (full 8320x40000 image)
    headline = dict(
        image=dict(W=8320, H=40000, seed=0, amp=8,
                   raw_sha256="1998cc55af5f5559357771c352637589ae4aa408e1e1b65196a5da2be4b02ddd"),
        encodes=[dict(css=3, quality=95, optimize=0, jpeg_len=127740823, jpeg_sha256_128="a4262baf890ecced73c1ccd0e6e220eb",
                      decoded_sha256_128="14c6bda4b603fea0c4e8d67e10572e95"),
                 dict(css=1, quality=95, optimize=1, jpeg_len=148073301, jpeg_sha256_128="25313424729423734b295197a59dc3fb",
                      decoded_sha256_128="307604274a930a494b5ef8d98cf7cf22"),
                 dict(css=0, quality=95, optimize=1, jpeg_len=220395670, jpeg_sha256_128="5e4fa049843b45b596525b55c7c362de",
                      decoded_sha256_128="50fafa9e24829e7b6885c0ea1c63096c")])
    # first 2000 rows of the headline image (cheap enough for the GPU suite to compare against)
    big = O.synth(8320, 2000, 0, 8)
    slab = []
    for css, q, opt in ((1, 95, 1), (3, 95, 0), (0, 95, 1), (2, 85, 1), (4, 75, 0)):
        jpg = cv_encode(big, css, q, opt)
        dec = cv2.imdecode(jpg, cv2.IMREAD_COLOR)
        slab.append(dict(W=8320, H=2000, seed=0, amp=8, css=css, quality=q, optimize=opt, jpeg_len=int(jpg.size),
                         jpeg_sha256=sha(jpg), decoded_sha256=sha(dec), psnr=float(cv2.PSNR(big, dec))))
    out = dict(generator="cv2 %s / %s" % (cv2.__version__, [l.strip() for l in cv2.getBuildInformation().splitlines()
                                                           if "JPEG:" in l][0]),
               cases=cases, files=files, headline=headline, slab=slab)
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", len(cases), "cases,", len(files), "files,", len(slab), "slab cases")


if __name__ == "__main__":
    main()
