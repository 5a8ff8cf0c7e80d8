#!/usr/bin/env python
"""Generates tests/golden/prog_*.jpg + golden_prog.json: small PROGRESSIVE (SOF2) streams written by cv2.imencode
(libjpeg-turbo 3.1.2, jpeg_simple_progression) with the digests of cv2.imdecode's pixels -- fixtures of the checker's
progressive decoder (oracle.decode, SURVEY.md 8f N4: the format the reference as shipped writes). Run from the repo
root:  python tests/golden/make_golden_prog.py
"""
import json
import os
import sys

import cv2

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import oracle as O  # noqa: E402
from make_golden import SF, sha  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    cv2.setNumThreads(1)
    files = []
    for (W, H, css, q, opt, rst) in ((50, 70, 1, 95, 1, 0), (33, 17, 4, 75, 0, 0), (64, 48, 3, 90, 1, 5), (40, 40, 0, 100, 1, 0),
                                     (24, 56, 2, 60, 0, 3)):
        img = O.synth(W, H, seed=11, amp=8)
        p = [cv2.IMWRITE_JPEG_QUALITY, q, cv2.IMWRITE_JPEG_OPTIMIZE, opt, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, SF[css],
             cv2.IMWRITE_JPEG_PROGRESSIVE, 1] + ([cv2.IMWRITE_JPEG_RST_INTERVAL, rst] if rst else [])
        ok, jpg = cv2.imencode(".jpg", img, p)
        assert ok
        jpg = jpg.ravel()
        name = f"prog_{W}x{H}_css{css}_q{q}_opt{opt}_rst{rst}.jpg"
        jpg.tofile(os.path.join(HERE, name))
        dec = cv2.imdecode(jpg, cv2.IMREAD_COLOR)
        files.append(dict(file=name, W=W, H=H, css=css, quality=q, optimize=opt, restart_interval=rst, jpeg_sha256=sha(jpg),
                          decoded_sha256=sha(dec)))
    with open(os.path.join(HERE, "golden_prog.json"), "w") as f:
        json.dump(dict(generator="tests/golden/make_golden_prog.py, cv2 %s" % cv2.__version__, files=files), f, indent=1)
    print(len(files), "files")


if __name__ == "__main__":
    main()
