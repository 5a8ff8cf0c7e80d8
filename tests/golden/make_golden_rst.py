#!/usr/bin/env python
"""Generates tests/golden/golden_rst.json: digests of cv2.imencode streams WITH restart markers
(IMWRITE_JPEG_RST_INTERVAL, libjpeg-turbo 3.1.2) and of their cv2.imdecode pixels, for the restart-interval half of
the checker (oracle.encode(..., restart_interval=n)). Run from the repo root:
    python tests/golden/make_golden_rst.py
"""
import json
import os
import sys

import cv2

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import oracle as O  # noqa: E402
from make_golden import cv_encode, sha  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    cv2.setNumThreads(1)
    cases = []
    for (W, H) in ((96, 80), (17, 33), (135, 121), (640, 360)):
        seed = W * 31 + H
        img = O.synth(W, H, seed=seed, amp=8)
        for css in range(5):
            g = O.geometry(W, H, css)
            for q, opt in ((95, 1), (75, 0)):
                for rst in (1, 5, g.mcux, 2 * g.mcux, 60000):   # 1 MCU, ragged, one / two MCU rows, longer than the image
                    jpg = cv_encode(img, css, q, opt, rst)
                    dec = cv2.imdecode(jpg, cv2.IMREAD_COLOR)
                    cases.append(dict(W=W, H=H, seed=seed, amp=8, css=css, quality=q, optimize=opt, restart_interval=int(rst),
                                      jpeg_len=int(jpg.size), jpeg_sha256_128=sha(jpg)[:32], decoded_sha256_128=sha(dec)[:32]))
    with open(os.path.join(HERE, "golden_rst.json"), "w") as f:
        json.dump(dict(generator="tests/golden/make_golden_rst.py, cv2 %s" % cv2.__version__, cases=cases), f, separators=(",", ":"))
    print(len(cases), "cases")


if __name__ == "__main__":
    main()
