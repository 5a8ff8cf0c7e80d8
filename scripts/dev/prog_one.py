import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import nvjpeg_imagecompressor_b200 as P
import oracle as O
W, H, css = 1920, 1080, 3
img = O.synth(W, H, 2, 8)
jpg = O.encode_progressive(img, css, 95)
eng = P.Engine(W, H, 95, True, css)
for i in range(2):
    t0 = time.perf_counter(); out = eng.decode(jpg); print("ms", (time.perf_counter() - t0) * 1e3)
