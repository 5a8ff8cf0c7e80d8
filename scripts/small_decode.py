#!/usr/bin/env python
"""Where a 1920x1080 decode spends its time: host call durations and device stage times, 1 and 4 engines."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import nvjpeg_imagecompressor_b200 as P
from nvjpeg_imagecompressor_b200.synth import synth_rows
W, H = 1920, 1080
dev = torch.device("cuda", 0)
img = synth_rows(W, H, 0, H, 3, 8, dev)
for neng, pinned in ((1, False), (4, False), (8, False), (8, True), (16, True)):
    engs = [P.Engine(W, H, 95, True, "420") for _ in range(neng)]
    streams = [torch.cuda.Stream() for _ in engs]
    for e, s in zip(engs, streams):
        e.set_stream(s.cuda_stream)
    jpg = np.array(engs[0].encode(img.cpu().numpy()), copy=True)
    if pinned:
        keep = torch.empty(jpg.size, dtype=torch.uint8, pin_memory=True)
        keep.numpy()[:] = jpg
        jpg = keep.numpy()
    outs = [torch.empty((H, W, 3), dtype=torch.uint8, device=dev) for _ in engs]
    for k, e in enumerate(engs):
        e.decode_device(jpg, outs[k].data_ptr(), W * 3); e.decode_finish()
    N = 400
    t_call = t_fin = 0.0
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(N):
        e = engs[i % neng]
        if i >= neng:
            a = time.perf_counter(); e.decode_finish(); t_fin += time.perf_counter() - a
        a = time.perf_counter(); e.decode_device(jpg, outs[i % neng].data_ptr(), W * 3); t_call += time.perf_counter() - a
    for e in engs:
        e.decode_finish()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    engs[0].enable_timing(True)
    engs[0].decode_device(jpg, outs[0].data_ptr(), W * 3); engs[0].decode_finish()
    tm = {k: round(v, 4) for k, v in engs[0].timings().items() if k.startswith("dec_") or k == "total"}
    print(json.dumps(dict(engines=neng, pinned=pinned, images_s=round(N / dt, 1), us_per_image=round(dt / N * 1e6, 1),
                          host_us_decode_device=round(t_call / N * 1e6, 1), host_us_decode_finish=round(t_fin / N * 1e6, 1),
                          jpeg_bytes=int(jpg.size), stages_ms=tm)))
    for e in engs:
        e.close()
