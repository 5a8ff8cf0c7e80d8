#!/usr/bin/env python
"""Aggregate an `ncu --page source --print-source cuda,sass --csv` dump by CUDA source line."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
fname = None; hdr = None; data = []
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if len(r) > 8 and r[0] == "Line No": hdr = r; continue
    if hdr and len(r) == len(hdr) and r[2] == "-":   # a source-line summary row
        ii = hdr.index("Instructions Executed"); ti = hdr.index("Thread Instructions Executed"); si = hdr.index("# Samples")
        try: data.append((int(r[ii]), int(r[ti]), int(r[si]), fname, r[0], r[1][:100]))
        except ValueError: pass
tot = sum(d[0] for d in data); stot = sum(d[2] for d in data)
print("total warp instructions", tot, "samples", stot)
for n, t, s, f, l, src in sorted(data, reverse=True)[:top]:
    print(f"{n/tot*100:5.1f}% inst {s/max(stot,1)*100:5.1f}% smp thr/inst={t/max(n,1):5.1f} {f}:{l}: {src}")
