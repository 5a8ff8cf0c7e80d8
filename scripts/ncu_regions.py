#!/usr/bin/env python
"""Instruction share of k_fdct's stages from an .ncu-rep (regions found by the '// ----' markers in enc_fdct.cu)."""
import collections, csv, io, re, subprocess, sys
rep = sys.argv[1]; srcfile = sys.argv[2] if len(sys.argv) > 2 else "nvjpeg_imagecompressor_b200/csrc/enc_fdct.cu"
kern = sys.argv[3] if len(sys.argv) > 3 else "k_fdct"
nblocks = float(sys.argv[4]) if len(sys.argv) > 4 else 10.4e6
lines = open(srcfile).read().split("\n")
marks = [(1, "file head / helpers")]
for i, l in enumerate(lines, 1):
    m = re.match(r"\s*// ---- (.*)", l)
    if m: marks.append((i, m.group(1)[:50]))
    if re.match(r"__global__|template <int HS, int VS, bool DUMP>\s*$", l): marks.append((i, "kernel prologue"))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
fn = fname = hdr = None; acc = collections.Counter(); other = 0
for r in csv.reader(io.StringIO(src)):
    if len(r) >= 2 and r[0] == "Function Name": fn = r[1]
    if len(r) >= 2 and r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if len(r) > 8 and r[0] == "Line No": hdr = r; continue
    if hdr and len(r) == len(hdr) and r[2] == "-" and fn and kern in fn:
        try: n = int(r[hdr.index("Instructions Executed")]); ln = int(r[0])
        except ValueError: continue
        if fname == srcfile.split("/")[-1]: acc[ln] += n
        else: other += n
tot = sum(acc.values()) + other
print(f"total warp instructions {tot}, per block {tot*32/nblocks:.0f} thread-instr")
for j, (l0, name) in enumerate(marks):
    l1 = marks[j + 1][0] if j + 1 < len(marks) else 10**9
    s = sum(n for l, n in acc.items() if l0 <= l < l1)
    if s: print(f"  {s/tot*100:5.1f}%  {s*32/nblocks:7.0f}/block  lines {l0}-{l1-1}: {name}")
print(f"  {other/tot*100:5.1f}%  {other*32/nblocks:7.0f}/block  other files")
