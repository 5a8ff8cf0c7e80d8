#!/usr/bin/env python
"""Summarise an .ncu-rep on the CPU box: key raw metrics per kernel and the hottest CUDA source lines per kernel.

    python scripts/ncu_summary.py gpurun_out/prof.ncu-rep [top_lines] [kernel_substring]
"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
only = sys.argv[3] if len(sys.argv) > 3 else ""

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum.pct_of_peak_sustained_elapsed",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__grid_size",
        "smsp__warps_eligible.avg.per_cycle_active", "lts__t_bytes.sum"]
STALLS = "smsp__average_warps_issue_stalled_"

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
for r in rows[2:]:
    name = r[idx["Kernel Name"]].split("(")[0]
    if only and only not in name:
        continue
    print("==", name)
    for w in WANT:
        if w in idx:
            print(f"   {w:90s} {r[idx[w]]} {units[idx[w]]}")
    st = [(float(r[i]), h[len(STALLS):-len('_per_issue_active.ratio')]) for h, i in idx.items()
          if h.startswith(STALLS) and h.endswith("_per_issue_active.ratio") and "not_issued" not in h]
    print("   stalls/issue:", ", ".join(f"{n}={v:.2f}" for v, n in sorted(st, reverse=True)[:8]))

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
kern = fname = hdr = None
acc = collections.defaultdict(lambda: collections.defaultdict(lambda: [0, 0, 0, ""]))
for r in csv.reader(io.StringIO(src)):
    if len(r) >= 2 and r[0] == "Function Name":
        kern = r[1].split("(")[0]
    if len(r) >= 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if len(r) > 8 and r[0] == "Line No":
        hdr = r
        continue
    if hdr and len(r) == len(hdr) and r[2] == "-":
        ii, ti, si = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
        try:
            e = acc[kern][(fname, int(r[0]))]
            e[0] += int(r[ii]); e[1] += int(r[ti]); e[2] += int(r[si]); e[3] = r[1][:100]
        except ValueError:
            pass
for k, v in acc.items():
    if only and only not in k:
        continue
    tot = sum(e[0] for e in v.values()); stot = sum(e[2] for e in v.values())
    print(f"== {k}: {tot} warp instructions, {stot} samples")
    for (f, l), e in sorted(v.items(), key=lambda kv: -kv[1][2])[:top]:
        print(f"  {e[0] / max(tot, 1) * 100:5.1f}% inst {e[2] / max(stot, 1) * 100:5.1f}% smp thr/inst={e[1] / max(e[0], 1):5.1f} {f}:{l}: {e[3]}")
