#!/usr/bin/env python
"""Turn an .ncu-rep (ncu --set full of the encode kernels) into the committed evidence under profiles/:
   <tag>_raw.csv      ncu --page raw --csv (every metric of every captured launch)
   <tag>_summary.txt  key metrics, stall reasons and hottest source lines per kernel (scripts/ncu_summary.py)
   traffic.json       dram__bytes_read.sum + dram__bytes_write.sum per launch, read by bench.py (roofline.traffic)

    python scripts/ncu_export.py gpurun_out/prof.ncu-rep r01g
"""
import csv, io, json, os, subprocess, sys
rep, tag = sys.argv[1], sys.argv[2]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
open(os.path.join(root, "profiles", f"{tag}_raw.csv"), "w").write(raw)
summ = subprocess.run([sys.executable, os.path.join(root, "scripts", "ncu_summary.py"), rep, "30"], capture_output=True, text=True).stdout
open(os.path.join(root, "profiles", f"{tag}_summary.txt"), "w").write(summ)
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
def to_bytes(v, u):
    return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
def to_ms(v, u):
    return float(v) * {"ns": 1e-6, "us": 1e-3, "ms": 1, "s": 1e3, "usecond": 1e-3, "msecond": 1, "nsecond": 1e-6}[u]
kern = {}
for r in rows[2:]:
    name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "").strip()
    name = {"k_fdct<2, 1, 0>": "k_fdct<2,1>"}.get(name, name)
    rd = to_bytes(r[ix["dram__bytes_read.sum"]], units[ix["dram__bytes_read.sum"]])
    wr = to_bytes(r[ix["dram__bytes_write.sum"]], units[ix["dram__bytes_write.sum"]])
    kern[name] = {"dram_bytes": int(rd + wr), "dram_read": int(rd), "dram_write": int(wr),
                  "ncu_ms": round(to_ms(r[ix["gpu__time_duration.sum"]], units[ix["gpu__time_duration.sum"]]), 4),
                  "warp_instructions": int(float(r[ix["smsp__inst_executed.sum"]])),
                  "issue_active_pct": round(float(r[ix["smsp__issue_active.avg.pct_of_peak_sustained_active"]]), 1)}
import hashlib
_h = hashlib.sha256()   # the encode kernels' sources, as bench.py kernel_source_digest()
for _f in [os.path.join(root, "nvjpeg_imagecompressor_b200", "csrc", n) for n in ("common.cuh", "enc_fdct.cu", "enc_huff.cu")]:
    _h.update(os.path.basename(_f).encode() + b"\0" + open(_f, "rb").read())
json.dump({"source": f"profiles/{tag}_raw.csv (ncu --set full --clock-control none, headline config, one launch each)",
           "source_digest": _h.hexdigest()[:16],   # bench.py reports `traffic` only for the code this was captured on
           "kernels": kern}, open(os.path.join(root, "profiles", "traffic.json"), "w"), indent=1)
print(json.dumps(kern, indent=1))
