"""Per-kernel table (duration, DRAM bytes, GB/s, % of the measured HBM peak) from an ncu --csv log of
scripts/all_kernels_once.py:   python scripts/ncu_kernel_table.py gpurun_out/all_kernels.csv profiles/r1_all_kernels_ncu.txt"""
import collections
import csv
import json
import os
import re
import sys

src, dst = sys.argv[1], sys.argv[2]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    PEAK = 6559.7
rows = list(csv.reader(open(src)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr = rows[hi]
ix = {n: i for i, n in enumerate(hdr)}
per = collections.OrderedDict()
unit = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3}
for r in rows[hi + 1:]:
    if len(r) < len(hdr) or not r[0].isdigit():
        continue
    name = re.sub(r"\(.*", "", r[ix["Kernel Name"]]).replace("void ", "").replace("simt::", "")
    key = (name, r[ix["Grid Size"]].replace(" ", ""))
    v = float(r[ix["Metric Value"]].replace(",", "")) * unit.get(r[ix["Metric Unit"]], 1)
    per.setdefault(key, collections.defaultdict(list))[r[ix["Metric Name"]]].append(v)
out = ["# every kernel of libsimt_b200.so under ncu (--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum",
       "# --clock-control none; scripts/all_kernels_once.py: B=8, 19+4 channels, 65x129 -> 512x1024; eval kernels on 8 images of",
       "# 1024x2048).  Cold caches, serialised launches.  DRAM GB/s = (read + write) / duration, % of the measured HBM peak",
       f"# ({PEAK} GB/s); means over the launches of the run.  MODE: 0 fwd, 1 fwd+bwd, 3 Placeholder, 4 step (scale on device).",
       f"{'kernel':64s} {'grid':>12s} {'n':>3s} {'us':>9s} {'DRAM MB':>9s} {'GB/s':>8s} {'% peak':>7s}"]
for (name, grid), d in per.items():
    n = len(d["gpu__time_duration.sum"])
    t = sum(d["gpu__time_duration.sum"]) / n
    by = (sum(d["dram__bytes_read.sum"]) + sum(d["dram__bytes_write.sum"])) / n
    gbs = by / t / 1e3
    out.append(f"{name[:64]:64s} {grid:>12s} {n:3d} {t:9.2f} {by / 1e6:9.2f} {gbs:8.0f} {100 * gbs / PEAK:6.1f}%")
open(dst, "w").write("\n".join(out) + "\n")
print("\n".join(out))
