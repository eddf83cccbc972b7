"""Executed-instruction profile of an ncu SASS source page CSV (no line info needed):
   ncu -i rep --page source --csv --print-source sass > sass.csv ; python scripts/ncu_sass_hot.py sass.csv [nregions]
Prints totals by opcode and the instruction stream as regions of equal execution count (basic-block-like)."""
import csv, sys, re
from collections import defaultdict
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
col = {n: i for i, n in enumerate(rows[hi])}
data = rows[hi + 1:]
tot = sum(int(r[col["Instructions Executed"]]) for r in data)
samp = sum(int(r[col["# Samples"]]) for r in data)
print(f"warp-instructions executed {tot}, samples {samp}, SASS rows {len(data)}")
byop = defaultdict(lambda: [0, 0])
for r in data:
    src = r[col["Source"]].strip()
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_]+)", src)
    op = m.group(2) if m else src.split()[0]
    byop[op][0] += int(r[col["Instructions Executed"]]); byop[op][1] += int(r[col["# Samples"]])
print("by opcode:", ", ".join(f"{k} {v[0] / tot * 100:.1f}%" for k, v in sorted(byop.items(), key=lambda x: -x[1][0])[:28]))
# regions
regs = []
cur = None
for i, r in enumerate(data):
    n = int(r[col["Instructions Executed"]])
    if cur is None or abs(n - cur[2]) > 0.02 * max(n, cur[2], 1):
        cur = [i, i, n, 0, 0]
        regs.append(cur)
    cur[1] = i; cur[3] += n; cur[4] += int(r[col["# Samples"]])
print("regions (first row, #instr, exec count per instr, share of executed, share of samples):")
k = int(sys.argv[2]) if len(sys.argv) > 2 else 40
for g in sorted(regs, key=lambda g: -g[3])[:k]:
    print(f"  row {g[0]:5d} len {g[1] - g[0] + 1:4d}  x{g[2]:8d}  inst {g[3] / tot * 100:5.1f}%  samples {g[4] / max(samp, 1) * 100:5.1f}%   {data[g[0]][col['Source']].strip()[:50]}")
