"""Attribute an ncu SASS source-page CSV to CUDA source lines using nvdisasm line info.

    ncu -i rep.ncu-rep --page source --csv --print-source sass > sass.csv
    cuobjdump -xelf all libsimt_b200.so ; nvdisasm -g -c head.sm_100a.cubin > head.disasm
    python scripts/ncu_lines.py sass.csv head.disasm <mangled-kernel-substring> [top]
"""
import csv
import re
import sys
from collections import defaultdict

sass_csv, disasm, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40

# ---- line table from nvdisasm ------------------------------------------------------------------
lines = open(disasm).read().split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith(".text.") and kern in l and l.rstrip().endswith(":"))
line_of = []          # per instruction (in order)
cur = None
ins_re = re.compile(r"^\s+/\*([0-9a-f]{4,})\*/\s+(.*?);")
for l in lines[start + 1:]:
    if l.startswith("//-----") or l.startswith(".section") and "text" in l:
        break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = int(m.group(2)) if m.group(1).endswith("head.cu") or m.group(1).endswith(".cu") else cur
        continue
    m = ins_re.match(l)
    if m:
        line_of.append((cur, m.group(2)))

rows = list(csv.reader(open(sass_csv)))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
col = {n: i for i, n in enumerate(hdr)}
data = rows[hdr_i + 1:]
print(f"{len(data)} SASS rows in report, {len(line_of)} instructions in disasm")
by_line = defaultdict(lambda: [0, 0, defaultdict(int)])
by_op = defaultdict(lambda: [0, 0])
tot_inst = tot_samp = 0
stall_cols = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
for i, r in enumerate(data):
    ln = line_of[i][0] if i < len(line_of) else None
    inst = int(r[col["Instructions Executed"]] or 0)
    samp = int(r[col["# Samples"]] or 0)
    op = r[col["Source"]].split()[0] if r[col["Source"]].split() else "?"
    if op.startswith("@"):
        op = r[col["Source"]].split()[1]
    op = op.split(".")[0]
    by_line[ln][0] += inst
    by_line[ln][1] += samp
    for s in stall_cols:
        v = int(r[col[s]] or 0)
        if v:
            by_line[ln][2][s] += v
    by_op[op][0] += inst
    by_op[op][1] += samp
    tot_inst += inst
    tot_samp += samp
src = open("/root/repo/simt_b200/csrc/head.cu").read().split("\n") if "head" in disasm else []
print(f"total warp-instructions {tot_inst}, samples {tot_samp}")
print("\n== by source line (sorted by samples) ==")
for ln, (inst, samp, st) in sorted(by_line.items(), key=lambda kv: -kv[1][1])[:top]:
    tops = ", ".join(f"{k[6:]}={v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
    text = src[ln - 1].strip()[:70] if ln and ln <= len(src) else ""
    print(f"L{ln}: inst {100*inst/tot_inst:5.1f}%  samples {100*samp/tot_samp:5.1f}%  [{tops}]  | {text}")
print("\n== by opcode ==")
for op, (inst, samp) in sorted(by_op.items(), key=lambda kv: -kv[1][0])[:30]:
    print(f"{op:10s} inst {100*inst/tot_inst:5.1f}% ({inst})  samples {100*samp/tot_samp:5.1f}%")
