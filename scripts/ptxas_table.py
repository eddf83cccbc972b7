"""Parse `nvcc -Xptxas=-v` output (stdin or file) into one line per kernel: registers, spills, smem.
   nvcc ... -Xptxas=-v -c x.cu 2>&1 | python scripts/ptxas_table.py [filter]"""
import re, subprocess, sys
txt = sys.stdin.read()
flt = sys.argv[1] if len(sys.argv) > 1 else ""
cur = None
rows = []
for line in txt.splitlines():
    m = re.search(r"Compiling entry function '(\S+)'", line)
    if m:
        cur = {"name": m.group(1), "spill_st": 0, "spill_ld": 0, "regs": 0, "stack": 0}
        rows.append(cur)
        continue
    if cur is None:
        continue
    m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
    if m:
        cur["stack"], cur["spill_st"], cur["spill_ld"] = map(int, m.groups())
    m = re.search(r"Used (\d+) registers", line)
    if m:
        cur["regs"] = int(m.group(1))
names = [r["name"] for r in rows]
try:
    dem = subprocess.run(["cu++filt"] + names, capture_output=True, text=True).stdout.splitlines()
except Exception:
    dem = names
for r, d in zip(rows, dem):
    d = d.replace("simt::", "").replace("(simt::HeadArgs)", "").replace("unsigned char", "u8").replace("long long", "i64")
    if flt and flt not in d:
        continue
    print(f"{r['regs']:4d} regs  spill st/ld {r['spill_st']:4d}/{r['spill_ld']:4d}  stack {r['stack']:4d}  {d}")
