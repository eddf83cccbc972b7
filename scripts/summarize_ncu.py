"""Turn an .ncu-rep into the text summary committed under profiles/ (run here, no GPU needed).
   python scripts/summarize_ncu.py gpurun_out/x.ncu-rep profiles/x.txt [kernel-substr-for-line-attribution]"""
import subprocess, sys, os, re
rep, out = sys.argv[1], sys.argv[2]
kern = sys.argv[3] if len(sys.argv) > 3 else None
det = subprocess.run(["ncu", "-i", rep, "--page", "details"], capture_output=True, text=True).stdout
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
keep = re.compile(r"Duration|Elapsed Cycles|SM Frequency|SM Active Cycles|DRAM Throughput|Memory Throughput|Compute \(SM\) Throughput|"
                  r"Executed Ipc|Issue Slots Busy|Registers Per|Theoretical Occ|Achieved Occ|Active Warps Per Scheduler|Eligible Warps|"
                  r"Avg. Active Threads|Executed Instructions  |Grid Size|Block Size|Dynamic Shared|L1/TEX Hit|L2 Hit|Mem Busy|Max Bandwidth|"
                  r"^\s+void |^\s+[a-z_]+kernel")
lines = [l.rstrip() for l in det.split("\n") if keep.search(l)]
metrics = {}
rows = [r for r in raw.split("\n") if r.startswith('"')]
if len(rows) >= 3:
    import csv, io
    rd = list(csv.reader(io.StringIO("\n".join(rows))))
    hdr, units, vals = rd[0], rd[1], rd[2]
    want = ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_lsu.sum",
            "sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
            "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum"]
    for w in want:
        for i, h in enumerate(hdr):
            if h == w:
                metrics[w] = f"{vals[i]} {units[i]}"
with open(out, "w") as f:
    f.write(f"# ncu --set full --clock-control none summary of {os.path.basename(rep)}\n")
    f.write("\n".join(lines) + "\n\n# selected raw metrics\n")
    for k, v in metrics.items():
        f.write(f"{k}: {v}\n")
    if kern:
        sass = out + ".sass.csv"
        with open(sass, "w") as g:
            g.write(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout)
        dis = "build/scratch/elf/head.disasm"
        la = subprocess.run([sys.executable, "scripts/ncu_lines.py", sass, dis, kern, "25"], capture_output=True, text=True).stdout
        f.write("\n# warp-stall samples and executed instructions by source line / opcode (scripts/ncu_lines.py)\n" + la)
        os.remove(sass)
print(open(out).read()[:3000])
