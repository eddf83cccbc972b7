"""Warp-stall reason breakdown + pipe utilisation of an .ncu-rep (run here): python scripts/ncu_stalls.py rep.ncu-rep"""
import csv, subprocess, sys, io
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
st = {k: float(v[0].replace(',', '')) for k, v in d.items()
      if k.startswith('smsp__pcsamp_warps_issue_stalled_') and not k.endswith('_not_issued')}
tot = sum(st.values())
print("stall samples:", ", ".join(f"{k.replace('smsp__pcsamp_warps_issue_stalled_', '')} {v / tot * 100:.1f}%"
                                  for k, v in sorted(st.items(), key=lambda x: -x[1])[:12]))
for k in ['sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
          'sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
          'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
          'sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
          'smsp__inst_executed.sum', 'gpu__time_duration.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
          'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
          'smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct', 'dram__bytes_read.sum', 'dram__bytes_write.sum']:
    if k in d:
        print(f"  {k}: {d[k][0]} {d[k][1]}")
