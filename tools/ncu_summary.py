#!/usr/bin/env python
"""Key numbers of one kernel from an ncu report: `python tools/ncu_summary.py rep.ncu-rep`
(time, instructions, issue utilisation, pipe utilisation, stall reasons, memory traffic)."""
import csv
import subprocess
import sys

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, u, v = rows[0], rows[1], rows[2]
d = {n: (v[i], u[i]) for i, n in enumerate(h)}
keys = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_eligible.avg.per_cycle_active", "smsp__warps_active.avg.per_cycle_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "lts__t_sectors_srcunit_tex.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
for k in keys:
    if k in d:
        print(f"{k:75s} {d[k][0]:>16s} {d[k][1]}")
st = []
for n, (val, unit) in d.items():
    if "issue_stalled" in n and n.endswith("_per_warp_active.pct") and "not_issued" not in n:
        try:
            st.append((float(val), n))
        except ValueError:
            pass
print("--- stall reasons (% of warp-active cycles)")
for val, n in sorted(st, reverse=True)[:12]:
    print(f"{val:8.2f}  {n.replace('smsp__average_warps_issue_stalled_', '').replace('_per_warp_active.pct', '')}")
