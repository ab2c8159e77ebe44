#!/usr/bin/env python
"""Key numbers of one kernel from an ncu report: `python tools/ncu_summary.py rep.ncu-rep [launch]`
(time, instructions, issue utilisation, pipe utilisation, stall reasons, memory traffic)."""
import csv
import subprocess
import sys

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, u = rows[0], rows[1]
v = rows[2 + (int(sys.argv[2]) if len(sys.argv) > 2 else 0)]  # optional second argument: which launch of the report
print("kernel:", v[h.index("Kernel Name")])
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
keys2 = ["l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum",
         "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum",
         "l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_st.sum"]
for k in keys2:
    if k in d:
        print(f"{k:75s} {d[k][0]:>16s} {d[k][1]}")
# warp states from the PC samples (the per-issue ratios of newer ncu versions carry the same information)
st, tot = [], 0.0
for n, (val, unit) in d.items():
    if n.startswith("smsp__pcsamp_warps_issue_stalled_") and not n.endswith("_not_issued"):
        try:
            st.append((float(val.replace(",", "")), n))
            tot += float(val.replace(",", ""))
        except ValueError:
            pass
print("--- warp states (% of PC samples)")
for val, n in sorted(st, reverse=True)[:12]:
    print(f"{100 * val / max(tot, 1):8.2f}  {n.replace('smsp__pcsamp_warps_issue_stalled_', '')}")
