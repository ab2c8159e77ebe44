#!/usr/bin/env python
"""profiles/r2_lane_kernel_counters.json from an `ncu --set full` report of one step of bench.py (the four kernels of the
lane pipeline): DRAM bytes per step and per kernel for roofline.traffic, and the pipe utilisations of the walk kernel (the
dominant one) for the bench line's `pipes`.  usage: ncu_counters_json.py rep.ncu-rep NGAL"""
import csv
import json
import subprocess
import sys

rep, ngal = sys.argv[1], int(sys.argv[2])
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, units = rows[0], rows[1]


def val(r, name):
    v = float(r[h.index(name)].replace(",", ""))
    u = units[h.index(name)]
    return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}.get(u, 1.0)


kern, tot_r, tot_w, tot_t = {}, 0.0, 0.0, 0.0
walk = None
for r in rows[2:]:
    name = r[h.index("Kernel Name")]
    short = next((k for k in ("lane_plan_kernel", "lane_build_kernel", "lane_walk_kernel", "lane_finish_kernel") if k in name), name)
    rd, wr, t = val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum"), val(r, "gpu__time_duration.sum")
    kern[short] = {"dram_bytes_read": rd, "dram_bytes_write": wr, "duration_s_under_ncu": t}
    tot_r += rd; tot_w += wr; tot_t += t
    if short == "lane_walk_kernel":
        walk = r
d = {"ngal": ngal, "workload": "cfg2", "precision": "fp32c", "kernels": kern,
     "dram_bytes_read": tot_r, "dram_bytes_write": tot_w, "dram_bytes_per_launch": tot_r + tot_w,
     "dominant_kernel": "lane_walk_kernel", "dominant_kernel_share_of_step": kern["lane_walk_kernel"]["duration_s_under_ncu"] / tot_t,
     "dominant_kernel_dram_bytes": kern["lane_walk_kernel"]["dram_bytes_read"] + kern["lane_walk_kernel"]["dram_bytes_write"]}
if walk is not None:
    wf = val(walk, "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum")
    d.update({"issue_active_pct": val(walk, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
              "lsu_pipe_pct": val(walk, "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
              "smem_bank_conflict_frac": val(walk, "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum") / max(wf, 1.0),
              "fp64_pipe_pct": val(walk, "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
              "fma_pipe_pct": val(walk, "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
              "alu_pipe_pct": val(walk, "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active")})
d["source"] = ("ncu --set full --clock-control none -k regex:lane_(plan|build|walk|finish)_kernel of one step of "
               "`python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-extra` (tools/gpu_scripts/gpu_round2_s.sh); the pipe numbers are the "
               "walk kernel's, the DRAM bytes the sum over the step's four kernels")
d["note"] = ("the DRAM traffic above the catalog columns (0.114 GB) is the rest-frame SED rows [node][32 galaxies] of every batch: written "
             "once by the build kernel, read by every band table whose window covers them (about 30 KB per galaxy read in total); "
             "the SEDs of the batches in flight do not fit the 126 MB L2")
print(json.dumps(d, indent=1))
