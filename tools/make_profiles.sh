#!/bin/sh
# Turns the scratch outputs of tools/gpu_scripts/gpu_round2_s.sh (gpurun_out/r2b) into the tracked summaries under profiles/.
# usage: tools/make_profiles.sh [tag]   (tag defaults to r2b)
set -e
cd "$(dirname "$0")/.."
T=${1:-r2b}; R=gpurun_out/r2b; P=profiles
for f in bench_n1 bench_n1_fp64 bench_n1_2M bench_reference_arm; do cp $R/$f.json $P/${T}_$f.json; done
cp $R/pytest_gpu.txt $P/${T}_pytest_gpu.txt
cp $R/launches_bench.csv $P/${T}_launches_bench.csv
for c in b30_rf3_defaults b40_hd_cs17 b4_defaults; do grep -E "max rel|max abs" $R/parity_sweep_$c.txt; done > $P/${T}_parity_sweep_1e5.txt
i=0
for k in plan build walk finish; do
  python tools/ncu_summary.py $R/lane_full.ncu-rep $i > $P/${T}_${k}_kernel_summary.txt
  i=$((i+1))
done
python tools/ncu_summary.py $R/walk_fp64_full.ncu-rep 0 > $P/${T}_walk_kernel_fp64_summary.txt
ncu -i $R/lane_full.ncu-rep --page details --kernel-name regex:lane_walk_kernel > $P/${T}_walk_kernel_details.txt
ncu -i $R/lane_full.ncu-rep --page details --kernel-name regex:lane_build_kernel > $P/${T}_build_kernel_details.txt
ncu -i $R/lane_full.ncu-rep --page source --print-source cuda,sass --csv --kernel-name regex:lane_walk_kernel | python tools/ncu_by_line.py 60 > $P/${T}_walk_kernel_by_line.txt
ncu -i $R/lane_full.ncu-rep --page source --print-source cuda,sass --csv --kernel-name regex:lane_build_kernel | python tools/ncu_by_line.py 30 > $P/${T}_build_kernel_by_line.txt
python tools/ncu_counters_json.py $R/lane_full.ncu-rep 195000 > $P/r2_lane_kernel_counters.json
