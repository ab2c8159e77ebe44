# final evidence of the round: tests, bench lines, launch list, full captures of the three kernels that carry the step
mkdir -p gpurun_out/r2b  # run from the repo root: gpurun -- bash tools/gpu_scripts/<this file>
O=gpurun_out/r2b
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.txt 2>&1; tail -3 $O/pytest_gpu.txt
( time timeout 1500 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err ) 2> $O/bench_n1.time; tail -3 $O/bench_n1.time
timeout 900 python bench.py --impl reference > $O/bench_reference_arm.json 2> $O/bench_reference_arm.err
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu --no-extra --precision fp64 > $O/bench_n1_fp64.json 2> $O/bench_n1_fp64.err
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu --no-extra --ngal 2000000 > $O/bench_n1_2M.json 2> $O/bench_n1_2M.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 27 -c 30 --csv --log-file $O/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-extra > $O/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"lane_walk_kernel|lane_build_kernel|lane_finish_kernel|lane_plan_kernel" -s 12 -c 4 -o $O/lane_full -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-extra > $O/ncu_full.log 2>&1; tail -2 $O/ncu_full.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"lane_walk_kernel" -s 3 -c 1 -o $O/walk_fp64_full -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-extra --precision fp64 > $O/ncu_full_fp64.log 2>&1; tail -2 $O/ncu_full_fp64.log
for c in b30_rf3_defaults b40_hd_cs17 b4_defaults; do timeout 900 python tools/parity_sweep.py $c 100000 > $O/parity_sweep_$c.txt 2>&1; grep -E "max rel" $O/parity_sweep_$c.txt | cut -c1-170; done
