set -x
mkdir -p gpurun_out
timeout 600 python tools/parity_sweep.py b30_rf3_defaults 20000 > gpurun_out/p_b30.txt 2>&1; echo "rc=$?" >> gpurun_out/p_b30.txt; grep -E "max rel|rc=" gpurun_out/p_b30.txt | cut -c1-150
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-extra > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; tail -2 gpurun_out/bench_quick.err; head -c 200 gpurun_out/bench_quick.json
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --no-extra --ngal 2000000 > gpurun_out/bench_2M.json 2> gpurun_out/bench_2M.err; head -c 200 gpurun_out/bench_2M.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"lane_walk_kernel|lane_build_kernel" -s 6 -c 2 -o gpurun_out/walk_full -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-extra > gpurun_out/ncu_full.log 2>&1; tail -3 gpurun_out/ncu_full.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 20 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-extra > gpurun_out/ncu_launches.log 2>&1; tail -10 gpurun_out/launches.csv | cut -d\" -f10,30
