set -x
mkdir -p gpurun_out
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_lane_512.json 2> gpurun_out/bench_lane_512.err; tail -2 gpurun_out/bench_lane_512.err
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --ngal 2000000 > gpurun_out/bench_lane_2M_512.json 2> gpurun_out/bench_lane_2M_512.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --precision fp64 > gpurun_out/bench_lane_fp64.json 2> gpurun_out/bench_lane_fp64.err
EGGPHOT_KERNEL=node timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_node.json 2> gpurun_out/bench_node.err
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1; tail -5 gpurun_out/pytest_gpu.txt
for cs in "b30_rf3_defaults 300000" "b40_hd_cs17 100000" "b4_flambda_nonorm 300000"; do timeout 900 python tools/parity_sweep.py $cs >> gpurun_out/parity_sweep_3e5.txt 2>&1; done; cat gpurun_out/parity_sweep_3e5.txt
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_run.py 4096 > gpurun_out/sanitizer_memcheck.txt 2>&1; tail -5 gpurun_out/sanitizer_memcheck.txt
timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_run.py 2048 > gpurun_out/sanitizer_racecheck.txt 2>&1; tail -5 gpurun_out/sanitizer_racecheck.txt
timeout 900 compute-sanitizer --tool synccheck python tools/sanitize_run.py 2048 > gpurun_out/sanitizer_synccheck.txt 2>&1; tail -5 gpurun_out/sanitizer_synccheck.txt
