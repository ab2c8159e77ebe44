set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1; tail -5 gpurun_out/pytest_gpu.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-extra --precision fp64 > gpurun_out/bench_fp64.json 2> gpurun_out/bench_fp64.err; tail -2 gpurun_out/bench_fp64.err; head -c 200 gpurun_out/bench_fp64.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lane_build_kernel -s 3 -c 1 -o gpurun_out/build_full -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-extra > gpurun_out/ncu_build.log 2>&1; tail -3 gpurun_out/ncu_build.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lane_finish_kernel -s 3 -c 1 -o gpurun_out/finish_full -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-extra > gpurun_out/ncu_finish.log 2>&1; tail -3 gpurun_out/ncu_finish.log
