mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"lane_walk_kernel" -s 3 -c 1 -o gpurun_out/walk_small -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-extra --ngal 24384 > gpurun_out/ncu_small.log 2>&1; tail -3 gpurun_out/ncu_small.log
