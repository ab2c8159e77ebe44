set -x
mkdir -p gpurun_out
for ch in 32768 98304; do
EGGPHOT_CHUNK=$ch EGGPHOT_TRACE=1 timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu --no-extra > gpurun_out/trace_$ch.json 2> gpurun_out/trace_$ch.err; tail -8 gpurun_out/trace_$ch.err
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"lane_walk_kernel" -s 3 -c 1 -o gpurun_out/walk_full2 -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-extra > gpurun_out/ncu_full2.log 2>&1; tail -3 gpurun_out/ncu_full2.log
