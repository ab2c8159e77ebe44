mkdir -p gpurun_out/r2b
FUZZ_GPU=1 timeout 900 python tools/fuzz_filters.py 3 220 > gpurun_out/r2b/fuzz_gpu_plain.txt 2>&1; tail -3 gpurun_out/r2b/fuzz_gpu_plain.txt
FUZZ_GPU=1 timeout 900 python tools/fuzz_filters.py 4 220 wide > gpurun_out/r2b/fuzz_gpu_wide.txt 2>&1; tail -3 gpurun_out/r2b/fuzz_gpu_wide.txt
FUZZ_GPU=1 timeout 900 python tools/fuzz_filters.py 5 300 wide > gpurun_out/r2b/fuzz_gpu_wide2.txt 2>&1; tail -3 gpurun_out/r2b/fuzz_gpu_wide2.txt
