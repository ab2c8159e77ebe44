mkdir -p gpurun_out/r2b
df -h /tmp | tail -1
timeout 1500 python -m pytest tests/test_cli.py tests/test_gpu_parity.py -m gpu -q -k "cli or save_sed or config5 or input_cat or full_run" > gpurun_out/pytest_new.txt 2>&1; tail -5 gpurun_out/pytest_new.txt
timeout 1200 python tools/config5_run.py 50000 > gpurun_out/r2b/config5_cli.txt 2>&1; tail -4 gpurun_out/r2b/config5_cli.txt
