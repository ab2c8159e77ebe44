mkdir -p gpurun_out/r2b
timeout 1500 python tools/config5_run.py 195000 > gpurun_out/r2b/config5_cli_full.txt 2>&1; tail -4 gpurun_out/r2b/config5_cli_full.txt
