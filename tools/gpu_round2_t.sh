mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "node_sets or config5" > gpurun_out/pytest_new.txt 2>&1; tail -15 gpurun_out/pytest_new.txt
