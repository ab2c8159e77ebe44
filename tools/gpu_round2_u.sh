mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -k "maglim" > gpurun_out/pytest_new.txt 2>&1; tail -15 gpurun_out/pytest_new.txt
