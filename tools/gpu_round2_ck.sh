mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_cli.py -m gpu -q -k "check_mode or input_cat_generates" > gpurun_out/pytest_new.txt 2>&1; tail -30 gpurun_out/pytest_new.txt
