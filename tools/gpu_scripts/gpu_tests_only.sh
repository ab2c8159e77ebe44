mkdir -p gpurun_out/r2b
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2b/pytest_gpu.txt 2>&1; tail -3 gpurun_out/r2b/pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
