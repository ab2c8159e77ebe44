# last check of the round with the committed code: GPU tests, smoke, the driver's bench invocation, the parity report
mkdir -p gpurun_out/r2b
O=gpurun_out/r2b
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.txt 2>&1; tail -3 $O/pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.txt 2>&1; tail -3 $O/smoke.txt
timeout 900 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; python -c "import json;d=json.load(open('$O/bench_n1.json'));print('%.4g %.4g'%(d['value'], d['e2e']['value']))"
timeout 1200 python tools/parity_report.py 4000 > $O/parity_report.json 2> $O/parity_report.err; tail -2 $O/parity_report.err
