mkdir -p gpurun_out
for c in b30_rf3_defaults b40_hd_cs17; do timeout 600 python tools/parity_sweep.py $c 20000 > gpurun_out/p_$c.txt 2>&1; echo "rc=$?" >> gpurun_out/p_$c.txt; grep -E "max rel|rc=" gpurun_out/p_$c.txt | cut -c1-150; done
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-extra > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; tail -2 gpurun_out/bench_quick.err
python -c "import json;d=json.load(open('gpurun_out/bench_quick.json'));print('quick %.4g %.4g'%(d['value'], d['e2e']['value']), d['roofline']['kernel_ms'], d['roofline']['kernel_share_of_step'])"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-extra --precision fp64 > gpurun_out/bench_fp64.json 2> gpurun_out/bench_fp64.err; python -c "import json;d=json.load(open('gpurun_out/bench_fp64.json'));print('fp64 %.4g'%(d['value']))"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 27 -c 10 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-extra > gpurun_out/ncu_launches.log 2>&1; tail -8 gpurun_out/launches.csv | cut -d\" -f10,30 | cut -c1-60,100-200
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1; tail -5 gpurun_out/pytest_gpu.txt
