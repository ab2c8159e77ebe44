# refresh of the round's bench lines with the final code (1 GPU)
mkdir -p gpurun_out/r2b
O=gpurun_out/r2b
( time timeout 1500 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err ) 2> $O/bench_n1.time; tail -3 $O/bench_n1.time
timeout 900 python bench.py --impl reference > $O/bench_reference_arm.json 2> $O/bench_reference_arm.err
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu --no-extra --precision fp64 > $O/bench_n1_fp64.json 2> $O/bench_n1_fp64.err
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu --no-extra --ngal 2000000 > $O/bench_n1_2M.json 2> $O/bench_n1_2M.err
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.txt 2>&1; tail -3 $O/pytest_gpu.txt
python -c "
import json
d=json.load(open('$O/bench_n1.json')); print('n1 %.4g e2e %.4g'%(d['value'], d['e2e']['value']), d['roofline']['kernel'], d['roofline']['kernel_ms'], d['roofline']['kernel_share_of_step'], d['roofline']['frac'], d['pipes'])
"
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.txt 2>&1; tail -4 $O/smoke.txt
