mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-extra $EXTRA > gpurun_out/bench_e_$tag.json 2> gpurun_out/bench_e_$tag.err
  python -c "import json;d=json.load(open('gpurun_out/bench_e_$tag.json'));print('$tag', '%.4g %.4g %.3f'%(d['value'], d['e2e']['value'], d['e2e']['ms_per_step']))"; }
run base A=1
run split EGGPHOT_WALK_SPLIT=1
run split_c32 EGGPHOT_WALK_SPLIT=1 EGGPHOT_CHUNK=32768
run split_c65 EGGPHOT_WALK_SPLIT=1 EGGPHOT_CHUNK=65536
EGGPHOT_WALK_SPLIT=1 EGGPHOT_TRACE=1 timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu --no-extra > gpurun_out/trace_split.json 2> gpurun_out/trace_split.err; tail -5 gpurun_out/trace_split.err
EXTRA="--ngal 2000000 --steps 4"
run big A=1
run big_split EGGPHOT_WALK_SPLIT=1
