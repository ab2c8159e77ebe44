mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-extra $EXTRA > gpurun_out/bench_e_$tag.json 2> gpurun_out/bench_e_$tag.err
  python -c "import json;d=json.load(open('gpurun_out/bench_e_$tag.json'));print('$tag', '%.4g %.4g %.3f'%(d['value'], d['e2e']['value'], d['e2e']['ms_per_step']))"; }
run default A=1
run notaper EGGPHOT_NO_TAPER=1
run cs2 EGGPHOT_CSTREAMS=2
run c65536 EGGPHOT_CHUNK=65536
run c40960 EGGPHOT_CHUNK=40960
run c32768 EGGPHOT_CHUNK=32768
EGGPHOT_TRACE=1 timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu --no-extra > gpurun_out/trace_default.json 2> gpurun_out/trace_default.err; tail -5 gpurun_out/trace_default.err
EXTRA="--ngal 2000000"
run big A=1
run big_cs2 EGGPHOT_CSTREAMS=2
