mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --no-extra $EXTRA > gpurun_out/bench_e_$tag.json 2> gpurun_out/bench_e_$tag.err
  python -c "import json;d=json.load(open('gpurun_out/bench_e_$tag.json'));print('$tag', '%.4g %.4g %.3f'%(d['value'], d['e2e']['value'], d['e2e']['ms_per_step']))"; }
for rep in 1 2; do
run base$rep A=1
run c65k_notaper$rep EGGPHOT_CHUNK=65024 EGGPHOT_NO_TAPER=1
run c65k_taper$rep EGGPHOT_CHUNK=65024
run c98k_taper$rep EGGPHOT_CHUNK=97504
run c49k_notaper$rep EGGPHOT_CHUNK=48768 EGGPHOT_NO_TAPER=1
done
