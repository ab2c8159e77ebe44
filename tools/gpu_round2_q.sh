mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-extra $EXTRA > gpurun_out/bench_e_$tag.json 2> gpurun_out/bench_e_$tag.err
  python -c "import json;d=json.load(open('gpurun_out/bench_e_$tag.json'));print('$tag', '%.4g %.4g %.3f'%(d['value'], d['e2e']['value'], d['e2e']['ms_per_step']))"; }
timeout 600 python tools/parity_sweep.py b30_rf3_defaults 20000 > gpurun_out/p_b30.txt 2>&1; echo "rc=$?" >> gpurun_out/p_b30.txt; grep -E "max rel|rc=" gpurun_out/p_b30.txt | cut -c1-150
run w_default A=1
run w_cs2 EGGPHOT_CSTREAMS=2
run w_c65536 EGGPHOT_CHUNK=65536
run w_c65536_cs2 EGGPHOT_CHUNK=65536 EGGPHOT_CSTREAMS=2
EXTRA="--ngal 24384"
run w_small A=1
EXTRA="--ngal 2000000 --steps 4"
run w_big A=1
