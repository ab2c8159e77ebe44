# usage: gpurun -- bash tools/gpu_scripts/gpu_variant.sh "" _name1 _name2 ...   (libraries built by tools/build_variant.sh)
mkdir -p gpurun_out
for v in "$@"; do
  lib=egg_b200/libeggphot$v.so
  EGGPHOT_LIB=$PWD/$lib timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-extra --no-e2e > gpurun_out/bench_bv$v.json 2> gpurun_out/bench_bv$v.err
  python -c "import json;d=json.load(open('gpurun_out/bench_bv$v.json'));print('variant[$v]', '%.4g'%d['value'], 'step %.3f walk %.3f groups %d'%(d['ms_per_step'], d['roofline']['kernel_ms'], d['config']['ngroups']))"
done
