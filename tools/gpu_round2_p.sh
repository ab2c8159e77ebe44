mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-extra $EXTRA > gpurun_out/bench_e_$tag.json 2> gpurun_out/bench_e_$tag.err
  python -c "import json;d=json.load(open('gpurun_out/bench_e_$tag.json'));print('$tag', '%.4g %.4g %.3f'%(d['value'], d['e2e']['value'], d['e2e']['ms_per_step']))"; }
run im0 EGGPHOT_ITEM_MODE=0
run im1 EGGPHOT_ITEM_MODE=1
run im2 EGGPHOT_ITEM_MODE=2
EXTRA="--ngal 24384"
run s_im0 EGGPHOT_ITEM_MODE=0
run s_im1 EGGPHOT_ITEM_MODE=1
run s_im2 EGGPHOT_ITEM_MODE=2
EXTRA="--ngal 2000000 --steps 4"
run b_im0 EGGPHOT_ITEM_MODE=0
run b_im1 EGGPHOT_ITEM_MODE=1
run b_im2 EGGPHOT_ITEM_MODE=2
