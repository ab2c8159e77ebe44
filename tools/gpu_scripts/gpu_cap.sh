mkdir -p gpurun_out
for mb in 4096 8192 16384 32768; do
  EGGPHOT_LANE_SCRATCH_MB=$mb timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-extra --no-e2e --precision fp64 > gpurun_out/cap_fp64_$mb.json 2>gpurun_out/cap.err
  EGGPHOT_LANE_SCRATCH_MB=$mb timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu --no-extra --ngal 2000000 > gpurun_out/cap_2M_$mb.json 2>gpurun_out/cap.err
  EGGPHOT_LANE_SCRATCH_MB=$mb timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu --no-extra --workload cfg4 > gpurun_out/cap_cfg4_$mb.json 2>gpurun_out/cap.err
  python -c "
import json
a=json.load(open('gpurun_out/cap_fp64_$mb.json')); b=json.load(open('gpurun_out/cap_2M_$mb.json')); c=json.load(open('gpurun_out/cap_cfg4_$mb.json'))
print($mb, 'fp64 %.4g | 2M %.4g e2e %.4g | cfg4 %.4g e2e %.4g'%(a['value'], b['value'], b['e2e']['value'], c['value'], c['e2e']['value']))"
done
