set -x
mkdir -p gpurun_out
timeout 600 python tools/parity_sweep.py b30_rf3_defaults 20000 > gpurun_out/p_b30.txt 2>&1; echo "rc=$?" >> gpurun_out/p_b30.txt; grep -E "max rel|rc=" gpurun_out/p_b30.txt | cut -c1-150
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-extra > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; tail -2 gpurun_out/bench_quick.err; head -c 200 gpurun_out/bench_quick.json
for ch in 16384 32768 49152 65536 98304 195008; do
  EGGPHOT_CHUNK=$ch timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-extra > gpurun_out/bench_chunk_$ch.json 2> gpurun_out/bench_chunk_$ch.err
  python -c "import json;d=json.load(open('gpurun_out/bench_chunk_$ch.json'));print($ch, d['value'], d['e2e']['value'], d['e2e']['ms_per_step'])"
done
( time timeout 1500 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err ) 2> gpurun_out/bench_full.time; tail -3 gpurun_out/bench_full.time; tail -2 gpurun_out/bench_full.err
( time timeout 900 python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err ) 2> gpurun_out/bench_ref.time; tail -3 gpurun_out/bench_ref.time; cat gpurun_out/bench_ref.json | cut -c1-600
