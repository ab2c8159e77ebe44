set -x
mkdir -p gpurun_out
EGGPHOT_TRACE=1 timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu --no-extra > gpurun_out/trace_default.json 2> gpurun_out/trace_default.err; tail -4 gpurun_out/trace_default.err
for ch in 0 24576 32768 65536; do
  EGGPHOT_CHUNK=$ch timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-extra > gpurun_out/bench_chunk_$ch.json 2> gpurun_out/bench_chunk_$ch.err
  python -c "import json;d=json.load(open('gpurun_out/bench_chunk_$ch.json'));print($ch, '%.4g %.4g %.3f'%(d['value'], d['e2e']['value'], d['e2e']['ms_per_step']))"
done
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1; tail -5 gpurun_out/pytest_gpu.txt
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_run.py 2048 > gpurun_out/sanitizer_memcheck.txt 2>&1; tail -4 gpurun_out/sanitizer_memcheck.txt
timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_run.py 1024 > gpurun_out/sanitizer_racecheck.txt 2>&1; tail -4 gpurun_out/sanitizer_racecheck.txt
