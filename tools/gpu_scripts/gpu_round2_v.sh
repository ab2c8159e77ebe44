mkdir -p gpurun_out/r2b
O=gpurun_out/r2b
nvidia-smi topo -m > $O/topo.txt 2>&1
timeout 600 python -m pytest tests/test_cli.py -m gpu -q -k "two_gpu" > $O/pytest_2gpu_cli.txt 2>&1; tail -3 $O/pytest_2gpu_cli.txt
for n in 8 4 2; do
  extra="--no-extra --no-cpu"; [ $n = 8 ] && extra=""
  ( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 10 --warmup 3 $extra > $O/bench_n$n.json 2> $O/bench_n$n.err ) 2> $O/bench_n$n.time
  tail -1 $O/bench_n$n.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print($n, '%.4g'%d['value'], '%.4g'%d['e2e']['value'], (d.get('configs3') or {}).get('value'), ((d.get('configs3') or {}).get('e2e') or {}).get('value'), (d.get('cpu_baseline') or {}).get('cores'))"
  grep real $O/bench_n$n.time
done
