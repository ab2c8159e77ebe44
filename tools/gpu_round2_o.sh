mkdir -p gpurun_out
for n in 24384 48768 97536; do
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 10 --csv --log-file gpurun_out/launches_$n.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-extra --ngal $n > gpurun_out/ncu_launches_$n.log 2>&1
echo "== $n"; tail -10 gpurun_out/launches_$n.csv | cut -d\" -f10,30 | cut -c1-40,100-200
done
