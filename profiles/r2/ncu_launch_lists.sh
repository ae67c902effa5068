mkdir -p gpurun_out/ncu
B0="python bench.py --no-cpu-baseline --e2e-steps 1"
for w in cfg3 cfg2 cfg5; do
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/ncu/launches_$w.csv $B0 --workload $w --steps 1 --warmup 3 > gpurun_out/ncu/launches_$w.log 2>&1
  tail -1 gpurun_out/ncu/launches_$w.log | cut -c1-150
done
ls -la gpurun_out/ncu
