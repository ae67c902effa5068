# ncu evidence of the predictive speculative kernel (cfg 2), after bench.py --workload cfg2 ran clean without ncu
set -x
mkdir -p gpurun_out/ncu
B0="python bench.py --no-cpu-baseline --e2e-steps 1 --no-sub-records"
timeout 300 $B0 --workload cfg2 --steps 1 --warmup 3 > gpurun_out/ncu/plain_cfg2.json 2> gpurun_out/ncu/plain_cfg2.err || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/ncu/launches_cfg2.csv $B0 --workload cfg2 --steps 1 --warmup 3 > gpurun_out/ncu/launches_cfg2.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:rsf_mcmc_spec_kernel -s 2 -c 1 -f -o gpurun_out/ncu/r2_cfg2_spec $B0 --workload cfg2 --steps 1 --warmup 3 > gpurun_out/ncu/full_cfg2.log 2>&1
ncu -i gpurun_out/ncu/r2_cfg2_spec.ncu-rep --page raw --csv > gpurun_out/ncu/r2_cfg2_spec_raw.csv 2>/dev/null
ncu -i gpurun_out/ncu/r2_cfg2_spec.ncu-rep --page details --csv > gpurun_out/ncu/r2_cfg2_spec_details.csv 2>/dev/null
ncu -i gpurun_out/ncu/r2_cfg2_spec.ncu-rep --page source --csv --print-source sass > gpurun_out/ncu/r2_cfg2_spec_source_sass.csv 2>/dev/null
gzip -f gpurun_out/ncu/r2_cfg2_spec_source_sass.csv
rm -f gpurun_out/ncu/r2_cfg2_spec.ncu-rep
ls -la gpurun_out/ncu
