# cfg-4 (reduced) launch list + full capture; bench.py must see >= 4 draws for the diagnostics, hence --steps 2
mkdir -p gpurun_out/ncu
B="python bench.py --no-cpu-baseline --e2e-steps 1 --workload cfg4r"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/ncu/launches_cfg4r.csv $B --steps 4 --warmup 3 > gpurun_out/ncu/launches_cfg4r.log 2>&1
timeout 1200 ncu --set full --import-source on --clock-control none -k regex:rsf_mcmc_kernel -s 1 -c 1 -f -o gpurun_out/ncu/r2_cfg4r $B --steps 4 --warmup 3 > gpurun_out/ncu/full_cfg4r.log 2>&1
f=r2_cfg4r
ncu -i gpurun_out/ncu/$f.ncu-rep --page raw --csv > gpurun_out/ncu/${f}_raw.csv 2>/dev/null
ncu -i gpurun_out/ncu/$f.ncu-rep --page details --csv > gpurun_out/ncu/${f}_details.csv 2>/dev/null
ncu -i gpurun_out/ncu/$f.ncu-rep --page source --csv --print-source sass > gpurun_out/ncu/${f}_source_sass.csv 2>/dev/null
gzip -f gpurun_out/ncu/${f}_source_sass.csv; rm -f gpurun_out/ncu/$f.ncu-rep
tail -3 gpurun_out/ncu/full_cfg4r.log; ls -la gpurun_out/ncu | grep cfg4r
