# cfg 4 reduced under ncu: launch list of the bench workload and one full capture of its MCMC kernel (two lanes per chain)
mkdir -p gpurun_out/ncu
B="python bench.py --no-cpu-baseline --e2e-steps 1 --no-sub-records --workload cfg4r"
timeout 300 $B --steps 1 --warmup 3 > gpurun_out/ncu/plain_cfg4r.json 2> gpurun_out/ncu/plain_cfg4r.err || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/ncu/launches_cfg4r.csv $B --steps 1 --warmup 3 > gpurun_out/ncu/launches_cfg4r.log 2>&1
timeout 1200 ncu --set full --import-source on --clock-control none -k regex:rsf_mcmc_spec_kernel -s 1 -c 1 -f -o gpurun_out/ncu/r2_cfg4r $B --steps 1 --warmup 3 > gpurun_out/ncu/full_cfg4r.log 2>&1
f=r2_cfg4r
ncu -i gpurun_out/ncu/$f.ncu-rep --page raw --csv > gpurun_out/ncu/${f}_raw.csv 2>/dev/null
ncu -i gpurun_out/ncu/$f.ncu-rep --page details --csv > gpurun_out/ncu/${f}_details.csv 2>/dev/null
ncu -i gpurun_out/ncu/$f.ncu-rep --page source --csv --print-source sass > gpurun_out/ncu/${f}_source_sass.csv 2>/dev/null
gzip -f gpurun_out/ncu/${f}_source_sass.csv; rm -f gpurun_out/ncu/$f.ncu-rep
tail -3 gpurun_out/ncu/full_cfg4r.log; ls -la gpurun_out/ncu | grep cfg4r
