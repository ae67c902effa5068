# ncu evidence of round 2 (run on the GPU box through gpurun; every command only after bench.py ran clean without ncu)
set -x
mkdir -p gpurun_out/ncu
B0="python bench.py --no-cpu-baseline --e2e-steps 1"                    # launch lists: the bench as it runs (chain groups on)
B="$B0 --chain-groups 1"       # full captures: one launch per adaptation interval (ncu serialises kernels anyway)
export_rep() {   # the .ncu-rep with imported source is 25 MB: export what is read, drop the report
  ncu -i gpurun_out/ncu/$1.ncu-rep --page raw --csv > gpurun_out/ncu/$1_raw.csv 2>/dev/null
  ncu -i gpurun_out/ncu/$1.ncu-rep --page details --csv > gpurun_out/ncu/$1_details.csv 2>/dev/null
  ncu -i gpurun_out/ncu/$1.ncu-rep --page source --csv --print-source cuda > gpurun_out/ncu/$1_source_cuda.csv 2>/dev/null
  ncu -i gpurun_out/ncu/$1.ncu-rep --page source --csv --print-source sass > gpurun_out/ncu/$1_source_sass.csv 2>/dev/null
  gzip -f gpurun_out/ncu/$1_source_sass.csv gpurun_out/ncu/$1_source_cuda.csv
  rm -f gpurun_out/ncu/$1.ncu-rep
}
if [ "$1" != "full-only" ]; then
for w in cfg3 cfg2 cfg5 cfg4r; do
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/ncu/launches_$w.csv $B0 --workload $w --steps 1 --warmup 3 > gpurun_out/ncu/launches_$w.log 2>&1
done
fi
timeout 600 ncu --set full --import-source on --clock-control none -k regex:rsf_mcmc_spec_kernel -s 2 -c 1 -f -o gpurun_out/ncu/r2_cfg2_spec $B --workload cfg2 --steps 1 --warmup 3 > gpurun_out/ncu/full_cfg2.log 2>&1
export_rep r2_cfg2_spec
timeout 600 ncu --set full --import-source on --clock-control none -k regex:rsf_mcmc_kernel -s 40 -c 1 -f -o gpurun_out/ncu/r2_cfg3 $B --workload cfg3 --steps 1 --warmup 3 > gpurun_out/ncu/full_cfg3.log 2>&1
export_rep r2_cfg3
timeout 600 ncu --set full --import-source on --clock-control none -k regex:rsf_mcmc_kernel -s 25 -c 1 -f -o gpurun_out/ncu/r2_cfg5 $B --workload cfg5 --steps 1 --warmup 3 > gpurun_out/ncu/full_cfg5.log 2>&1
export_rep r2_cfg5
timeout 900 ncu --set full --import-source on --clock-control none -k regex:rsf_mcmc_kernel -s 1 -c 1 -f -o gpurun_out/ncu/r2_cfg4r $B --workload cfg4r --iters 1 --steps 1 --warmup 3 > gpurun_out/ncu/full_cfg4r.log 2>&1
export_rep r2_cfg4r
ls -la gpurun_out/ncu
cat gpurun_out/ncu/full_cfg4r.log | tail -5
du -sh gpurun_out
