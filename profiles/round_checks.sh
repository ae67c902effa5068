set -x
python -m pytest tests -m gpu -x -q > gpurun_out/t_all.txt 2>&1
python bench.py > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err
python bench.py --chains 131072 --iters 20 --no-cpu-baseline > gpurun_out/bench_c131072.json 2> gpurun_out/bench_c131072.err
python bench.py --impl reference > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
python bench.py --no-cpu-baseline > gpurun_out/bench_plain.json 2>/dev/null && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
export STIFF_N=1000
python profiles/microbench/forward_stiff_once.py > gpurun_out/stiff_once.txt 2>&1 && \
ncu --section SourceCounters --section WarpStateStats --section SchedulerStats --section LaunchStats --section InstructionStats --section Occupancy --clock-control none --import-source on -k regex:rsf_forward -s 1 -c 1 -f -o gpurun_out/prof_stiff7 python profiles/microbench/forward_stiff_once.py > gpurun_out/ncu_stiff7.log 2>&1
python profiles/microbench/run_configs.py cfg4 2 > gpurun_out/cfg4.json 2> gpurun_out/cfg4.err
python profiles/microbench/forward_stiff_ab.py > gpurun_out/stiff_ab.txt 2>&1
