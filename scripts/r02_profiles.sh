#!/bin/bash
# round-2 evidence run: GPU suite, bench line, ncu launch list of the same command, ncu --set full of the step kernels
mkdir -p gpurun_out
O=gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -4 > $O/r02_pytest_final.log; cat $O/r02_pytest_final.log
python bench.py --steps 10 --warmup 3 > $O/r02_bench_line.json 2> $O/r02_bench_line.err; tail -c 600 $O/r02_bench_line.json; tail -2 $O/r02_bench_line.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-dropin > $O/r02_ncu_launch.log 2>&1
python scripts/launch_list.py $O/r02_launches.csv > $O/r02_bench_launch_list.csv; cat $O/r02_bench_launch_list.csv
# full capture of one launch of each step kernel at the bench size (4th step)
ncu --set full --import-source on --clock-control none -k regex:'cg2d_kernel|dyn_tma_uv|thermo_pipe|corr_kernel|rhs_kernel|phihyd' --launch-skip 18 --launch-count 6 \
  -o $O/r02_step_kernels_2048 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-dropin > $O/r02_ncu_full.log 2>&1
tail -2 $O/r02_ncu_full.log
python scripts/ncu_summary.py $O/r02_step_kernels_2048.ncu-rep > $O/r02_step_kernels_2048x2048x50.txt; grep "==\|time_duration\|dram__bytes\|dram_throughput" $O/r02_step_kernels_2048x2048x50.txt
