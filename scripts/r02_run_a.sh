#!/bin/bash
# round-2 GPU run A: full GPU suite, CG2D size sweep (sync overhead), SR timings, bench
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/r02_pytest_a.log
cat gpurun_out/r02_pytest_a.log
for n in 256 512 1024 2048; do python scripts/cg2d_perf.py $n 200 3 0; done 2>&1 | grep "^N=" > gpurun_out/r02_cg2d_sizes.log
for n in 256 1024 2048; do python scripts/cg2d_perf.py $n 200 3 1; done 2>&1 | grep "^N=" >> gpurun_out/r02_cg2d_sizes.log
cat gpurun_out/r02_cg2d_sizes.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_a.json 2> gpurun_out/r02_bench_a.err
tail -c 3000 gpurun_out/r02_bench_a.json; tail -5 gpurun_out/r02_bench_a.err
