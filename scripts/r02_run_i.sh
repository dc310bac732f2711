#!/bin/bash
L=gpurun_out/r02_cg2d_512.log; : > $L
for lib in "" $PWD/build/variants/lib_cg512.so; do
 for cs in 0 1; do
  for n in 256 1024 2048; do
    echo "== lib=$(basename "$lib") coopsync=$cs N=$n" >> $L
    MITGCM_B200_LIB=$lib MITGCM_B200_CG2D_COOPSYNC=$cs timeout 120 python scripts/cg2d_perf.py $n 200 3 0 2>&1 | grep "^N=" | tail -1 >> $L
  done
 done
done
echo "== SR 512 flags 2048" >> $L
MITGCM_B200_LIB=$PWD/build/variants/lib_cg512.so timeout 120 python scripts/cg2d_perf.py 2048 200 3 1 2>&1 | grep "^N=" | tail -1 >> $L
cat $L | cut -c1-140
MITGCM_B200_LIB=$PWD/build/variants/lib_cg512.so timeout 300 python -m pytest tests/test_cg2d_gpu.py -x -q -m gpu 2>&1 | tail -3
