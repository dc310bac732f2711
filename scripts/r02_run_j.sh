#!/bin/bash
timeout 900 python -m pytest tests/test_cg2d_gpu.py -x -q -m gpu 2>&1 | tail -3
for v in "X=1" "MITGCM_B200_CG2D_NODEFERX=1"; do
  env $v timeout 120 python scripts/cg2d_perf.py 2048 200 3 0 2>&1 | grep "^N=" | tail -1 | cut -c1-150
done
