#!/bin/bash
mkdir -p gpurun_out
for st in 2 4; do
MITGCM_B200_DYN_TMA_STAGES=$st timeout 600 ncu --set full --import-source on --clock-control none -k regex:dyn_tma_uv --launch-skip 2 --launch-count 1 \
  -o gpurun_out/r02_dyn_uv_st$st -f python bench.py --nx 1024 --ny 1024 --steps 1 --warmup 3 --no-cpu-baseline --no-dropin > gpurun_out/ncu_e_$st.log 2>&1
tail -3 gpurun_out/ncu_e_$st.log
done
ls -la gpurun_out/r02_dyn_uv_st*.ncu-rep
