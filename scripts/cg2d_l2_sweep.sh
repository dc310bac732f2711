#!/bin/bash
# A/B of the CG2D L2 eviction-priority classes (MITGCM_B200_CG2D_L2 bit mask) and the persisting window
out=gpurun_out/r02_cg2d_l2_sweep.log
: > $out
for m in 0 1 3 16 17 19 23 27 31 7; do
  echo "== MITGCM_B200_CG2D_L2=$m" >> $out
  MITGCM_B200_CG2D_L2=$m python scripts/cg2d_perf.py 2048 200 3 >> $out 2>&1
done
for wv in 100 60; do
  echo "== MITGCM_B200_CG2D_L2WIN=$wv" >> $out
  MITGCM_B200_CG2D_L2WIN=$wv python scripts/cg2d_perf.py 2048 200 3 >> $out 2>&1
  echo "== MITGCM_B200_CG2D_L2WIN=$wv L2=16" >> $out
  MITGCM_B200_CG2D_L2WIN=$wv MITGCM_B200_CG2D_L2=16 python scripts/cg2d_perf.py 2048 200 3 >> $out 2>&1
done
echo "== NORESIDENT L2=19" >> $out
MITGCM_B200_CG2D_NORESIDENT=1 MITGCM_B200_CG2D_L2=19 python scripts/cg2d_perf.py 2048 200 3 >> $out 2>&1
echo "== 4096 L2=0" >> $out
python scripts/cg2d_perf.py 4096 100 2 >> $out 2>&1
echo "== 4096 L2=17" >> $out
MITGCM_B200_CG2D_L2=17 python scripts/cg2d_perf.py 4096 100 2 >> $out 2>&1
cat $out | grep -v "^N=.*\(time\).*" | head -5
grep "^==\|^N=\|persisting" $out
