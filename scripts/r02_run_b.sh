#!/bin/bash
# round-2 GPU run B: flag barrier vs cooperative grid.sync in CG2D / CG2D_SR
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_cg2d_gpu.py tests/test_bench_scale_gpu.py tests/test_step_gpu.py -x -q -m gpu 2>&1 | tail -6 > gpurun_out/r02_pytest_b.log
cat gpurun_out/r02_pytest_b.log
L=gpurun_out/r02_cg2d_flagbarrier.log; : > $L
for n in 256 1024 2048; do
  for sr in 0 1; do
    echo "== flag barrier N=$n sr=$sr" >> $L; timeout 120 python scripts/cg2d_perf.py $n 200 3 $sr 2>&1 | grep "^N=" | tail -1 >> $L
    echo "== grid.sync    N=$n sr=$sr" >> $L; MITGCM_B200_CG2D_COOPSYNC=1 timeout 120 python scripts/cg2d_perf.py $n 200 3 $sr 2>&1 | grep "^N=" | tail -1 >> $L
  done
done
cat $L
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_b.json 2> gpurun_out/r02_bench_b.err
python - <<PY
import json
j = json.loads(open("gpurun_out/r02_bench_b.json").read().strip().splitlines()[-1])
print(j["value"], j["ms_per_step"], j["cg2d"], j["phase_ms_per_step"], j["e2e"]["value"])
PY
tail -3 gpurun_out/r02_bench_b.err
