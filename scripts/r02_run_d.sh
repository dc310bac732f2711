#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_step_gpu.py -x -q -m gpu -k "dyn_tma" 2>&1 | tail -4
for v in "MITGCM_B200_DYN_TMA_STAGES=2" "MITGCM_B200_DYN_TMA_STAGES=3" "MITGCM_B200_DYN_TMA_STAGES=4" "MITGCM_B200_DYN_TMA_NOSPLIT=1"; do
  env $v timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-dropin > gpurun_out/r02_bench_d.json 2> gpurun_out/r02_bench_d.err
  python - <<PY
import json
j = json.loads(open("gpurun_out/r02_bench_d.json").read().strip().splitlines()[-1])
print("$v", j["value"], j["ms_per_step"], j["phase_ms_per_step"])
PY
  tail -2 gpurun_out/r02_bench_d.err
done
