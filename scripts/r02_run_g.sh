#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_step_gpu.py tests/test_config2_resident_gpu.py tests/test_bench_scale_gpu.py -x -q -m gpu 2>&1 | tail -6
for v in "X=1" "MITGCM_B200_NO_PHIFUSE=1"; do
  env $v timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-dropin 2>/dev/null | python -c "
import json,sys
r=json.loads(sys.stdin.readline()); p=r['phase_ms_per_step']
print('$v', 'value %.3f step %.2f ms |' % (r['value'], r['ms_per_step']), ' '.join('%s %.2f' % (k,v) for k,v in p.items()))"
done
