#!/bin/bash
# prints per-phase ms of the bench step for the given environment settings: scripts/phase_times.sh [VAR=val ...]
for cfg in "$@"; do
  env $cfg python bench.py --steps 4 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
r=json.loads(sys.stdin.readline()); p=r['phase_ms_per_step']
print('$cfg', 'step %.2f ms |' % r['ms_per_step'], ' '.join('%s %.2f' % (k,v) for k,v in p.items()))"
done
