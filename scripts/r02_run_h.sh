#!/bin/bash
for lib in "" build/variants/lib_A.so build/variants/lib_B.so build/variants/lib_C.so; do
  for v in "X=1" "MITGCM_B200_NO_COLGEOM=1"; do
    if [ -n "$lib" ] && [ "$v" != "X=1" ]; then continue; fi
    env MITGCM_B200_LIB=$([ -n "$lib" ] && echo $PWD/$lib) $v timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-dropin 2>/dev/null | python -c "
import json,sys
r=json.loads(sys.stdin.readline()); p=r['phase_ms_per_step']
print('$lib $v', 'step %.2f ms |' % r['ms_per_step'], ' '.join('%s %.2f' % (k,v) for k,v in p.items()))"
  done
done
