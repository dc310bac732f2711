#!/bin/bash
# round-2 GPU run C (N GPUs): multi-rank parity + weak bench with the flag barrier and with grid.sync
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
O=gpurun_out; mkdir -p $O
for cs in 0 1; do
  export MITGCM_B200_CG2D_COOPSYNC=$cs
  timeout 300 $TR scripts/dist_check.py 64 48 4 45 > $O/r02c_dist${N}_cs$cs.log 2>&1; echo "dist cs=$cs rc=$?"; grep -a "DIST_CHECK" $O/r02c_dist${N}_cs$cs.log | cut -c1-160
  timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 5 > $O/r02c_bench${N}_cs$cs.out 2> $O/r02c_bench${N}_cs$cs.err; echo "bench cs=$cs rc=$?"
done
unset MITGCM_B200_CG2D_COOPSYNC
python - <<PY
import json
for cs in (0, 1):
    try:
        j = json.loads(open("$O/r02c_bench${N}_cs%d.out" % cs).read().strip().splitlines()[-1])
        print("coopsync", cs, "value", round(j["value"], 3), "ms/step", round(j["ms_per_step"], 3), "e2e", round(j["e2e"]["value"], 3),
              "iters", j["cg2d"]["iters_per_step"], "us/iter", round(j["cg2d"]["us_per_iter"], 2), j["phase_ms_per_step"], j.get("multi_rank_check", {}).get("ok"))
    except Exception as e:
        print(cs, "no line:", e)
PY
