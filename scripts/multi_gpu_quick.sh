#!/bin/bash
# scripts/multi_gpu_quick.sh N [tag]: the two runs that matter most -- full-scale equivalence + the driver's bench command
N=${1:-8}; TAG=${2:-r02}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
O=gpurun_out
timeout 300 $TR scripts/dist_check.py 64 48 4 45 > $O/${TAG}_dist${N}_peer.log 2>&1; echo "dist peer rc=$?"; grep -a "DIST_CHECK" $O/${TAG}_dist${N}_peer.log | cut -c1-200
timeout 400 $TR scripts/weak_equiv_check.py --steps 6 > $O/${TAG}_weak_equiv${N}.log 2>&1; echo "weak equiv rc=$?"; grep -a "WEAK_EQUIV\|final\|step\": 5" $O/${TAG}_weak_equiv${N}.log | cut -c1-300
timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 5 > $O/${TAG}_bench${N}_weak.out 2> $O/${TAG}_bench${N}_weak.err; echo "bench weak rc=$?"
timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 5 --scaling strong --no-selfcheck > $O/${TAG}_bench${N}_strong.out 2> $O/${TAG}_bench${N}_strong.err; echo "bench strong rc=$?"
python - <<PY
import json
for kind in ("weak", "strong"):
    try:
        j = json.loads(open("$O/${TAG}_bench${N}_%s.out" % kind).read().strip().splitlines()[-1])
        print(kind, "value", round(j["value"], 3), j["unit"], "ms/step", round(j["ms_per_step"], 3), "e2e", round(j["e2e"]["value"], 3),
              "iters", j["cg2d"]["iters_per_step"], "us/iter", round(j["cg2d"]["us_per_iter"], 2), j["health"]["after_timed"], j["phase_ms_per_step"], j.get("multi_rank_check", {}).get("ok"))
    except Exception as e:
        print(kind, "no line:", e)
PY
