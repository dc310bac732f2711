#!/bin/bash
# scripts/r02_run_n.sh N: N-rank parity (45 steps) + the driver's weak bench command + strong scaling, logs in gpurun_out/
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
O=gpurun_out; mkdir -p $O
timeout 300 $TR scripts/dist_check.py 64 48 4 45 > $O/r02f_dist${N}.log 2>&1; echo "dist rc=$?"; grep -a "DIST_CHECK" $O/r02f_dist${N}.log | cut -c1-160
timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 5 > $O/r02f_bench${N}_weak.out 2> $O/r02f_bench${N}_weak.err; echo "bench weak rc=$?"
timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 5 --scaling strong --no-selfcheck > $O/r02f_bench${N}_strong.out 2> $O/r02f_bench${N}_strong.err; echo "bench strong rc=$?"
python - <<PY
import json
for kind in ("weak", "strong"):
    try:
        j = json.loads(open("$O/r02f_bench${N}_%s.out" % kind).read().strip().splitlines()[-1])
        print(kind, "value", round(j["value"], 3), j["unit"], "ms/step", round(j["ms_per_step"], 3), "e2e", round(j["e2e"]["value"], 3),
              "iters", j["cg2d"]["iters_per_step"], "us/iter", round(j["cg2d"]["us_per_iter"], 2), j["health"]["after_timed"]["finite"], j["phase_ms_per_step"], j.get("multi_rank_check", {}).get("ok"))
    except Exception as e:
        print(kind, "no line:", e)
PY
