#!/bin/bash
# Multi-GPU checks and benches on N GPUs of one box: scripts/multi_gpu_suite.sh N [tag]
# Logs go to gpurun_out/ (copy what should be judged into profiles/).
N=${1:-2}; TAG=${2:-r02}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
O=gpurun_out
# N-rank vs 1-rank parity of the resident step, 45 steps (peer transport, then the wide option set, then NCCL)
timeout 300 $TR scripts/dist_check.py 64 48 4 45 > $O/${TAG}_dist${N}_peer.log 2>&1; echo "dist peer rc=$?"; grep -a "DIST_CHECK" $O/${TAG}_dist${N}_peer.log | cut -c1-400
DIST_CHECK_WIDE=1 timeout 300 $TR scripts/dist_check.py 96 72 5 12 > $O/${TAG}_dist${N}_wide.log 2>&1; echo "dist wide rc=$?"; grep -a "DIST_CHECK" $O/${TAG}_dist${N}_wide.log | cut -c1-300
MITGCM_B200_TRANSPORT=nccl timeout 300 $TR scripts/dist_check.py 64 48 4 45 > $O/${TAG}_dist${N}_nccl.log 2>&1; echo "dist nccl rc=$?"; grep -a "DIST_CHECK" $O/${TAG}_dist${N}_nccl.log | cut -c1-300
# full-scale equivalence of the weak-scaled workload with the one-block run (every rank steps both)
timeout 400 $TR scripts/weak_equiv_check.py --steps 8 > $O/${TAG}_weak_equiv${N}.log 2>&1; echo "weak equiv rc=$?"; grep -a "WEAK_EQUIV\|final" $O/${TAG}_weak_equiv${N}.log | cut -c1-300
# weak-scaled bench (the driver's command), strong scaling of one 2048^2 x 50 domain, and the NCCL transport for comparison
timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 5 > $O/${TAG}_bench${N}_weak.out 2> $O/${TAG}_bench${N}_weak.err; echo "bench weak rc=$?"; tail -c 300 $O/${TAG}_bench${N}_weak.err
timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 5 --scaling strong --no-selfcheck > $O/${TAG}_bench${N}_strong.out 2> $O/${TAG}_bench${N}_strong.err; echo "bench strong rc=$?"
MITGCM_B200_TRANSPORT=nccl timeout 600 $TR bench.py --gpus $N --steps 10 --warmup 5 --no-selfcheck > $O/${TAG}_bench${N}_nccl.out 2> $O/${TAG}_bench${N}_nccl.err; echo "bench nccl rc=$?"; tail -c 600 $O/${TAG}_bench${N}_nccl.err | grep -a error
python - <<PY
import json
for kind in ("weak", "strong", "nccl"):
    try:
        j = json.loads(open("$O/${TAG}_bench${N}_%s.out" % kind).read().strip().splitlines()[-1])
        print(kind, "value", round(j["value"], 3), j["unit"], "ms/step", round(j["ms_per_step"], 3), "e2e", round(j["e2e"]["value"], 3),
              "iters", j["cg2d"]["iters_per_step"], "us/iter", round(j["cg2d"]["us_per_iter"], 2), j["health"]["after_timed"], j["phase_ms_per_step"])
    except Exception as e:
        print(kind, "no line:", e)
PY
