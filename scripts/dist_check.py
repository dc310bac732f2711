"""Multi-GPU parity check, run under torchrun: the same global problem is stepped (a) on N ranks and (b) on rank 0
alone as a single process holding all tiles; results must agree (mitgcm_b200.distributed.selfcheck).
usage: torchrun --nproc-per-node N scripts/dist_check.py [NXg NYg NR steps]
env: DIST_CHECK_WIDE=1 (MOM_VECINV + salt + DST3 at overlap 3), DIST_CHECK_BUOYANCY=0, MITGCM_B200_TRANSPORT=nccl"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from mitgcm_b200 import distributed

local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
NXg, NYg, NR, nsteps = (int(a) for a in (sys.argv[1:5] + ["64", "48", "4", "3"][len(sys.argv) - 1:]))
r = distributed.selfcheck(NXg, NYg, NR, nsteps, wide=bool(int(os.environ.get("DIST_CHECK_WIDE", "0"))),
                          buoyancy=bool(int(os.environ.get("DIST_CHECK_BUOYANCY", "1"))), verbose=True,
                          land_frac=float(os.environ.get("DIST_CHECK_LAND", "0.15")))
if dist.get_rank() == 0:
    print("DIST_CHECK", "PASS" if r["ok"] else "FAIL", json.dumps(r))
dist.destroy_process_group()
sys.exit(0 if r["ok"] else 1)
