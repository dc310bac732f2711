"""Where does the N-rank weak-scaled step first differ from the one-block step?  (debug aid of weak_equiv_check.py)"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import bench
from mitgcm_b200 import distributed, runtime as rt
ap = argparse.ArgumentParser()
ap.add_argument("--nx", type=int, default=2048); ap.add_argument("--ny", type=int, default=2048); ap.add_argument("--nr", type=int, default=50)
a = ap.parse_args()
args = argparse.Namespace(nx=a.nx, ny=a.ny, nr=a.nr, scaling="weak", momentum="fluxform", temp_adv_scheme=2)
local = int(os.environ.get("LOCAL_RANK", 0)); torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size(); dev = torch.device("cuda", local)
NAMES2 = ("etaN", "cg2d_b", "cg2d_x"); NAMES3 = ("gU", "gV", "uVel", "vVel", "theta", "phiHyd", "wVel")
def snap(W):
    out = {}
    for n in NAMES2 + NAMES3:
        b = torch.empty((W.d.PY, W.d.PX) if n in NAMES2 else (W.NR, W.d.PY, W.d.PX), device=dev, dtype=torch.float64)
        rt.get_field(n, b); out[n] = b[..., :, :].cpu() if n in NAMES2 else b[:2].cpu()      # top two levels of 3-D fields
    return out
def run(multi):
    W = bench.setup_workload(args, rank if multi else 0, world if multi else 1, local)
    s0 = snap(W)
    r = (distributed.forward_step if multi else rt.forward_step)(0)
    s1 = snap(W)
    if multi: distributed.teardown()
    rt.finalize()
    return s0, s1, r, W
a0, a1, ra, W = run(False); dist.barrier(); b0, b1, rb, W = run(True)
if rank == 0:
    print("solver", ra, rb)
    for tag, x, y in (("after setup", a0, b0), ("after step 0", a1, b1)):
        for n in x:
            d = (x[n] - y[n]).abs(); m = float(d.max()); sc = float(x[n].abs().max())
            idx = np.unravel_index(int(d.argmax()), tuple(d.shape))
            OL = 2
            inter = d[..., OL:-OL, OL:-OL]; mi = float(inter.max())
            print(f"{tag:13s} {n:8s} max|diff| {m:.3e} (interior {mi:.3e}) of {sc:.3e} at {tuple(int(v) for v in idx)}")
dist.destroy_process_group()
