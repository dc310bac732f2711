import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from mitgcm_b200 import runtime as rt, distributed
from mitgcm_b200.grid import Dims, exch_xyz
from mitgcm_b200.model import LIB_PARAMS, ini_cg2d, make_channel
from mitgcm_b200.parallel import process_grid
from oracle.pyoracle import Oracle
local = int(os.environ.get("LOCAL_RANK", 0)); torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
nPx, nPy = process_grid(world)
sNx, sNy, NR = 32, 24, 3
gG, P, sG = make_channel(sNx, sNy, NR, nSx=nPx, nSy=nPy, land_frac=0.15)
opG = ini_cg2d(gG, P)
px, py = rank % nPx, rank // nPx
d = Dims(sNx=sNx, sNy=sNy, OLx=2, OLy=2, Nr=NR, nPx=nPx, nPy=nPy, myPx=px, myPy=py)
mine = lambda a: np.array(a[py:py + 1, px:px + 1], copy=True)
rt.init(d, local)
distributed.setup(d)
# 1. exchange test
rng = np.random.default_rng(5)
A3 = rng.standard_normal(gG.d.shape3); A2 = rng.standard_normal(gG.d.shape2)
rt.set_field("theta", mine(A3)); rt.set_field("etaN", mine(A2))
distributed.exchange("theta"); distributed.exchange("etaN")
R3 = exch_xyz(gG.d, A3.copy()); R2 = exch_xyz(gG.d, A2.copy())
g3 = rt.get_field("theta", np.zeros(d.shape3)); g2 = rt.get_field("etaN", np.zeros(d.shape2))
print(rank, "exch3 equal", np.array_equal(g3, mine(R3)), "exch2 equal", np.array_equal(g2, mine(R2)), flush=True)
# 2. cg2d test vs oracle on the global tiling
rt.set_cg2d_operator({k: (mine(v) if isinstance(v, np.ndarray) else v) for k, v in opG.items()})
b = np.zeros(gG.d.shape2); jj, ii = gG.d.interior()
b[:, :, jj, ii] = rng.standard_normal((nPy, nPx, sNy, sNx)); b *= gG.maskC[:, :, 0] * gG.rA / 1200.0
x = 0.1 * rng.standard_normal(gG.d.shape2) * gG.maskC[:, :, 0]
o = Oracle(gG, dict(globalArea=P["globalArea"]))
for nit in (1, 3, 500):
    bo, xo = b.copy(), x.copy()
    ro = o.cg2d(opG, bo, xo, nit, -1)
    bm, xm = mine(b), mine(x)
    rg = rt.cg2d(bm, xm, nit, -1)
    err = np.abs(xm[:, :, jj, ii] - mine(xo)[:, :, jj, ii]).max() / np.abs(xo).max()
    print(rank, "cg2d", nit, rg["numIters"], ro["numIters"], rg["firstResidual"], ro["firstResidual"], rg["lastResidual"], ro["lastResidual"], "xerr", err, flush=True)
rt.finalize(); dist.barrier(); dist.destroy_process_group()
