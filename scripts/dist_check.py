"""Multi-GPU parity check, run under torchrun: the same global problem is stepped (a) on N ranks
and (b) on rank 0 alone as a single tile; results must agree (CG2D sums are rank-ordered, so only
the summation shape differs).  usage: torchrun --nproc-per-node N scripts/dist_check.py [NXg NYg NR steps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from mitgcm_b200 import runtime as rt, distributed
from mitgcm_b200.grid import Dims, cartesian_grid, exch_xyz
from mitgcm_b200.model import DEFAULTS, LIB_PARAMS, ini_cg2d, make_channel, Model
from mitgcm_b200.parallel import process_grid

local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
NXg, NYg, NR, nsteps = (int(a) for a in (sys.argv[1:5] + ["64", "48", "4", "3"][len(sys.argv) - 1:]))
nPx, nPy = process_grid(world)
sNx, sNy = NXg // nPx, NYg // nPy
# the single-process reference set-up of the same global domain, tiled nPx x nPy
BUOY = int(os.environ.get("DIST_CHECK_BUOYANCY", "1"))      # the bench workload couples theta to the flow
# DIST_CHECK_WIDE=1: the wider option set across ranks -- MOM_VECINV, SALT_INTEGRATE, both tracers through
# GAD_ADVECTION (DST3, multi-dimensional), overlap 3
WIDE = int(os.environ.get("DIST_CHECK_WIDE", "0"))
OL = 3 if WIDE else 2
extra = dict(vectorInvariantMomentum=1, saltStepping=1, tempAdvScheme=33, saltAdvScheme=33, diffKhS=5e2, diffKrS=2e-5,
             sBeta=7.4e-4) if WIDE else {}
gG, P, sG = make_channel(sNx, sNy, NR, nSx=nPx, nSy=nPy, OL=OL, land_frac=0.15, buoyancyLinear=BUOY, **extra)
if WIDE:
    rng = np.random.default_rng(7)
    sG["salt"] = exch_xyz(gG.d, (35.0 + np.linspace(-0.5, 0.5, NR)[None, None, :, None, None]
                                 + 0.05 * rng.standard_normal(gG.d.shape3)) * gG.maskC)
opG = ini_cg2d(gG, P)
px, py = rank % nPx, rank // nPx
d = Dims(sNx=sNx, sNy=sNy, OLx=OL, OLy=OL, Nr=NR, nPx=nPx, nPy=nPy, myPx=px, myPy=py)


def mine(a):            # (nSy,nSx,...) global tiling -> my tile as a 1x1 tiling
    return np.ascontiguousarray(a[py:py + 1, px:px + 1])


rt.init(d, local)
for n in rt.GRID_FIELD_NAMES:
    if n in gG.a:
        rt.set_field(n, mine(gG.a[n]))
for n in ("drF", "drC", "recip_drF", "recip_drC"):
    v = np.zeros(NR + 1)
    v[:len(gG.a[n])] = gG.a[n]
    rt.set_field(n, v)
rt.set_params(**{k: P[k] for k in LIB_PARAMS if k in P})
adv = 33 if WIDE else 2
rt.set_params(deltaTtracer=P["deltaTtracer"], tempAdvScheme=adv, tempVertAdvScheme=adv, saltAdvScheme=adv,
              saltVertAdvScheme=adv, nIter0=0)
distributed.setup(d)
distributed.set_salt_stepping(bool(WIDE))
rt.set_cg2d_operator({k: (mine(v) if isinstance(v, np.ndarray) else v) for k, v in opG.items()})
for n in ("uVel", "vVel", "wVel", "theta", "etaN", "surfForcU", "surfForcV"):
    rt.set_field(n, mine(sG[n]))
rt.fill_field("kappaRU", P["viscAr"]); rt.fill_field("kappaRV", P["viscAr"]); rt.fill_field("kappaRT", P["diffKrT"])
for n in ("tRef", "sRef", "rF", "rC"):      # linear EOS + CALC_PHI_HYD inputs
    src = sG.get(n, gG.a.get(n))
    v = np.zeros(NR + 1)
    v[:len(src)] = src
    rt.set_field(n, v)
for n in ("gU", "gV", "guNm1", "gvNm1", "gtNm1", "theta2", "cg2d_b", "cg2d_x"):
    rt.fill_field(n, 0.0)
FIELDS = ("uVel", "vVel", "wVel", "theta", "etaN") + (("salt",) if WIDE else ())
if WIDE:
    rt.set_field("salt", mine(sG["salt"]))
    rt.fill_field("kappaRS", P["diffKrS"]); rt.fill_field("gsNm1", 0.0); rt.fill_field("salt2", 0.0)
res = [distributed.forward_step(it) for it in range(nsteps)]
out = {n: rt.get_field(n, np.zeros(d.shape2 if n == "etaN" else d.shape3)) for n in FIELDS}
rt.finalize()
dist.barrier()
ok = True
# every rank checks its tile against the single-process run of the same global domain (rank 0 computes it)
if rank == 0:
    m = Model(gG, P, sG, opG, device=local)
    ref_res = [m.step() for _ in range(nsteps)]
    ref = {n: m.get(n) for n in out}
    m.close()
    payload = [ref, ref_res]
else:
    payload = [None, None]
dist.broadcast_object_list(payload, src=0)
ref, ref_res = payload
for n, a in out.items():
    b = ref[n][py:py + 1, px:px + 1]
    err = np.abs(a - b).max() / max(np.abs(ref[n]).max(), 1e-300)
    if err > 1e-9:
        ok = False
    at = np.unravel_index(np.abs(a - b).argmax(), a.shape)[2:]
    print(f"rank {rank} {n}: rel err {err:.2e} at (k,j,i)={tuple(int(x) for x in at)} tile sum {np.abs(a).sum():.6e}")
print(f"rank {rank} iters {[r['numIters'] for r in res]} ref {[r['numIters'] for r in ref_res]}")
if any(abs(r['numIters'] - q['numIters']) > 1 for r, q in zip(res, ref_res)):
    ok = False
flag = torch.tensor([0 if ok else 1], device="cuda")
dist.all_reduce(flag)
if rank == 0:
    print("DIST_CHECK", "PASS" if flag.item() == 0 else "FAIL")
dist.destroy_process_group()
sys.exit(0 if flag.item() == 0 else 1)
