"""Quick CG2D timing on a synthetic doubly-periodic flat-bottom grid (device-resident b, x).
usage: python scripts/cg2d_perf.py [N=2048] [iters=200] [reps=3] [sr=0]"""
import sys
import time

import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from mitgcm_b200 import runtime as rt
from mitgcm_b200.grid import Dims

N = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 200
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
sr = bool(int(sys.argv[4])) if len(sys.argv) > 4 else False


def flat_operator(d, dx=20e3, H=5000.0, dt=1200.0, g=9.81):
    """INI_CG2D closed form for a flat-bottom uniform Cartesian grid (ini_cg2d.F:85-229)."""
    aW = (dx * H / dx)
    norm = 1.0 / aW
    aWn = aW * norm
    aC = -(aWn + aWn + aWn + aWn + 1.0 * norm * (1.0 / g) * (dx * dx) / dt / dt)
    pC = 1.0 / aC
    t = 0.51 * (aC + aC)
    pW = -aWn / (t * t)
    full = lambda v: np.full(d.shape2, v)
    return dict(aW2d=full(aWn), aS2d=full(aWn), aC2d=full(aC), pW=full(pW), pS=full(pW), pC=full(pC),
                cg2dNorm=norm, cg2dTolerance_sq=0.0, cg2dNormaliseRHS=True)


d = Dims(sNx=N, sNy=N, OLx=2, OLy=2)
rt.init(d)
rt.set_cg2d_operator(flat_operator(d))
rng = np.random.default_rng(20261018)
b = np.zeros(d.shape2)
jj, ii = d.interior()
b[:, :, jj, ii] = rng.standard_normal((1, 1, N, N))
b -= b[:, :, jj, ii].mean() * (b != 0)
b0 = torch.from_numpy(b).cuda()
x0 = torch.zeros_like(b0)
for rep in range(reps):
    bd, xd = b0.clone(), x0.clone()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    r = rt.cg2d(bd, xd, iters, -1, sr=sr)
    dt = time.perf_counter() - t0
    n = r["numIters"]
    gbs = 136.0 * N * N * n / dt / 1e9
    print(f"N={N} sr={sr} iters={n} time={dt*1e3:.2f} ms  {n/dt:.1f} it/s  {dt/n*1e6:.1f} us/it  "
          f"{gbs:.1f} GB/s (136 B/pt/it, SURVEY 8d) = {gbs/6556.2*100:.1f}% of measured HBM peak; "
          f"res {r['firstResidual']:.3e}->{r['lastResidual']:.3e}")
rt.finalize()
