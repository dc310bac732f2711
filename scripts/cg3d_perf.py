"""CG3D timing on a synthetic doubly-periodic flat-bottom box (device-resident b, x).
usage: python scripts/cg3d_perf.py [N=512] [Nr=50] [iters=50] [reps=3] [sweep|smem]
sweep: time the four-sweep kernel and every variant of the fused kernel (csrc/cg3d.cu) in one process.
smem: the default variant with 0, 4, 8, ... levels of z' in shared memory."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from mitgcm_b200 import runtime as rt
from mitgcm_b200.grid import Dims

N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
Nr = int(sys.argv[2]) if len(sys.argv) > 2 else 50
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 50
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
d = Dims(sNx=N, sNy=N, OLx=2, OLy=2, Nr=Nr)
rt.init(d)
dx, dz, dt, gB = 20.0, 20.0, 20.0, 10.0
# INI_CG3D closed form for a uniform grid without land (ini_cg3d.F): aW = aS = dy*dz/dx, aV = dx*dy/dz, normalised
aH, aV = dx * dz / dx, dx * dx / dz
norm = 1.0 / max(aH, aV)
full = lambda v: np.full(d.shape3, v)
aVk = full(aV * norm); aVk[:, :, 0] = 0.0
aC = np.zeros(d.shape3)
for k in range(Nr):
    aU = aV if k > 0 else 0.0
    aL = aV if k < Nr - 1 else 0.0
    c = -aH - aH - aH - aH - aU - aL
    if k == 0:
        c = c - 1.0 * (1.0 / gB) * dx * dx * 1.0 / dt / dt
    aC[:, :, k] = c * norm
zMC, zML, zMU = aC.copy(), aVk.copy(), np.zeros(d.shape3)
zMU[:, :, :-1] = aVk[:, :, 1:]
zMC[:, :, 0] = 1.0 / zMC[:, :, 0]
zMU[:, :, 0] *= zMC[:, :, 0]
for k in range(1, Nr):
    zMC[:, :, k] = 1.0 / (zMC[:, :, k] - zML[:, :, k] * zMU[:, :, k - 1])
    zMU[:, :, k] *= zMC[:, :, k]
rt.set_cg3d_operator(dict(aW3d=full(aH * norm), aS3d=full(aH * norm), aV3d=aVk, aC3d=aC, zMC=zMC, zML=zML, zMU=zMU,
                          cg3dNorm=norm, cg3dTolerance_sq=0.0, cg3dNormaliseRHS=True))
rt.fill_field("maskC", 1.0)
g = torch.Generator(device="cuda").manual_seed(1)
jj, ii = d.interior()
sweep = len(sys.argv) > 5 and sys.argv[5] == "sweep"
configs = [("default", {})]
if sweep:
    configs = [("four-sweep", {"MITGCM_B200_CG3D_UNFUSED": "1"})] + [
        (f"fused variant {v}", {"MITGCM_B200_CG3D_VARIANT": str(v)}) for v in range(4)]
if len(sys.argv) > 5 and sys.argv[5] == "smem":
    configs = [(f"fused, <= {n} levels in shared memory", {"MITGCM_B200_CG3D_SMEM_LEVELS": str(n)}) for n in (0, 4, 8, 12, 16, 24, 32)]
for label, env in [c for c in configs for _ in range(reps)]:
    for k in ("MITGCM_B200_CG3D_UNFUSED", "MITGCM_B200_CG3D_VARIANT", "MITGCM_B200_CG3D_SMEM_LEVELS"):
        os.environ.pop(k, None)
    os.environ.update(env)
    b = torch.zeros(d.shape3, dtype=torch.float64, device="cuda")
    b[..., jj, ii] = torch.randn((1, 1, Nr, N, N), dtype=torch.float64, device="cuda", generator=g)
    x = torch.zeros_like(b)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    r = rt.cg3d(b, x, iters)
    torch.cuda.synchronize()
    t = time.perf_counter() - t0
    cells = N * N * Nr
    print(f"{label}: N={N} Nr={Nr} iters={r['numIters']} time={t * 1e3:.2f} ms  {t / r['numIters'] * 1e6:.1f} us/it  "
          f"{168.0 * cells * r['numIters'] / t / 1e9:.1f} GB/s (168 B/cell/it) = {168.0 * cells * r['numIters'] / t / 1e9 / 6556.2 * 100:.1f}% of measured HBM peak; "
          f"res {r['firstResidual']:.3e}->{r['lastResidual']:.3e}")
rt.finalize()
