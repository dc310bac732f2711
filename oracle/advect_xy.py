"""Oracle run of verification/advect_xy (salt only): multi-dimensional advection (GAD_ADVECTION,
saltAdvScheme = 33, built with GAD_MULTIDIM_COMPRESSIBLE) of a blob by the uniform flow u = v = 1 m/s on
a doubly periodic 20 x 20 x 1 Cartesian grid (2 tiles of 20 x 10, OL = 3), momStepping = F, 80 steps.

TEST INFRASTRUCTURE ONLY.  Pins GAD_ADVECTION and the DST3 flux-limiter leaves (gad_dst3fl_adv_{x,y}.F,
shared with GAD_CALC_RHS) against the experiment's golden output (results/output.txt:
%MON dynstat_salt_{max,min,mean,sd} at steps 0, 16, ..., 80).  The experiment's temperature uses the
second-order-moment scheme (80), which is outside the path and is not run.
Set-up read from input/data, code/SIZE.h, code/ini_salt.F, code/ini_vel.F: dXspacing = dYspacing = 10 km,
delZ = 10 km, deltaT = 2500 s, salt = 35 + 1 inside the 60 km sphere around (40 km, 40 km, -50 km).
`advect` replaces the oracle's GAD_ADVECTION (same signature) so a GPU test can put the CUDA kernel in."""
from __future__ import annotations

import numpy as np

from mitgcm_b200.grid import Dims, cartesian_grid, masks_from_depth
from .pyoracle import Oracle
from .baroclinic_gyre import mon_stats


def setup(nSx=1, nSy=2):
    d = Dims(sNx=20 // nSx, sNy=20 // nSy, OLx=3, OLy=3, nSx=nSx, nSy=nSy, Nr=1)
    g = cartesian_grid(d, [10e3] * 20, [10e3] * 20, [10e3], f0=0.0, beta=0.0)
    masks_from_depth(g, -10e3 * np.ones((20, 20)), hFacMin=1.0, hFacMinDr=0.0)
    xC, yC, rC = g.a["xC"], g.a["yC"], g.a["rC"]
    rD = np.sqrt((xC - 40e3) ** 2 + (yC - 40e3) ** 2 + (rC[0] + 50e3) ** 2)
    salt = np.where(rD <= 60e3, 36.0, 35.0)[:, :, None].copy()
    return d, g, salt


def run(nSteps=80, scheme=33, compressible=True, advect=None, nSx=1, nSy=2, every=16):
    """Returns the list of monitor statistics of salt at steps 0, every, 2*every, ..."""
    d, g, salt = setup(nSx, nSy)
    o = Oracle(g, {})
    o.exch_xyz(salt, d.Nr)
    uVel = np.ones(d.shape3) * g.maskW
    vVel = np.ones(d.shape3) * g.maskS
    wVel = np.zeros(d.shape3)
    dT = np.full(d.Nr, 2500.0)
    maskInC = g.maskC[:, :, 0]
    fn = advect or o.gad_advection
    out = [mon_stats(d, salt, g.hFacC, maskInC, g.rA, g.drF)]
    for it in range(nSteps):
        new = salt.copy()
        for bj in range(1, d.nSy + 1):
            for bi in range(1, d.nSx + 1):
                gS = np.zeros((d.Nr, d.PY, d.PX))
                rc = fn(bi, bj, scheme, scheme, 0, int(compressible), dT, uVel, vVel, wVel, salt, gS)
                assert not rc
                new[bj - 1, bi - 1] = salt[bj - 1, bi - 1] + dT[:, None, None] * gS     # TIMESTEP_TRACER
        salt = new
        o.exch_xyz(salt, d.Nr)
        if (it + 1) % every == 0:
            out.append(mon_stats(d, salt, g.hFacC, maskInC, g.rA, g.drF))
    return out
