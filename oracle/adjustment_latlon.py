"""End-to-end oracle run of verification/adjustment.128x64x1: barotropic (gravity-wave) adjustment of a one-layer
atmosphere in p coordinates on the global spherical-polar grid, 128 x 64 cells of 2.8125 degrees from pole to pole,
2 x 2 tiles of 64 x 32, OL = 2, Nr = 1.

TEST INFRASTRUCTURE ONLY.  Pins, against the experiment's golden output (results/output.txt, written by an older
SOLVE_FOR_PRESSURE: `cg2d_init_res`, `cg2d_iters`, `cg2d_res` per step, and %MON dynstat_{eta,uvel,vvel}_*):
  * INI_SPHERICAL_POLAR_GRID from pole to pole (the zero-width cell faces at yG = +-90) with INI_CG2D on it,
  * CG2D at 128 x 64 with 2 x 2 tiles, tolerance 1e-12,
  * MOM_FLUXFORM with momAdvection = F, useCoriolis = F: metric terms only,
  * the p-coordinate free surface (uniformLin_PhiSurf: Bo_surf = 1/rhoConst = 1, rkSign = -1).
Run-time switches from input/data: deltaT = 450, abEps = 0.1, 24 steps, initial eta from ps.init.
theta = tRef is uniform and not stepped: the hydrostatic pressure gradient is zero to the bit."""
from __future__ import annotations

import os

import numpy as np

from mitgcm_b200.grid import Dims, spherical_polar_grid, set_hfac, global_area
from .pyoracle import Oracle
from .baroclinic_gyre import mon_stats, tile_field

FIXTURE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "inputs", "adjustment_128x64_ps_init.npy")


def setup(nSx=2, nSy=2):
    d = Dims(sNx=128 // nSx, sNy=64 // nSy, OLx=2, OLy=2, nSx=nSx, nSy=nSy, Nr=1)
    g = spherical_polar_grid(d, [2.8125] * 128, [2.8125] * 64, [1.0e5], xgOrigin=0.0, ygOrigin=-90.0, rSphere=6370e3,
                             rotationPeriod=86400.0, gBaro=9.81)
    one = np.ones(d.shape3)
    # ADD_WALLS2MASKS (add_walls2masks.F:37-57): the zero-width faces at the poles (dxG = 0) are closed
    hS = np.where(g.a["dxG"][:, :, None] == 0.0, 0.0, 1.0)
    hW = np.where(g.a["dyG"][:, :, None] == 0.0, 0.0, 1.0)
    set_hfac(g, one, hW, hS)
    g.a["Bo_surf"] = np.ones(d.shape2)                       # uniformLin_PhiSurf: 1/rhoConst
    g.a["recip_Bo"] = np.ones(d.shape2)
    P = dict(deltaTMom=450.0, deltaTFreeSurf=450.0, rkSign=-1.0, cg2dTargetResidual=1e-12, momAdvection=0, momViscosity=1,
             no_slip_sides=0, no_slip_bottom=0, viscAhD=0.0, viscAhZ=0.0, selectBotDragQuadr=-1, cfFacMom=0.0,
             usingSphericalPolarGrid=1, selectMetricTerms=1, recip_rSphere=1.0 / 6370e3, globalArea=global_area(g))
    return d, g, P, np.load(FIXTURE)


def run(nSteps=24, nSx=2, nSy=2, engine=None):
    """Returns (cg2dNorm, [per-step dict of solver scalars and monitor statistics])."""
    d, g, P, ps = setup(nSx, nSy)
    o = Oracle(g, P)
    e = engine or o
    op = o.ini_cg2d()
    if engine is not None and hasattr(engine, "setup"):
        engine.setup(g, o.params, op)
    abEps = 0.1
    ns = (d.PY, d.PX)
    tiles = [(bi, bj) for bj in range(1, nSy + 1) for bi in range(1, nSx + 1)]
    z3 = lambda: np.zeros(d.shape3)
    uVel, vVel, wVel, gU, gV, guNm1, gvNm1 = (z3() for _ in range(7))
    etaN = tile_field(d, ps)
    o.exch_xyz(etaN)                                        # ini_psurf.F: _EXCH_XY_RL(etaN)
    kap = np.zeros((d.Nr + 1,) + ns)
    sfU = np.zeros(d.shape2)
    zero = np.zeros(ns)
    maskInC, maskInW, maskInS = g.maskC[:, :, 0], g.maskW[:, :, 0], g.maskS[:, :, 0]
    out = []
    for it in range(nSteps):
        abFac = 0.0 if it == 0 else 0.5 + abEps
        for bi, bj in tiles:
            fVerU, fVerV = np.zeros((2,) + ns), np.zeros((2,) + ns)
            guDiss, gvDiss = np.zeros(ns), np.zeros(ns)
            e.mom_fluxform(bi, bj, 1, 0, d.sNx + 1, 0, d.sNy + 1, kap, kap, fVerU[1], fVerV[1], fVerU[0], fVerV[0],
                           guDiss, gvDiss, uVel, vVel, wVel, gU, gV)
            o.timestep(bi, bj, 1, 0, d.sNx + 1, 0, d.sNy + 1, zero, zero, guDiss, gvDiss, sfU, sfU, 1, 1, abFac,
                       uVel, vVel, gU, gV, guNm1, gvNm1)
        b, x = np.zeros(d.shape2), np.zeros(d.shape2)
        for bi, bj in tiles:
            o.solve_rhs(bi, bj, etaN, gU, gV, b, x)
        res = e.cg2d(op, b, x, 600, -1)
        o.exch_xyz(x)
        etaN = g.recip_Bo * x
        for bi, bj in tiles:
            o.correction_step(bi, bj, etaN, gU, gV, uVel, vVel)
            o.integrate_for_w(bi, bj, uVel, vVel, wVel)
        for f in (uVel, vVel, wVel):
            o.exch_xyz(f, d.Nr)
        rec = dict(res)
        rec["eta"] = mon_stats(d, etaN[:, :, None], maskInC[:, :, None], maskInC, g.rA, [g.drF[0]])
        rec["uvel"] = mon_stats(d, uVel, g.hFacW, maskInW, g.rAw, g.drF)
        rec["vvel"] = mon_stats(d, vVel, g.hFacS, maskInS, g.rAs, g.drF)
        out.append(rec)
    return op["cg2dNorm"], out
