"""End-to-end oracle run of verification/adjustment.cs-32x32x1: barotropic adjustment on the cs32
cubed sphere (6 facets of 32x32 cut into 48 tiles of 16x8, OL = 2, pkg/exch2), Nr = 1.

TEST INFRASTRUCTURE ONLY.  Pins, against the experiment's golden output (results/output.txt:
`CG2D normalisation factor`, 24 steps of `cg2d: Sum(rhs),rhsMax`, cg2d_init_res, iteration counts and
%MON dynstat_{eta,uvel,vvel,wvel}_*), what no other reachable experiment exercises:
  * the exch2 tile graph in a time-stepping run: EXCH_XY_RL (etaN, wVel), EXCH_UV_XYZ_RL with signs
    (uVel, vVel), EXCH_UV_XY_RS without signs (grid metrics, hFacW/S, CG2D operators);
  * CG2D on the cubed sphere (EXCH2_S3D_RX inside the solver), tolerance 1e-13;
  * the Crank-Nicolson free surface: implicSurfPress = implicDiv2DFlow = 0.5 with exactConserv
    (timestep.F psFac term, calc_div_ghat.F exactConserv branch, update_etah.F, correction_step.F);
  * MOM_FLUXFORM on a curvilinear grid with momAdvection = F, no viscosity: Coriolis only.
Run-time switches read from input/data and the golden's parameter summary: deltaT = 900, abEps = 0.1,
gravity = gBaro = 9.8184, rhoConst = 1000, delR = 1366, f = 2 Omega sin(lat), no forcing,
tempStepping = saltStepping = F, cg2dTargetResidual = 1e-13, cg2dMaxIters = 600.

Sequence per step (forward_step.F, non-staggered): DYNAMICS (CALC_GRAD_PHI_SURF because
implicSurfPress != 1, MOM_FLUXFORM, TIMESTEP), SOLVE_FOR_PRESSURE, MOMENTUM_CORRECTION_STEP,
INTEGR_CONTINUITY (dEtaHdt, etaN, wVel, UPDATE_ETAH), DO_FIELDS_BLOCKING_EXCHANGES.
`engine` as in baroclinic_gyre.py (mom_fluxform / cg2d from the CUDA library).
"""
from __future__ import annotations

import os

import numpy as np

from mitgcm_b200.grid import Dims, cubed_sphere_grid, cube_masks_from_depth, global_area
from mitgcm_b200.exch2 import cubed_sphere_topology
from mitgcm_b200.model import ini_cg2d_tilegraph
from . import exch2_oracle as eo
from .pyoracle import Oracle
from .baroclinic_gyre import mon_stats

FIXTURE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "inputs",
                       "cs32_grid_bathy.npz")
KEEP = "xC yC rA xG yG dxC dyC dxG dyG rAw rAs".split()


def tile_from_xstack(T, d, glob):
    """W2_mapIO = -1 global layout (facets stacked along x) -> interior of a tiled array."""
    a = np.zeros(d.shape2)
    n = glob.shape[0]
    for t in range(T.nTiles):
        f, bx, by = int(T.myFace[t]) - 1, int(T.tBasex[t]), int(T.tBasey[t])
        a[0, t, d.OLy:d.OLy + d.sNy, d.OLx:d.OLx + d.sNx] = glob[by:by + d.sNy, f * n + bx:f * n + bx + d.sNx]
    return a


def setup():
    z = np.load(FIXTURE)
    faces = [{n: z[f"{n}_{f}"] for n in KEEP} for f in range(6)]
    T = cubed_sphere_topology(32, 16, 8)
    d = Dims(sNx=16, sNy=8, OLx=2, OLy=2, nSx=48, nSy=1, Nr=1)
    g = cubed_sphere_grid(d, T, faces, [1366.0], gBaro=9.8184)
    cube_masks_from_depth(g, T, z["adj_bathy_f2"], hFacMin=1.0, hFacMinDr=0.0)
    P = dict(deltaTMom=900.0, deltaTFreeSurf=900.0, implicSurfPress=0.5, implicDiv2DFlow=0.5,
             cg2dTargetResidual=1e-13, momAdvection=0, momViscosity=0, no_slip_sides=0, no_slip_bottom=0,
             viscAhD=0.0, viscAhZ=0.0, selectBotDragQuadr=-1, globalArea=global_area(g))
    return T, d, g, P, z["adj_ssh_eq"]


def run(nSteps=24, engine=None):
    """Returns (cg2dNorm, [per-step dict of solver scalars and monitor statistics])."""
    T, d, g, P, ssh = setup()
    OL = d.OLx
    o = Oracle(g, P)
    hook = eo.Exch2Hook(o, T, OL)                 # EXCH2 inside the oracle's CG2D
    e = engine or o
    op = ini_cg2d_tilegraph(g, P, T)
    if engine is not None and hasattr(engine, "setup"):
        engine.setup(g, o.params, op, T)
    if engine is not None and getattr(engine, "fb", 0) is None:
        engine.fb = o                                # routines the engine does not replace stay on this oracle
    abEps = 0.1
    ns = (d.PY, d.PX)
    tiles = [(bi, 1) for bi in range(1, d.nSx + 1)]
    z3 = lambda: np.zeros(d.shape3)
    uVel, vVel, wVel, gU, gV, guNm1, gvNm1 = (z3() for _ in range(7))
    etaN = tile_from_xstack(T, d, ssh)
    eo.exch2_3d(T, etaN[0][:, None], OL)           # ini_psurf.F: _EXCH_XY_RL(etaN)
    etaH = etaN.copy()
    dEtaHdt = np.zeros(d.shape2)
    kap = np.zeros((d.Nr + 1,) + ns)
    sfU = np.zeros(d.shape2)
    zero = np.zeros(ns)
    maskInC, maskInW, maskInS = g.maskC[:, :, 0], g.maskW[:, :, 0], g.maskS[:, :, 0]
    out = []
    try:
        for it in range(nSteps):
            abFac = 0.0 if it == 0 else 0.5 + abEps
            # ---- DYNAMICS
            for bi, bj in tiles:
                fVerU, fVerV = np.zeros((2,) + ns), np.zeros((2,) + ns)
                phiX, phiY = np.zeros(ns), np.zeros(ns)
                o.calc_grad_phi_surf(bi, bj, 0, d.sNx + 1, 0, d.sNy + 1, etaN, phiX, phiY)
                guDiss, gvDiss = np.zeros(ns), np.zeros(ns)
                e.mom_fluxform(bi, bj, 1, 0, d.sNx + 1, 0, d.sNy + 1, kap, kap, fVerU[1], fVerV[1], fVerU[0], fVerV[0],
                               guDiss, gvDiss, uVel, vVel, wVel, gU, gV)
                o.timestep(bi, bj, 1, 0, d.sNx + 1, 0, d.sNy + 1, zero, zero, guDiss, gvDiss, sfU, sfU, 1, 1, abFac,
                           uVel, vVel, gU, gV, guNm1, gvNm1, phiX, phiY)
            # ---- SOLVE_FOR_PRESSURE
            b, x = np.zeros(d.shape2), np.zeros(d.shape2)
            for bi, bj in tiles:
                o.solve_rhs(bi, bj, etaN, gU, gV, b, x, etaH=etaH)
            res = e.cg2d(op, b, x, 600, -1)
            eo.exch2_3d(T, x[0][:, None], OL)
            etaN = g.recip_Bo * x
            # ---- MOMENTUM_CORRECTION_STEP, INTEGR_CONTINUITY
            for bi, bj in tiles:
                o.correction_step(bi, bj, etaN, gU, gV, uVel, vVel)
            for bi, bj in tiles:
                o.integr_continuity_ec(bi, bj, uVel, vVel, etaH, dEtaHdt, etaN, True)
                o.integrate_for_w(bi, bj, uVel, vVel, wVel)
            eo.exch2_3d(T, etaN[0][:, None], OL)
            # UPDATE_ETAH with implicDiv2DFlow != 1 (update_etah.F:60-68): interior, then EXCH_XY_RL
            jj, ii = d.interior()
            etaH = etaH.copy()
            etaH[:, :, jj, ii] = etaN[:, :, jj, ii] + (1.0 - P["implicDiv2DFlow"]) * dEtaHdt[:, :, jj, ii] * P["deltaTFreeSurf"]
            eo.exch2_3d(T, etaH[0][:, None], OL)
            # ---- DO_FIELDS_BLOCKING_EXCHANGES
            eo.exch2_uv_3d(T, uVel[0], vVel[0], OL, True)
            eo.exch2_3d(T, wVel[0], OL)
            rec = dict(res)
            rec["eta"] = mon_stats(d, etaN[:, :, None], maskInC[:, :, None], maskInC, g.rA, [g.drF[0]])
            rec["uvel"] = mon_stats(d, uVel, g.hFacW, maskInW, g.rAw, g.drF)
            rec["vvel"] = mon_stats(d, vVel, g.hFacS, maskInS, g.rAs, g.drF)
            rec["wvel"] = mon_stats(d, wVel, g.maskC, maskInC, g.rA, g.drC[:1])
            out.append(rec)
    finally:
        hook.close()
    return op["cg2dNorm"], out
