"""End-to-end oracle run of verification/tutorial_advection_in_gyre: the barotropic gyre of config 1 RESTARTED from
a 10-year spin-up (pickup.0000259200: |u| up to 0.26 m/s), 4 steps.

TEST INFRASTRUCTURE ONLY.  Pins, against the experiment's golden output (results/output.txt: `cg2d: Sum(rhs),rhsMax`,
cg2d_init_res, iteration counts 11 10 10 11, %MON dynstat_{eta,uvel,vvel,wvel}_* at steps 259200 .. 259204), what config 1
(started from rest) only touches with round-off-sized numbers:
  * MOM_FLUXFORM's advective terms on a developed, non-linear flow, with harmonic viscosity and no-slip sides;
  * the no-slip BOTTOM drag (no_slip_bottom with viscAz = 1e-2: mom_u_botdrag_coeff / MOM_{U,V}_BOTTOMDRAG);
  * a restart: AB2 continues from the pickup's GuNm1 / GvNm1 with the regular weight (adams_bashforth2.F:61-65,
    startAB = 1), the first guess of CG2D is Bo_surf * etaN of the pickup (solve_for_pressure.F).
Parameters from input/data and the golden's parameter summary: deltaT = 1200, abEps = 0.1, viscAh = 400,
viscAz = 0.01, beta = 1e-11, f0 = 1e-4, gravity = gBaro = 9.81, rhoConst = 999.8, cg2dTargetResidual = 1e-10,
2 x 2 tiles of 30 x 30, OL = 4.  theta is uniform (20 degC: no buoyancy forcing, dynstat_theta_sd = 0 in the golden) and
salt is zero, so the thermodynamics is not stepped here; the passive ptracer of the experiment uses the second-order
moment scheme (80), which is outside the path.
"""
from __future__ import annotations

import os

import numpy as np

from mitgcm_b200.grid import Dims, cartesian_grid, masks_from_depth, global_area
from .barotropic_gyre import mon_stats, tile_field
from .pyoracle import Oracle

INPUTS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "inputs")
FIXTURE = os.path.join(INPUTS, "advection_in_gyre.npz")
# tutorial_advection_in_gyre; oracle/matrix_example.py runs the same sequence with its own numbers
CONFIG = dict(fixture=FIXTURE, n=60, dx=20e3, tiles=(2, 2), OL=4, ygOrigin=0.0, deltaT=1200.0, viscAh=400.0, viscAr=1e-2,
              abEps=0.1, rhoConst=999.8, tol=1e-10)


def setup(cfg=CONFIG):
    z = np.load(cfg["fixture"])
    n, (nSx, nSy) = cfg["n"], cfg["tiles"]
    d = Dims(sNx=n // nSx, sNy=n // nSy, OLx=cfg["OL"], OLy=cfg["OL"], nSx=nSx, nSy=nSy, Nr=1)
    g = cartesian_grid(d, [cfg["dx"]] * n, [cfg["dx"]] * n, [5000.0], ygOrigin=cfg["ygOrigin"], f0=1e-4, beta=1e-11,
                       gBaro=9.81)
    masks_from_depth(g, z["topog"], hFacMin=1.0, hFacMinDr=0.0)
    P = dict(deltaTMom=cfg["deltaT"], deltaTFreeSurf=cfg["deltaT"], viscAhD=cfg["viscAh"], viscAhZ=cfg["viscAh"],
             no_slip_sides=1, sideDragFactor=2.0, no_slip_bottom=1, selectBotDragQuadr=-1,
             cg2dTargetResidual=cfg["tol"], globalArea=global_area(g))
    return z, d, g, P


def stats(d, g, etaN, uVel, vVel, wVel):
    maskInC, maskInW, maskInS = g.maskC[:, :, 0], g.maskW[:, :, 0], g.maskS[:, :, 0]
    return dict(eta=mon_stats(d, etaN[:, :, None], maskInC[:, :, None], maskInC, g.rA, g.drF),
                uvel=mon_stats(d, uVel, g.hFacW, maskInW, g.rAw, g.drF),
                vvel=mon_stats(d, vVel, g.hFacS, maskInS, g.rAs, g.drF),
                wvel=mon_stats(d, wVel, g.maskC, maskInC, g.rA, g.drC[:1]))


def run(nSteps=4, engine=None, cfg=CONFIG):
    """Returns (cg2dNorm, statistics of the pickup state, [per-step dict]).  `engine` as in baroclinic_gyre.py
    (mom_fluxform / cg2d from the CUDA library; it must not be used to claim oracle parity)."""
    z, d, g, P = setup(cfg)
    o = Oracle(g, P)
    e = engine or o
    op = o.ini_cg2d()
    if engine is not None and hasattr(engine, "setup"):
        engine.setup(g, o.params, op)
    if engine is not None and getattr(engine, "fb", 0) is None:
        engine.fb = o                                # routines the engine does not replace stay on this oracle
    viscAr, abEps, rhoConst = cfg["viscAr"], cfg["abEps"], cfg["rhoConst"]
    t3 = lambda a: tile_field(d, a)[:, :, None].copy()
    uVel, vVel = t3(z["Uvel"]), t3(z["Vvel"])
    guNm1, gvNm1 = t3(z["GuNm1"]), t3(z["GvNm1"])
    etaN = tile_field(d, z["EtaN"])
    sfU = tile_field(d, z["windx"]) * (1.0 / rhoConst)
    sfV = np.zeros(d.shape2)
    wVel, gU, gV = (np.zeros(d.shape3) for _ in range(3))
    for bj in range(1, d.nSy + 1):          # INTEGR_CONTINUITY at start-up (ini_fields / initialise_varia): w of the pickup flow
        for bi in range(1, d.nSx + 1):
            o.integrate_for_w(bi, bj, uVel, vVel, wVel)
    o.exch_xyz(wVel, d.Nr)
    kappaR = np.full((d.Nr + 1, d.PY, d.PX), viscAr)
    dPhi = np.zeros((d.PY, d.PX))
    first = stats(d, g, etaN, uVel, vVel, wVel)
    out = []
    abFac = 0.5 + abEps                      # restart: startAB = 1
    for it in range(nSteps):
        for bj in range(1, d.nSy + 1):
            for bi in range(1, d.nSx + 1):
                fVerU = np.zeros((2, d.PY, d.PX))
                fVerV = np.zeros((2, d.PY, d.PX))
                for k in range(1, d.Nr + 1):
                    kUp, kDown = 1 + (k + 1) % 2, 1 + k % 2
                    guDiss, gvDiss = np.zeros((d.PY, d.PX)), np.zeros((d.PY, d.PX))
                    e.mom_fluxform(bi, bj, k, 0, d.sNx + 1, 0, d.sNy + 1, kappaR, kappaR,
                                   fVerU[kUp - 1], fVerV[kUp - 1], fVerU[kDown - 1], fVerV[kDown - 1],
                                   guDiss, gvDiss, uVel, vVel, wVel, gU, gV)
                    o.timestep(bi, bj, k, 0, d.sNx + 1, 0, d.sNy + 1, dPhi, dPhi, guDiss, gvDiss, sfU, sfV,
                               1, 1, abFac, uVel, vVel, gU, gV, guNm1, gvNm1)
        b, x = np.zeros(d.shape2), np.zeros(d.shape2)
        for bj in range(1, d.nSy + 1):
            for bi in range(1, d.nSx + 1):
                o.solve_rhs(bi, bj, etaN, gU, gV, b, x)
        res = e.cg2d(op, b, x, 1000, -1)
        o.exch_xyz(x)
        etaN = g.recip_Bo * x
        for bj in range(1, d.nSy + 1):
            for bi in range(1, d.nSx + 1):
                o.correction_step(bi, bj, etaN, gU, gV, uVel, vVel)
                o.integrate_for_w(bi, bj, uVel, vVel, wVel)
        o.exch_xyz(uVel, d.Nr)
        o.exch_xyz(vVel, d.Nr)
        o.exch_xyz(wVel, d.Nr)
        rec = dict(res)
        rec.update(stats(d, g, etaN, uVel, vVel, wVel))
        out.append(rec)
    return op["cg2dNorm"], first, out
