"""End-to-end oracle run of verification/flt_example (the ocean underneath the float package, which only follows the
flow): a wind-driven f-plane channel over a Gaussian bump, 80 x 42 x 8 cells of 5 km x 562.5 m, PARTIAL CELLS
(hFacMin = hFacMinDr = 0.2), 2 x 2 tiles of 40 x 21, OL = 2, stratified start from rest, 18 steps.

TEST INFRASTRUCTURE ONLY.  Pins, against the experiment's golden output (results/output.with_flt.txt), time stepping on
partial cells -- hFacC / hFacW / hFacS < 1 in MOM_FLUXFORM (advection, harmonic + vertical viscosity, free slip), in
GAD_CALC_RHS (centred advection, Laplacian and EXPLICIT vertical diffusion), in the continuity integration and in the CG2D
operator -- which the other reachable experiments have at the operator level only (configs 3 and 4).
Sequence per step (forward_step.F, non-staggered, no exactConserv): FIND_RHO, THERMODYNAMICS (TEMP_INTEGRATE; salinity
is uniform and sBeta = 0, so SALT_INTEGRATE does not feed back and is not stepped), DYNAMICS (CALC_PHI_HYD,
MOM_FLUXFORM, TIMESTEP), SOLVE_FOR_PRESSURE, MOMENTUM_CORRECTION_STEP, INTEGR_CONTINUITY, exchanges.
Parameters from input/data and the golden's summary: deltaT = 600, abEps = 0.1, viscAh = 1e3, viscAz = 1e-3,
diffKhT = 1e3, diffKzT = 1e-5, f0 = 1e-4, beta = 0, tAlpha = 2e-4, gravity = 9.81, rhoConst = rhoNil = 999.8,
cg2dTargetResidual = 1e-9.  `engine` as in baroclinic_gyre.py.
"""
from __future__ import annotations

import os

import numpy as np

from mitgcm_b200.grid import Dims, cartesian_grid, masks_from_depth, global_area
from .baroclinic_gyre import mon_stats
from .barotropic_gyre import tile_field
from .pyoracle import Oracle, Eos

FIXTURE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "inputs", "flt_example.npz")
NX, NY, NR = 80, 42, 8
TREF = [0.696834, 0.497738, 0.298643, 0.0995477, -0.0995477, -0.298643, -0.497738, -0.696834]


def setup():
    z = np.load(FIXTURE)
    d = Dims(sNx=40, sNy=21, OLx=2, OLy=2, nSx=2, nSy=2, Nr=NR)
    g = cartesian_grid(d, [5e3] * NX, [5e3] * NY, [562.5] * NR, f0=1e-4, beta=0.0, gBaro=9.81)
    masks_from_depth(g, z["topog"], hFacMin=0.2, hFacMinDr=0.2)
    P = dict(deltaTMom=600.0, deltaTFreeSurf=600.0, viscAhD=1e3, viscAhZ=1e3, no_slip_sides=0, sideDragFactor=2.0,
             no_slip_bottom=0, selectBotDragQuadr=-1, implicitDiffusion=0, cg2dTargetResidual=1e-9,
             globalArea=global_area(g))
    return z, d, g, P


def run(nSteps=18, engine=None, sr=False):
    """Returns (cg2dNorm, statistics of the start state, [per-step dict]).  sr: CG2D_SR instead of CG2D (useSRCGSolver)."""
    z, d, g, P = setup()
    rhoConst = 999.8
    return step_loop(d, g, P, nSteps, engine, tRef=TREF, salt0=35.0, rhoConst=rhoConst, tAlpha=2e-4, gravity=9.81, abEps=0.1,
                     viscAr=1e-3, diffKhT=1e3, diffKrT=1e-5, deltaT=600.0,
                     sfU=tile_field(d, z["windx"]) * (1.0 / rhoConst),          # external_forcing_surf.F:214
                     phi0surf=np.zeros(d.shape2), sr=sr)


def step_loop(d, g, P, nSteps, engine, *, tRef, salt0, rhoConst, tAlpha, gravity, abEps, viscAr, diffKhT, diffKrT, deltaT,
              sfU, phi0surf, sr=False):
    """The hydrostatic step of a linear-EOS box started from rest with theta = tRef(k) (no exactConserv, explicit
    diffusion, AB2 on the tendencies); shared with oracle/inverted_barometer.py."""
    NR = d.Nr
    o = Oracle(g, P)
    e = engine or o
    op = o.ini_cg2d()
    if engine is not None and hasattr(engine, "setup"):
        engine.setup(g, o.params, op)
    if engine is not None and getattr(engine, "fb", 0) is None:
        engine.fb = o
    rhoNil = rhoConst
    recip_rhoConst = 1.0 / rhoConst
    eos = Eos(rhoNil, rhoConst, tAlpha, 0.0)
    tRef, sRef = np.array(tRef), np.full(NR, salt0)
    dT = np.full(NR, deltaT)
    zr = np.zeros(NR)
    ns = (d.PY, d.PX)
    tiles = [(bi, bj) for bj in range(1, d.nSy + 1) for bi in range(1, d.nSx + 1)]
    sfV, sfT = (np.zeros(d.shape2) for _ in range(2))
    z3 = lambda: np.zeros(d.shape3)
    uVel, vVel, wVel, gU, gV, guNm1, gvNm1, gtNm1, rhoInSitu, ivdc = (z3() for _ in range(10))
    theta = np.where(g.maskC != 0.0, tRef[None, None, :, None, None], 0.0)      # ini_theta.F
    salt = np.where(g.maskC != 0.0, salt0, 0.0)
    etaN = np.zeros(d.shape2)
    kapU = np.full((NR + 1,) + ns, viscAr)
    kbl, dkr = np.zeros(NR), np.full(NR, diffKrT)
    maskInC, maskInW, maskInS = g.maskC[:, :, 0], g.maskW[:, :, 0], g.maskS[:, :, 0]
    drF, drC = g.drF, g.drC[:NR]

    def stats():
        return dict(eta=mon_stats(d, etaN[:, :, None], maskInC[:, :, None], maskInC, g.rA, [drF[0]]),
                    uvel=mon_stats(d, uVel, g.hFacW, maskInW, g.rAw, drF),
                    vvel=mon_stats(d, vVel, g.hFacS, maskInS, g.rAs, drF),
                    wvel=mon_stats(d, wVel, g.maskC, maskInC, g.rA, drC),
                    theta=mon_stats(d, theta, g.hFacC, maskInC, g.rA, drF))
    first = stats()
    out = []
    for it in range(nSteps):
        abFac = 0.0 if it == 0 else 0.5 + abEps
        for bi, bj in tiles:
            o.density_ivdc(eos, bi, bj, theta, salt, tRef, sRef, rhoInSitu, ivdc)
        # ---- THERMODYNAMICS / TEMP_INTEGRATE (explicit vertical diffusion)
        for bi, bj in tiles:
            ti = (bj - 1, bi - 1)
            kappaRk = np.zeros((NR,) + ns)
            o.calc_3d_diffusivity(bi, bj, ivdc, 0.0, kbl, dkr, kappaRk)
            gT = np.zeros((NR,) + ns)
            fV = np.zeros((2,) + ns)
            rTrans = np.zeros(ns)
            sl = {n: np.zeros(ns) for n in "xA yA maskUp uFld vFld wFld uTrans vTrans rTransKp1 fZon fMer".split()}
            th = np.ascontiguousarray(theta[ti])
            gNm1 = np.ascontiguousarray(gtNm1[ti])
            for k in range(NR, 0, -1):
                kUp, kDown = 1 + (k + 1) % 2, 1 + k % 2
                o.calc_adv_flow(bi, bj, k, uVel, vVel, wVel, sl["xA"], sl["yA"], sl["maskUp"], sl["uFld"],
                                sl["vFld"], sl["wFld"], sl["uTrans"], sl["vTrans"], rTrans, sl["rTransKp1"])
                e.gad_calc_rhs(bi, bj, 0, d.sNx + 1, 0, d.sNy + 1, k, max(1, k - 1), kUp, kDown, sl["xA"],
                               sl["yA"], sl["maskUp"], sl["uFld"], sl["vFld"], sl["wFld"], sl["uTrans"],
                               sl["vTrans"], rTrans, sl["rTransKp1"], diffKhT, 0.0, kappaRk[k - 1], zr, th, gNm1,
                               dT, 2, 2, 1, 0, 0, 0, sl["fZon"], sl["fMer"], fV, gT)
                ab = abFac * (gT[k - 1] - gNm1[k - 1])         # ADAMS_BASHFORTH2 on gT
                gNm1[k - 1] = gT[k - 1]
                gT[k - 1] = gT[k - 1] + ab
            theta[ti] = np.ascontiguousarray(th + dT[:, None, None] * gT)      # TIMESTEP_TRACER, CYCLE_TRACER
            gtNm1[ti] = gNm1
        # ---- DYNAMICS
        for bi, bj in tiles:
            fVerU, fVerV = np.zeros((2,) + ns), np.zeros((2,) + ns)
            phiHydF, phiHydC, dPx, dPy = (np.zeros(ns) for _ in range(4))
            for k in range(1, NR + 1):
                kUp, kDown = 1 + (k + 1) % 2, 1 + k % 2
                o.calc_phi_hyd(bi, bj, 0, d.sNx + 1, 0, d.sNy + 1, k, rhoInSitu, g.rF, g.rC, gravity,
                               recip_rhoConst, phi0surf, phiHydF, phiHydC, dPx, dPy)
                guDiss, gvDiss = np.zeros(ns), np.zeros(ns)
                e.mom_fluxform(bi, bj, k, 0, d.sNx + 1, 0, d.sNy + 1, kapU, kapU, fVerU[kUp - 1], fVerV[kUp - 1],
                               fVerU[kDown - 1], fVerV[kDown - 1], guDiss, gvDiss, uVel, vVel, wVel, gU, gV)
                o.timestep(bi, bj, k, 0, d.sNx + 1, 0, d.sNy + 1, dPx, dPy, guDiss, gvDiss, sfU, sfV,
                           1, 1, abFac, uVel, vVel, gU, gV, guNm1, gvNm1)
        # ---- SOLVE_FOR_PRESSURE
        b, x = np.zeros(d.shape2), np.zeros(d.shape2)
        for bi, bj in tiles:
            o.solve_rhs(bi, bj, etaN, gU, gV, b, x)
        res = e.cg2d(op, b, x, 1000, -1, sr=sr)
        o.exch_xyz(x)
        etaN = g.recip_Bo * x
        for bi, bj in tiles:
            o.correction_step(bi, bj, etaN, gU, gV, uVel, vVel)
            o.integrate_for_w(bi, bj, uVel, vVel, wVel)
        for a in (uVel, vVel, wVel, theta):
            o.exch_xyz(a, NR)
        rec = dict(res)
        rec.update(stats())
        out.append(rec)
    return op["cg2dNorm"], first, out
