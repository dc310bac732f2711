"""End-to-end oracle run of verification/tutorial_deep_convection: the NON-HYDROSTATIC step around CG3D.

TEST INFRASTRUCTURE ONLY.  100 x 100 x 50 cells of 20 m, doubly periodic, flat bottom, f plane, 2 x 2 tiles of
50 x 50 (code/SIZE.h), deltaT = 20 s, surface cooling Qnet in a disc, start state read from the experiment's own
T / U / V / Eta files (kept as tests/golden/inputs/deep_convection.npz by tests/golden/make_input_fixtures.py).
Order of model/src/forward_step.F with nonHydrostatic = T (non-staggered, exactConserv = F):

  EXTERNAL_FORCING_SURF  surfaceForcingT = -Qnet recip_Cp mass2rUnit              (external_forcing_surf.F:218-223)
  DO_OCEANIC_PHYS        FIND_RHO_2D (LINEAR)
  THERMODYNAMICS         TEMP_INTEGRATE: c2 advection + Laplacian diffusion + surface flux, AB2, explicit in the vertical
  DYNAMICS               CALC_PHI_HYD, MOM_FLUXFORM, TIMESTEP per level; then CALC_GW + TIMESTEP_WVEL (dynamics.F:638-655)
  SOLVE_FOR_PRESSURE     CALC_DIV_GHAT -> cg2d_b AND cg3d_b, old-style free-surface term with phi_nh(ks)
                         (oldFreeSurfTerm = use3Dsolver .AND. .NOT.exactConserv), CG2D with cg2dUseMinResSol = 1,
                         etaN = cg2d_x / Bo, PRE_CG3D, CG3D (cg3dMaxIters = 100 is reached on every step), EXCH(phi_nh)
  MOMENTUM_CORRECTION_STEP with the gradients of the surface pressure and of phi_nh
  INTEGR_CONTINUITY      INTEGRATE_FOR_W (the corrected flow is 3-D non-divergent to the solver's tolerance)
  DO_FIELDS_BLOCKING_EXCHANGES

The golden output (results/output.txt, 3 steps) prints for every step `cg2d: Sum(rhs),rhsMax`, cg2d_init_res,
cg2d_iters(min,last), cg2d_min_res / last_res and `cg3d: Sum(rhs),rhsMax`, cg3d_init_res, cg3d_last_res after 100
iterations, plus the monitor statistics: reproducing them PINS the CG3D restatement (cg3d_oracle.c) -- see
tests/test_oracle_golden.py -- and, with `engine`, the CUDA CG3D (never to claim oracle parity).
"""
from __future__ import annotations

import os

import numpy as np

from mitgcm_b200.grid import Dims, cartesian_grid, masks_from_depth, global_area
from .pyoracle import Oracle, Eos
from .barotropic_gyre import tile_field
from .baroclinic_gyre import mon_stats

NX = NY = 100
NR = 50
FIX = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "inputs", "deep_convection.npz")


def tile3(d, glob, o):
    """global (Nr, Ny, Nx) -> tiled (nSy, nSx, Nr, PY, PX) with halos (READ_FLD_XYZ_RL + _EXCH_XYZ_RL)."""
    a = np.zeros(d.shape3)
    for bj in range(d.nSy):
        for bi in range(d.nSx):
            a[bj, bi, :, d.OLy:d.OLy + d.sNy, d.OLx:d.OLx + d.sNx] = \
                glob[:, bj * d.sNy:(bj + 1) * d.sNy, bi * d.sNx:(bi + 1) * d.sNx]
    o.exch_xyz(a, d.Nr)
    return a


def setup(nSx=2, nSy=2):
    d = Dims(sNx=NX // nSx, sNy=NY // nSy, OLx=2, OLy=2, nSx=nSx, nSy=nSy, Nr=NR)
    g = cartesian_grid(d, [20.0] * NX, [20.0] * NY, [20.0] * NR, f0=1e-4, beta=0.0, gBaro=10.0)
    masks_from_depth(g, -1000.0 * np.ones((NY, NX)), hFacMin=1.0)
    P = dict(deltaTMom=20.0, deltaTFreeSurf=20.0, viscAhD=4e-2, viscAhZ=4e-2, no_slip_sides=0, no_slip_bottom=0,
             selectBotDragQuadr=-1, cg2dTargetResidual=1e-9, globalArea=global_area(g))
    return d, g, P


def run(nSteps=3, nSx=2, nSy=2, engine=None, want_state=False):
    """Returns (op2d, op3d, [per-step dict: cg2d scalars, cg3d scalars, monitor statistics])."""
    d, g, P = setup(nSx, nSy)
    o = Oracle(g, P)
    e = engine or o
    op = o.ini_cg2d()
    op3 = o.ini_cg3d(1.0, 1e-9, -1.0)
    if engine is not None and hasattr(engine, "setup"):
        engine.setup(g, o.params, op, op3=op3)
    fx = np.load(FIX)
    rhoConst = rhoNil = 1000.0
    gravity = 10.0
    recip_rhoConst = 1.0 / rhoConst
    mass2rUnit = recip_rhoConst
    recip_Cp = 1.0 / 4000.0
    eos = Eos(rhoNil, rhoConst, 2e-4, 0.0)
    tRef = np.full(NR, 20.0)
    sRef = np.full(NR, 35.0)
    abEps, viscAr, diffKhT, diffKrT = 0.1, 4e-2, 4e-2, 4e-2
    dT = np.full(NR, 20.0)
    zr = np.zeros(NR)
    ns = (d.PY, d.PX)
    tiles = [(bi, bj) for bj in range(1, nSy + 1) for bi in range(1, nSx + 1)]
    z3 = lambda: np.zeros(d.shape3)
    uVel = tile3(d, fx["U"].astype(np.float64), o) * g.maskW
    vVel = tile3(d, fx["V"].astype(np.float64), o) * g.maskS
    theta = tile3(d, fx["T"].astype(np.float64), o)
    etaN = tile_field(d, fx["Eta"].astype(np.float64))
    Qnet = tile_field(d, fx["Qnet"].astype(np.float64))
    salt = np.where(g.maskC != 0.0, 35.0, 0.0)
    wVel, gU, gV, gW, guNm1, gvNm1, gwNm1, gtNm1, rhoInSitu, ivdc, phi_nh = (z3() for _ in range(11))
    sfU, sfV = np.zeros(d.shape2), np.zeros(d.shape2)
    sfT = 0.0 - Qnet * recip_Cp * mass2rUnit                       # external_forcing_surf.F:218-223
    phi0surf = np.zeros(d.shape2)
    kapU = np.full((NR + 1,) + ns, viscAr)                         # calc_viscosity.F: viscArNr
    kapT = np.full(ns, diffKrT)
    # GRID.h depths of the flat-bottom column (ini_depths.F, ini_masks_etc.F: rLowW = max of the two neighbours ...)
    R_low = np.full(d.shape2, -1000.0)
    Ro_surf = np.zeros(d.shape2)
    rLowW, rLowS, rSurfW, rSurfS = R_low.copy(), R_low.copy(), Ro_surf.copy(), Ro_surf.copy()
    rC = np.ascontiguousarray(g.a["rC"][:NR])
    maskInC, maskInW, maskInS = g.maskC[:, :, 0], g.maskW[:, :, 0], g.maskS[:, :, 0]
    drF, drC = g.drF, g.drC[:NR]
    # INITIALISE_VARIA: INTEGR_CONTINUITY diagnoses the start w (initialise_varia.F:336)
    for bi, bj in tiles:
        o.integrate_for_w(bi, bj, uVel, vVel, wVel)
    o.exch_xyz(wVel, NR)

    def stats():
        return dict(eta=mon_stats(d, etaN[:, :, None], maskInC[:, :, None], maskInC, g.rA, [drF[0]]),
                    uvel=mon_stats(d, uVel, g.hFacW, maskInW, g.rAw, drF), vvel=mon_stats(d, vVel, g.hFacS, maskInS, g.rAs, drF),
                    wvel=mon_stats(d, wVel, g.maskC, maskInC, g.rA, drC), theta=mon_stats(d, theta, g.hFacC, maskInC, g.rA, drF))
    rec0 = stats()
    out = []
    for it in range(nSteps):
        abFac = 0.0 if it == 0 else 0.5 + abEps
        # ---- DO_OCEANIC_PHYS
        for bi, bj in tiles:
            o.density_ivdc(eos, bi, bj, theta, salt, tRef, sRef, rhoInSitu, ivdc)
        # ---- THERMODYNAMICS / TEMP_INTEGRATE
        for bi, bj in tiles:
            ti = (bj - 1, bi - 1)
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
                gtForc = np.zeros(ns)
                o.apply_forcing_T(bi, bj, k, sfT, gtForc)
                e.gad_calc_rhs(bi, bj, 0, d.sNx + 1, 0, d.sNy + 1, k, max(1, k - 1), kUp, kDown, sl["xA"],
                               sl["yA"], sl["maskUp"], sl["uFld"], sl["vFld"], sl["wFld"], sl["uTrans"],
                               sl["vTrans"], rTrans, sl["rTransKp1"], diffKhT, 0.0, kapT, zr, th, gNm1,
                               dT, 2, 2, 1, 0, 0, 0, sl["fZon"], sl["fMer"], fV, gT)
                gT[k - 1] = gT[k - 1] + gtForc                 # tracForcingOutAB = 0
                ab = abFac * (gT[k - 1] - gNm1[k - 1])         # ADAMS_BASHFORTH2 on gT
                gNm1[k - 1] = gT[k - 1]
                gT[k - 1] = gT[k - 1] + ab
            theta[ti] = th + dT[:, None, None] * gT            # TIMESTEP_TRACER + CYCLE_TRACER
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
            # non-hydrostatic: vertical momentum (dynamics.F:638-655)
            rc = o.calc_gw(bi, bj, R_low, Ro_surf, rLowW, rSurfW, rLowS, rSurfS, rC, kapU, kapU, 4e-2, 0.0, 1, abFac,
                           uVel, vVel, wVel, gW, gwNm1)
            assert rc == 0
            o.timestep_wvel(bi, bj, gW, wVel)
        # ---- SOLVE_FOR_PRESSURE
        b, x, b3 = np.zeros(d.shape2), np.zeros(d.shape2), z3()
        for bi, bj in tiles:
            o.solve_rhs_nh(bi, bj, etaN, phi_nh, gU, gV, b, x, b3)
        res = e.cg2d(op, b, x, 1000, 0)                        # nIterMin = cg2dUseMinResSol - 1 = 0
        o.exch_xyz(x)
        etaN = g.recip_Bo * x
        for bi, bj in tiles:
            o.pre_cg3d(bi, bj, x, etaN, wVel, b3)
        res3 = e.cg3d(op3, b3, phi_nh, 100)
        o.exch_xyz(phi_nh, NR)
        # ---- MOMENTUM_CORRECTION_STEP, INTEGR_CONTINUITY
        for bi, bj in tiles:
            o.correction_step_nh(bi, bj, etaN, phi_nh, gU, gV, uVel, vVel)
        for bi, bj in tiles:
            o.integrate_for_w(bi, bj, uVel, vVel, wVel)
        # ---- DO_FIELDS_BLOCKING_EXCHANGES
        for a in (uVel, vVel, wVel, theta):
            o.exch_xyz(a, NR)
        rec = dict(res)
        rec["cg3d"] = res3
        rec.update(stats())
        out.append(rec)
    if want_state:
        return op, op3, out, rec0, dict(uVel=uVel, vVel=vVel, wVel=wVel, theta=theta, etaN=etaN, phi_nh=phi_nh)
    return op, op3, out, rec0
