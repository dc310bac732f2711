"""End-to-end oracle run of verification/tutorial_baroclinic_gyre (config 2).

TEST INFRASTRUCTURE ONLY.  Drives the C restatements in the order of
model/src/forward_step.F (non-staggered, exactConserv) for the 62x62x15
spherical-polar gyre of verification/tutorial_baroclinic_gyre (input/data,
code/SIZE.h: 2x2 tiles of 31x31, OL = 2):

  DO_OCEANIC_PHYS   EXTERNAL_FORCING_SURF (SST relaxation, wind), FIND_RHO_2D, CALC_IVDC
  THERMODYNAMICS    TEMP_INTEGRATE: CALC_3D_DIFFUSIVITY, CALC_ADV_FLOW, APPLY_FORCING_T,
                    GAD_CALC_RHS, ADAMS_BASHFORTH2 (on gT), TIMESTEP_TRACER, GAD_IMPLICIT_R,
                    CYCLE_TRACER                              (temp_integrate.F:230-560)
  DYNAMICS          CALC_PHI_HYD, MOM_FLUXFORM, TIMESTEP      (dynamics.F:420-560)
  SOLVE_FOR_PRESSURE (exactConserv: etaH in the free-surface term), CG2D
  MOMENTUM_CORRECTION_STEP, INTEGR_CONTINUITY (dEtaHdt, etaN, wVel, etaH)
  DO_FIELDS_BLOCKING_EXCHANGES

so that GAD_CALC_RHS and the Nr > 1 branches of MOM_FLUXFORM (vertical advection and
viscosity, metric terms) are PINNED against the experiment's golden output
(results/output.txt: cg2dNorm, per-step cg2d_init_res / iters / last_res / rhsMax and
%MON dynstat_{eta,uvel,vvel,wvel,theta}_*), see tests/test_oracle_golden.py.

`engine`, if given, supplies mom_fluxform / gad_calc_rhs / cg2d with the Oracle's method
signatures, so a GPU test can put the CUDA kernels in the loop (never to claim oracle parity).
"""
from __future__ import annotations

import numpy as np

from mitgcm_b200.grid import Dims, spherical_polar_grid, masks_from_depth, exch_xyz, global_area
from .pyoracle import Oracle, Eos
from .barotropic_gyre import tile_field

NX = NY = 62
NR = 15
DELR = [50., 60., 70., 80., 90., 100., 110., 120., 130., 140., 150., 160., 170., 180., 190.]
TREF = [30., 27., 24., 21., 18., 15., 13., 11., 9., 7., 6., 5., 4., 3., 2.]


def gen_inputs():
    """input/gendata.py of the experiment: bathymetry, zonal wind stress, restoring SST (float32)."""
    Ho, nx, ny, xo, yo, dx, dy = 1800, NX, NY, 0, 15, 1, 1
    xeast = xo + (nx - 2) * dx
    ynorth = yo + (ny - 2) * dy
    h = -Ho * np.ones((ny, nx))
    h[:, [0, -1]] = 0
    h[[0, -1], :] = 0
    x = np.linspace(xo - dx, xeast, nx)
    y = np.linspace(yo - dy, ynorth, ny) + dy / 2
    Y, _ = np.meshgrid(y, x, indexing='ij')
    tau = -0.1 * np.cos(2 * np.pi * ((Y - yo) / (ny - 2) / dy))
    Trest = (30 - 0) / (ny - 2) / dy * (ynorth - Y) + 0
    return h.astype('>f4'), tau.astype('>f4'), Trest.astype('>f4')


def mon_stats(d, arr, hfac, mask2, area, dr):
    """pkg/monitor/mon_calc_stats_rl.F: min, max, volume-weighted mean and sd with the
    reference's summation order (k, j, i inside a tile, tiles bi-fast)."""
    jj, ii = d.interior()
    a = arr[:, :, :, jj, ii]
    m = mask2[:, :, None, jj, ii] * hfac[:, :, :, jj, ii]
    vol = area[:, :, None, jj, ii] * np.asarray(dr)[None, None, :, None, None] * m
    sel = m > 0
    mn, mx = a[sel].min(), a[sel].max()

    def tsum(x):
        tot = 0.0
        for bj in range(d.nSy):
            for bi in range(d.nSx):
                s = 0.0
                for v in x[bj, bi][sel[bj, bi]]:
                    s = s + v
                tot = tot + s
        return tot
    theVol = tsum(vol)
    mean = tsum(vol * a) / theVol
    sd = np.sqrt(tsum(vol * (a - mean) * (a - mean)) / theVol)
    return dict(min=mn, max=mx, mean=mean, sd=sd)


def setup(nSx=2, nSy=2):
    d = Dims(sNx=NX // nSx, sNy=NY // nSy, OLx=2, OLy=2, nSx=nSx, nSy=nSy, Nr=NR)
    g = spherical_polar_grid(d, [1.0] * NX, [1.0] * NY, DELR, xgOrigin=-1.0, ygOrigin=14.0,
                             rSphere=6370e3, rotationPeriod=86164.0, gBaro=9.81)
    bathy, wind, sst = gen_inputs()
    masks_from_depth(g, bathy.astype(np.float64), hFacMin=1.0, hFacMinDr=0.0)
    P = dict(deltaTMom=1200.0, deltaTFreeSurf=1200.0, viscAhD=5000.0, viscAhZ=5000.0,
             no_slip_sides=1, sideDragFactor=2.0, no_slip_bottom=0, selectBotDragQuadr=-1,
             usingSphericalPolarGrid=1, selectMetricTerms=1, recip_rSphere=1.0 / 6370e3,
             implicitDiffusion=1, cg2dTargetResidual=1e-7, globalArea=global_area(g))
    return d, g, P, wind, sst


def run(nSteps=10, nSx=2, nSy=2, engine=None, want_state=False):
    """Returns (cg2dNorm, [per-step dict of solver scalars and monitor statistics])."""
    d, g, P, wind, sst = setup(nSx, nSy)
    o = Oracle(g, P)
    e = engine or o
    op = o.ini_cg2d()
    if engine is not None and hasattr(engine, "setup"):
        engine.setup(g, o.params, op)
    rhoConst = rhoNil = 999.8
    gravity = 9.81
    recip_rhoConst = 1.0 / rhoConst
    mass2rUnit = recip_rhoConst
    recip_Cp = 1.0 / 3994.0
    eos = Eos(rhoNil, rhoConst, 2e-4, 0.0)
    tRef = np.array(TREF)
    sRef = np.full(NR, 30.0)
    abEps, viscAr, diffKhT, diffKrT, ivdc_kappa = 0.01, 1e-2, 1000.0, 1e-5, 1.0
    dT = np.full(NR, 1200.0)
    zr = np.zeros(NR)
    ns = (d.PY, d.PX)
    tiles = [(bi, bj) for bj in range(1, nSy + 1) for bi in range(1, nSx + 1)]

    fu = tile_field(d, wind.astype(np.float64))
    SST = tile_field(d, sst.astype(np.float64))
    sfU = fu * mass2rUnit                                     # external_forcing_surf.F:214
    sfV = np.zeros(d.shape2)
    sfT = np.zeros(d.shape2)
    lam = np.where(np.abs(g.yC) <= 180.0, 1.0 / 2592000.0, 0.0)   # ini_forcing.F:46-51
    phi0surf = np.zeros(d.shape2)
    z3 = lambda: np.zeros(d.shape3)
    uVel, vVel, wVel, gU, gV, guNm1, gvNm1, gtNm1, rhoInSitu, ivdc = (z3() for _ in range(10))
    theta = np.where(g.maskC != 0.0, tRef[None, None, :, None, None], 0.0)   # ini_theta.F
    salt = np.where(g.maskC != 0.0, 30.0, 0.0)
    etaN, etaH, dEtaHdt = (np.zeros(d.shape2) for _ in range(3))
    kapU = np.full((NR + 1,) + ns, viscAr)                     # calc_viscosity.F: viscArNr
    kbl = np.zeros(NR)                                         # diffKrBL79surf = deep = 0
    dkr = np.full(NR, diffKrT)
    maskInC, maskInW, maskInS = g.maskC[:, :, 0], g.maskW[:, :, 0], g.maskS[:, :, 0]
    drF, drC = g.drF, g.drC[:NR]
    rec0 = dict(theta=mon_stats(d, theta, g.hFacC, maskInC, g.rA, drF))
    out = []

    for it in range(nSteps):
        abFac = 0.0 if it == 0 else 0.5 + abEps
        # ---- DO_OCEANIC_PHYS
        for bi, bj in tiles:
            o.forcing_surf_relax_T(bi, bj, theta, SST, lam, recip_Cp, mass2rUnit, sfT)
            o.density_ivdc(eos, bi, bj, theta, salt, tRef, sRef, rhoInSitu, ivdc)
        # ---- THERMODYNAMICS / TEMP_INTEGRATE
        for bi, bj in tiles:
            ti = (bj - 1, bi - 1)
            kappaRk = np.zeros((NR,) + ns)
            o.calc_3d_diffusivity(bi, bj, ivdc, ivdc_kappa, kbl, dkr, kappaRk)
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
                               sl["vTrans"], rTrans, sl["rTransKp1"], diffKhT, 0.0, kappaRk[k - 1], zr, th, gNm1,
                               dT, 2, 2, 1, 0, 0, 0, sl["fZon"], sl["fMer"], fV, gT)
                gT[k - 1] = gT[k - 1] + gtForc                 # tracForcingOutAB = 0
                ab = abFac * (gT[k - 1] - gNm1[k - 1])         # ADAMS_BASHFORTH2 on gT
                gNm1[k - 1] = gT[k - 1]
                gT[k - 1] = gT[k - 1] + ab
            gT = th + dT[:, None, None] * gT                   # TIMESTEP_TRACER
            gT = np.ascontiguousarray(gT)
            err = o.gad_implicit_r(bi, bj, 0, d.sNx + 1, 0, d.sNy + 1, dT, kappaRk,
                                   np.ascontiguousarray(g.recip_hFacC[ti]), gT)
            assert err == 0
            theta[ti] = gT                                     # CYCLE_TRACER
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
        # ---- SOLVE_FOR_PRESSURE (exactConserv: etaH in the free-surface term)
        b, x = np.zeros(d.shape2), np.zeros(d.shape2)
        for bi, bj in tiles:
            o.solve_rhs(bi, bj, etaN, gU, gV, b, x, etaH=etaH)
        res = e.cg2d(op, b, x, 1000, -1)
        o.exch_xyz(x)
        etaN = g.recip_Bo * x
        # ---- MOMENTUM_CORRECTION_STEP, INTEGR_CONTINUITY
        for bi, bj in tiles:
            o.correction_step(bi, bj, etaN, gU, gV, uVel, vVel)
        for bi, bj in tiles:
            o.integr_continuity_ec(bi, bj, uVel, vVel, etaH, dEtaHdt, etaN, True)
            o.integrate_for_w(bi, bj, uVel, vVel, wVel)
        o.exch_xyz(etaN)
        etaH = etaN.copy()                                     # UPDATE_ETAH, implicDiv2DFlow = 1
        # ---- DO_FIELDS_BLOCKING_EXCHANGES
        for a in (uVel, vVel, wVel, theta):
            o.exch_xyz(a, NR)
        rec = dict(res)
        rec["eta"] = mon_stats(d, etaN[:, :, None], maskInC[:, :, None], maskInC, g.rA, [drF[0]])
        rec["uvel"] = mon_stats(d, uVel, g.hFacW, maskInW, g.rAw, drF)
        rec["vvel"] = mon_stats(d, vVel, g.hFacS, maskInS, g.rAs, drF)
        rec["wvel"] = mon_stats(d, wVel, g.maskC, maskInC, g.rA, drC)
        rec["theta"] = mon_stats(d, theta, g.hFacC, maskInC, g.rA, drF)
        out.append(rec)
    if want_state:
        return op, out, rec0, dict(uVel=uVel, vVel=vVel, wVel=wVel, theta=theta, etaN=etaN)
    return op["cg2dNorm"], out
