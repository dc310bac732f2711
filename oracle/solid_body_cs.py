"""End-to-end oracle run of verification/solid-body.cs-32x32x1: solid-body rotation of a one-layer
atmosphere (p coordinates, Nr = 1) on the cs32 cubed sphere, 6 tiles of 32x32, OL = 2, pkg/exch2,
with a passive tracer (salt) advected by the flow.

TEST INFRASTRUCTURE ONLY.  Pins MOM_VECINV (pkg/mom_vecinv) against
the experiment's golden output (results/output.txt: `CG2D normalisation factor`, 25 steps of
`cg2d: Sum(rhs),rhsMax`, cg2d_init_res, iteration counts and %MON dynstat_{eta,uvel,vvel,wvel,salt}_*):
  * vectorInvariantMomentum with the defaults selectVortScheme = 1, selectKEscheme = 0,
    selectCoriScheme = 0, useAbsVorticity = F: MOM_CALC_RELVORT3 with its three-cell facet corners,
    MOM_CALC_KE, MOM_VI_CORIOLIS, MOM_VI_{U,V}_CORIOLIS, MOM_VI_{U,V}_GRAD_KE, MOM_VI_{U,V}_VERTSHEAR
    (zero at Nr = 1), momViscosity = T with all coefficients zero (MOM_CALC_HDIV, MOM_VI_HDISSIP,
    MOM_{U,V}_RVISCFLUX executed);
  * GAD_CALC_RHS (centred 2nd order) + Adams-Bashforth on the tendency for salt on the cube.
Run-time switches from input/data and the golden's parameter summary: deltaT = 450, abEps = 0.1,
rotationPeriod = 108000, rSphere = 5500.4e3 with radius_fromHorizGrid = 6370e3, delR = 1e5 Pa,
rhoConst = 1 (uniformLin_PhiSurf: Bo_surf = 1/rhoConst), rkSign = -1 (p coordinates),
implicSurfPress = implicDiv2DFlow = 1, exactConserv = F, cg2dTargetResidual = 1e-12.
The experiment's own INI_VEL / INI_PSURF (code/ini_vel.F:38-75, code/ini_psurf.F:40-60) are restated in
`initial_state`.  theta = tRef is uniform and not stepped, so the hydrostatic pressure gradient is zero
to the bit and CALC_PHI_HYD is not called.

Sequence per step (forward_step.F, non-staggered): THERMODYNAMICS (SALT_INTEGRATE), DYNAMICS
(MOM_VECINV, TIMESTEP), SOLVE_FOR_PRESSURE, MOMENTUM_CORRECTION_STEP, INTEGR_CONTINUITY,
DO_FIELDS_BLOCKING_EXCHANGES.  `engine` as in adjustment_cs.py (mom_vecinv / gad_calc_rhs / cg2d from
the CUDA library).
"""
from __future__ import annotations

import os

import numpy as np

from mitgcm_b200.grid import Dims, cubed_sphere_grid, cs_corner_flags, set_hfac, global_area
from mitgcm_b200.exch2 import cubed_sphere_topology
from mitgcm_b200.model import ini_cg2d_tilegraph
from . import exch2_oracle as eo
from .pyoracle import Oracle
from .baroclinic_gyre import mon_stats

FIXTURE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "inputs",
                       "solid_body_cs32.npz")
KEEP = "xC yC dxF dyF rA xG yG dxV dyU rAz dxC dyC rAw rAs dxG dyG".split()
RSPHERE = 5500.4e3


def setup():
    z = np.load(FIXTURE)
    faces = [{n: z[f"{n}_{f}"] for n in KEEP} for f in range(6)]
    T = cubed_sphere_topology(32, 32, 32)
    d = Dims(sNx=32, sNy=32, OLx=2, OLy=2, nSx=6, nSy=1, Nr=1)
    g = cubed_sphere_grid(d, T, faces, [1.0e5], rotationPeriod=108000.0, rSphereFac=RSPHERE / 6370.0e3,
                          Bo_surf=1.0)
    one = np.ones(d.shape3)                    # no topography: hFac = 1 everywhere, halo included
    set_hfac(g, one, one.copy(), one.copy())
    P = dict(deltaTMom=450.0, deltaTFreeSurf=450.0, implicSurfPress=1.0, implicDiv2DFlow=1.0, rkSign=-1.0,
             cg2dTargetResidual=1e-12, momAdvection=1, momViscosity=1, no_slip_sides=0, no_slip_bottom=0,
             viscAhD=0.0, viscAhZ=0.0, selectBotDragQuadr=-1, selectKEscheme=0, selectVortScheme=1,
             selectCoriScheme=0, useAbsVorticity=0, useCoriolis=1, globalArea=global_area(g))
    # W2_mapIO = 1: facets stacked along y in the file
    salt0 = np.zeros(d.shape3)
    for t in range(6):
        salt0[0, t, 0, d.OLy:d.OLy + 32, d.OLx:d.OLx + 32] = z["S_init"][t * 32:(t + 1) * 32, :]
    return T, d, g, P, salt0


def initial_state(T, d, g):
    """code/ini_vel.F and code/ini_psurf.F of the experiment."""
    OL = d.OLx
    omega = g.a["omega"]
    omegaprime = 80.0 / RSPHERE
    fac = -(RSPHERE * RSPHERE) * omegaprime / (2.0 * omega)
    psi = fac * g.a["fCoriG"]
    ip1 = np.minimum(np.arange(d.PX) + 1, d.PX - 1)
    jp1 = np.minimum(np.arange(d.PY) + 1, d.PY - 1)
    u2 = 0.0 + (psi - psi[:, :, jp1, :]) * g.a["recip_dyG"]
    v2 = 0.0 + (psi[:, :, :, ip1] - psi) * g.a["recip_dxG"]
    uVel, vVel = np.ascontiguousarray(u2[:, :, None]), np.ascontiguousarray(v2[:, :, None])
    eo.exch2_uv_3d(T, uVel[0], vVel[0], OL, True)
    uVel *= g.maskW
    vVel *= g.maskS
    psFac = -(RSPHERE * RSPHERE) * omegaprime * (omega + omegaprime * 0.5)
    snFac = 1.0 / (4.0 * omega * omega)
    etaN = 0.0 + psFac * (snFac * g.a["fCori"] * g.a["fCori"] - 1.0 / 3.0) * g.a["recip_Bo"]
    return uVel, vVel, np.ascontiguousarray(etaN)


def run(nSteps=25, engine=None, want_state=False):
    """Returns (cg2dNorm, [per-step dict of solver scalars and monitor statistics])."""
    T, d, g, P, salt = setup()
    OL = d.OLx
    o = Oracle(g, P)
    hook = eo.Exch2Hook(o, T, OL)                 # EXCH2 inside the oracle's CG2D
    e = engine or o
    op = ini_cg2d_tilegraph(g, P, T)
    if engine is not None and hasattr(engine, "setup"):
        engine.setup(g, o.params, op, T)
    if engine is not None and getattr(engine, "fb", 0) is None:
        engine.fb = o
    corners = cs_corner_flags(T)
    abEps = 0.1
    ns = (d.PY, d.PX)
    tiles = [(bi, 1) for bi in range(1, d.nSx + 1)]
    z3 = lambda: np.zeros(d.shape3)
    wVel, gU, gV, guNm1, gvNm1, gsNm1 = (z3() for _ in range(6))
    uVel, vVel, etaN = initial_state(T, d, g)
    eo.exch2_3d(T, salt[0], OL)                    # ini_fields.F: _EXCH_XYZ_RL(salt)
    # INTEGR_CONTINUITY at start-up (initialise_varia.F:240-260): wVel of the initial flow
    for bi, bj in tiles:
        o.integrate_for_w(bi, bj, uVel, vVel, wVel)
    eo.exch2_3d(T, wVel[0], OL)
    kap = np.zeros((d.Nr + 1,) + ns)
    sfU = np.zeros(d.shape2)
    zero = np.zeros(ns)
    dT = np.full(d.Nr, 450.0)
    zr = np.zeros(d.Nr)
    kapS = np.zeros(ns)
    maskInC, maskInW, maskInS = g.maskC[:, :, 0], g.maskW[:, :, 0], g.maskS[:, :, 0]

    def stats(res):
        rec = dict(res)
        rec["eta"] = mon_stats(d, etaN[:, :, None], maskInC[:, :, None], maskInC, g.rA, [g.drF[0]])
        rec["uvel"] = mon_stats(d, uVel, g.hFacW, maskInW, g.rAw, g.drF)
        rec["vvel"] = mon_stats(d, vVel, g.hFacS, maskInS, g.rAs, g.drF)
        rec["wvel"] = mon_stats(d, wVel, g.maskC, maskInC, g.rA, g.drC[:1])
        rec["salt"] = mon_stats(d, salt, g.hFacC, maskInC, g.rA, g.drF)
        return rec

    out = [stats({})]
    try:
        for it in range(nSteps):
            abFac = 0.0 if it == 0 else 0.5 + abEps
            # ---- THERMODYNAMICS: SALT_INTEGRATE (salt_integrate.F: CALC_ADV_FLOW, GAD_CALC_RHS, AB2, TIMESTEP_TRACER)
            for bi, bj in tiles:
                ti = (bj - 1, bi - 1)
                gS = np.zeros((d.Nr,) + ns)
                fV = np.zeros((2,) + ns)
                rTrans = np.zeros(ns)
                sl = {n: np.zeros(ns) for n in "xA yA maskUp uFld vFld wFld uTrans vTrans rTransKp1 fZon fMer".split()}
                S = np.ascontiguousarray(salt[ti])
                for k in range(d.Nr, 0, -1):
                    kUp, kDown = 1 + (k + 1) % 2, 1 + k % 2
                    o.calc_adv_flow(bi, bj, k, uVel, vVel, wVel, sl["xA"], sl["yA"], sl["maskUp"], sl["uFld"],
                                    sl["vFld"], sl["wFld"], sl["uTrans"], sl["vTrans"], rTrans, sl["rTransKp1"])
                    e.gad_calc_rhs(bi, bj, 0, d.sNx + 1, 0, d.sNy + 1, k, max(1, k - 1), kUp, kDown, sl["xA"], sl["yA"],
                                   sl["maskUp"], sl["uFld"], sl["vFld"], sl["wFld"], sl["uTrans"], sl["vTrans"], rTrans,
                                   sl["rTransKp1"], 0.0, 0.0, kapS, zr, S, S, dT, 2, 2, 1, 0, 0, 0, sl["fZon"],
                                   sl["fMer"], fV, gS)
                    gNm1 = gsNm1[ti][k - 1]
                    ab = abFac * (gS[k - 1] - gNm1)
                    gNm1[...] = gS[k - 1]
                    gS[k - 1] = gS[k - 1] + ab
                salt[ti] = S + dT[:, None, None] * gS
            # ---- DYNAMICS
            for bi, bj in tiles:
                fVerU, fVerV = np.zeros((2,) + ns), np.zeros((2,) + ns)
                guDiss, gvDiss = np.zeros(ns), np.zeros(ns)
                e.mom_vecinv(bi, bj, 1, 0, d.sNx + 1, 0, d.sNy + 1, kap, kap, fVerU[1], fVerV[1], fVerU[0], fVerV[0],
                             guDiss, gvDiss, uVel, vVel, wVel, gU, gV, int(corners[bi - 1]), int(T.myFace[bi - 1]))
                o.timestep(bi, bj, 1, 0, d.sNx + 1, 0, d.sNy + 1, zero, zero, guDiss, gvDiss, sfU, sfU, 1, 1, abFac,
                           uVel, vVel, gU, gV, guNm1, gvNm1)
            # ---- SOLVE_FOR_PRESSURE
            b, x = np.zeros(d.shape2), np.zeros(d.shape2)
            for bi, bj in tiles:
                o.solve_rhs(bi, bj, etaN, gU, gV, b, x)
            res = e.cg2d(op, b, x, 600, -1)
            eo.exch2_3d(T, x[0][:, None], OL)
            etaN = g.recip_Bo * x
            # ---- MOMENTUM_CORRECTION_STEP, INTEGR_CONTINUITY
            for bi, bj in tiles:
                o.correction_step(bi, bj, etaN, gU, gV, uVel, vVel)
                o.integrate_for_w(bi, bj, uVel, vVel, wVel)
            # ---- DO_FIELDS_BLOCKING_EXCHANGES
            eo.exch2_uv_3d(T, uVel[0], vVel[0], OL, True)
            eo.exch2_3d(T, wVel[0], OL)
            eo.exch2_3d(T, salt[0], OL)
            out.append(stats(res))
    finally:
        hook.close()
    if want_state:
        return op["cg2dNorm"], out, dict(uVel=uVel, vVel=vVel, wVel=wVel, etaN=etaN, salt=salt)
    return op["cg2dNorm"], out
