"""Parity of the CUDA MOM_VECINV (through the C ABI, reference argument list, host buffers) against the
CPU oracle, level by level in the caller's order (dynamics.F:422 k = 1..Nr), and the whole
solid-body.cs-32x32x1 experiment with the CUDA kernels in the loop against the reference's golden output.

Tolerance: every expression is evaluated in the reference's order without FMA contraction, so the results
are expected to be bit-identical; the operator tests allow 1e-13 relative (north_star: 1e-12)."""
import json
import os

import numpy as np
import pytest

from helpers import make_grid, CudaEngine
from oracle.pyoracle import Oracle

pytestmark = pytest.mark.gpu


@pytest.fixture()
def rt():
    from mitgcm_b200 import runtime
    yield runtime
    runtime.finalize()


def close(a, b, tol=1e-13):
    scale = max(np.abs(b).max(), 1e-300)
    return np.abs(a - b).max() <= tol * scale


VI_CASES = [
    dict(),                                                                   # defaults of the reference
    dict(selectVortScheme=0, selectKEscheme=1),
    dict(selectVortScheme=2, selectKEscheme=2, selectCoriScheme=1),
    dict(selectVortScheme=3, selectKEscheme=3, selectCoriScheme=2),
    dict(selectKEscheme=-1, selectCoriScheme=3, upwindShear=1),
    dict(useAbsVorticity=1),
    dict(useAbsVorticity=1, useJamartMomAdv=1, selectVortScheme=2),
    dict(useAbsVorticity=1, momAdvection=0),
    dict(momAdvection=0, useCoriolis=0),
    dict(useCDscheme=1),
    dict(viscAhD=400.0, viscAhZ=300.0),
    dict(viscAhD=400.0, viscAhZ=300.0, viscA4D=1e11, viscA4Z=2e11, useBiharmonicVisc=1, no_slip_sides=1),
    dict(viscA4D=1e11, viscA4Z=2e11, useBiharmonicVisc=1, corners=15, face=3),
    dict(viscAhD=400.0, viscAhZ=300.0, no_slip_sides=1, no_slip_bottom=1, bottomDragLinear=1e-3),
    dict(selectBotDragQuadr=0, bottomDragQuadratic=2e-3, bottomVisc_pCell=1, no_slip_bottom=1),
    dict(selectBotDragQuadr=2, bottomDragQuadratic=2e-3, implicitViscosity=1),
    dict(momViscosity=0, corners=15, face=1),
    dict(corners=15, face=2),
    dict(corners=6, face=4),
    dict(corners=9, face=6, viscAhD=400.0, viscAhZ=300.0),
    # MOM_VI_{U,V}_CORIOLIS_C4: 4th-order / upwind interpolation of the vorticity
    dict(highOrderVorticity=1, selectVortScheme=0),
    dict(highOrderVorticity=1, selectVortScheme=2, useAbsVorticity=1, corners=15, face=2),
    dict(upwindVorticity=1, selectVortScheme=2),
    dict(upwindVorticity=1, selectVortScheme=0, corners=5, face=1, useCoriolis=0),
]


@pytest.mark.parametrize("case", VI_CASES, ids=lambda c: "-".join(f"{k}{v}" for k, v in c.items()) or "default")
@pytest.mark.parametrize("shape", [dict(sNx=31, sNy=17, OL=3, nSx=2, nSy=2, Nr=5), dict(sNx=66, sNy=34, OL=3, Nr=4)],
                         ids=["tiles2x2", "1tile"])
def test_mom_vecinv_matches_oracle(rt, case, shape):
    case = dict(case)
    corners, face = case.pop("corners", 0), case.pop("face", 0)
    g = make_grid(**shape, seed=31)
    d = g.d
    rng = np.random.default_rng(2)
    g.a["cosFacU"] = 0.5 + 0.5 * rng.random((d.nSy, d.nSx, d.PY))
    g.a["cosFacV"] = 0.5 + 0.5 * rng.random((d.nSy, d.nSx, d.PY))
    params = dict(viscAhD=0.0, viscAhZ=0.0, no_slip_sides=0, no_slip_bottom=0, sideDragFactor=2.0)
    params.update(case)
    o = Oracle(g, params)
    u = 0.2 * rng.standard_normal(d.shape3) * g.maskW
    v = 0.2 * rng.standard_normal(d.shape3) * g.maskS
    w = 1e-4 * rng.standard_normal(d.shape3) * g.maskC
    rt.init(d)
    rt.set_grid(g)
    rt.set_params(**params)
    ns = (d.PY, d.PX)
    iMin, iMax, jMin, jMax = 0, d.sNx + 1, 0, d.sNy + 1
    gU_o, gV_o, gU_g, gV_g = (np.full(d.shape3, 7.0) for _ in range(4))     # points outside the range stay
    for bj in range(1, d.nSy + 1):
        for bi in range(1, d.nSx + 1):
            kap = 1e-3 * (1 + rng.random((d.Nr + 1,) + ns))
            kav = 1e-3 * (1 + rng.random((d.Nr + 1,) + ns))
            fU_o, fV_o, fU_g, fV_g = (np.zeros((2,) + ns) for _ in range(4))
            for k in range(1, d.Nr + 1):
                kUp, kDown = 1 + (k + 1) % 2, 1 + k % 2
                gd_o, hd_o, gd_g, hd_g = (np.zeros(ns) for _ in range(4))
                o.mom_vecinv(bi, bj, k, iMin, iMax, jMin, jMax, kap, kav, fU_o[kUp - 1], fV_o[kUp - 1],
                             fU_o[kDown - 1], fV_o[kDown - 1], gd_o, hd_o, u, v, w, gU_o, gV_o, corners, face)
                rt.mom_vecinv(bi, bj, k, iMin, iMax, jMin, jMax, kap, kav, fU_g[kUp - 1], fV_g[kUp - 1],
                              fU_g[kDown - 1], fV_g[kDown - 1], gd_g, hd_g, u, v, w, gU_g, gV_g, corners, face)
                t = (bj - 1, bi - 1, k - 1)
                assert close(fU_g[kDown - 1], fU_o[kDown - 1]) and close(fV_g[kDown - 1], fV_o[kDown - 1]), (k, "fVer kp")
                assert close(gU_g[t], gU_o[t]) and close(gV_g[t], gV_o[t]), (k, "gU/gV")
                assert close(gd_g, gd_o) and close(hd_g, hd_o), (k, "dissipation")
    if case.get("momAdvection", 1) or case.get("useCoriolis", 1):
        assert np.abs(gU_o - 7.0).max() > 0


def test_mom_vecinv_rejects_unsupported_options(rt):
    g = make_grid(8, 8, 2, Nr=2, seed=1)
    d = g.d
    rt.init(d)
    rt.set_grid(g)
    z = np.zeros((d.PY, d.PX))
    z3 = np.zeros(d.shape3)
    k3 = np.zeros((3, d.PY, d.PX))
    args = lambda: (k3, k3, z.copy(), z.copy(), z.copy(), z.copy(), z.copy(), z.copy(), z3, z3, z3, z3.copy(), z3.copy())
    with pytest.raises(rt.B200Error):          # range reaches the outermost halo ring
        rt.mom_vecinv(1, 1, 1, 1 - d.OLx, d.sNx + d.OLx, 0, 9, *args())
    rt.set_params(highOrderVorticity=1)
    with pytest.raises(rt.B200Error):          # MOM_VI_U_CORIOLIS_C4 has no selectVortScheme = 1 (the default)
        rt.mom_vecinv(1, 1, 1, 0, 9, 0, 9, *args())
    rt.set_params(highOrderVorticity=0, selectVortScheme=4)
    with pytest.raises(rt.B200Error):
        rt.mom_vecinv(1, 1, 1, 0, 9, 0, 9, *args())


def test_solid_body_cs_with_cuda_kernels_matches_golden(rt):
    """verification/solid-body.cs-32x32x1 with MOM_VECINV, GAD_CALC_RHS and CG2D (exch2 tile graph) on the
    GPU: iteration counts identical, solver scalars and monitor statistics of all 25 steps as the golden
    output prints them (13 digits; dot products are summed in a different order on the GPU: 1e-11)."""
    from oracle import solid_body_cs as sbc
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "solid-body.cs-32x32x1.json")))
    eng = CudaEngine(rt, fallback=None)
    norm, out = sbc.run(25, engine=eng)
    assert [r["numIters"] for r in out[1:]] == gold["cg2d_iters"]
    for r, ir, (sr, rm) in zip(out[1:], gold["cg2d_init_res"], gold["sumRHS_rhsMax"]):
        assert r["firstResidual"] == pytest.approx(float(ir), rel=1e-11)
        assert r["rhsMax"] == pytest.approx(float(rm), rel=1e-12)
    for fld in ("eta", "uvel", "vvel", "wvel", "salt"):
        for st in ("max", "min", "sd"):
            for r, gv in zip(out, gold[f"dynstat_{fld}_{st}"]):
                assert r[fld][st] == pytest.approx(float(gv), rel=1e-11, abs=1e-13), (fld, st)


@pytest.mark.parametrize("case", [VI_CASES[0], VI_CASES[5], VI_CASES[10], VI_CASES[13], VI_CASES[19]],
                         ids=["default", "absvort", "harmonic", "noslip-drag", "corners-harmonic"])
def test_mom_vecinv_staged_path_matches_oracle(rt, case, monkeypatch):
    """The slab-staged kernels (the only path with biharmonic viscosity) forced for cases the one-launch
    re-evaluating kernel normally serves: both must agree with the oracle."""
    monkeypatch.setenv("MITGCM_B200_VI_STAGED", "1")
    test_mom_vecinv_matches_oracle(rt, case, dict(sNx=31, sNy=17, OL=3, nSx=2, nSy=2, Nr=5))


def test_resident_solid_body_cs_step_matches_the_golden():
    """solid-body.cs-32x32x1 stepped ENTIRELY on the device (mitgcm_b200_forward_step_ with
    MI_VECTORINVARIANTMOMENTUM on the exch2 tile graph, p coordinates: rkSign = -1, Bo_surf = 1/rhoConst).
    The passive tracer of the experiment is salt (SALT_INTEGRATE: centred advection, Adams-Bashforth on the tendency);
    theta is not stepped, as in the experiment.  25 steps against the golden output."""
    from mitgcm_b200.model import Model, ini_cg2d_tilegraph
    from oracle.baroclinic_gyre import mon_stats
    from oracle import solid_body_cs as sbc
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "solid-body.cs-32x32x1.json")))
    T, d, g, P, salt = sbc.setup()
    P = dict(P)
    P.update(abEps=0.1, deltaTtracer=450.0, viscAr=0.0, tempStepping=0, saltStepping=1, saltAdvScheme=2, cg2dMaxIters=600,
             momForcing=1, momDissip_In_AB=1, diffKhT=0.0, diffK4T=0.0, diffKrT=0.0, diffKhS=0.0, diffK4S=0.0, diffKrS=0.0,
             vectorInvariantMomentum=1)
    op = ini_cg2d_tilegraph(g, P, T)
    uVel, vVel, etaN = sbc.initial_state(T, d, g)
    sbc.eo.exch2_3d(T, salt[0], d.OLx)
    o = Oracle(g, P)
    wVel = np.zeros(d.shape3)
    for bi in range(1, d.nSx + 1):
        o.integrate_for_w(bi, 1, uVel, vVel, wVel)
    sbc.eo.exch2_3d(T, wVel[0], d.OLx)
    state = dict(uVel=uVel, vVel=vVel, wVel=wVel, theta=np.zeros(d.shape3), salt=salt, etaN=etaN, surfForcU=np.zeros(d.shape2),
                 surfForcV=np.zeros(d.shape2))
    m = Model(g, P, state, op, device=0, topo=T)
    maskInC, maskInW, maskInS = g.maskC[:, :, 0], g.maskW[:, :, 0], g.maskS[:, :, 0]
    try:
        for it in range(25):
            r = m.step()
            assert r["numIters"] == gold["cg2d_iters"][it], it
            assert r["firstResidual"] == pytest.approx(float(gold["cg2d_init_res"][it]), rel=1e-10), it
            st = dict(eta=mon_stats(d, m.get("etaN")[:, :, None], maskInC[:, :, None], maskInC, g.rA, [g.drF[0]]),
                      uvel=mon_stats(d, m.get("uVel"), g.hFacW, maskInW, g.rAw, g.drF),
                      vvel=mon_stats(d, m.get("vVel"), g.hFacS, maskInS, g.rAs, g.drF),
                      wvel=mon_stats(d, m.get("wVel"), g.maskC, maskInC, g.rA, g.drC[:1]),
                      salt=mon_stats(d, m.get("salt"), g.hFacC, maskInC, g.rA, g.drF))
            for f in ("eta", "uvel", "vvel", "wvel", "salt"):
                for s in ("max", "min", "sd"):
                    assert st[f][s] == pytest.approx(float(gold[f"dynstat_{f}_{s}"][it + 1]), rel=1e-10, abs=1e-13), (it, f, s)
    finally:
        m.close()
