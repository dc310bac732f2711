"""Parity of the CUDA GAD_CALC_RHS and MOM_FLUXFORM (through the C ABI, reference argument lists,
host buffers) against the CPU oracle, level by level in the callers' marching order
(temp_integrate.F:296 k = Nr..1; dynamics.F:422 k = 1..Nr).

Tolerance: every expression is evaluated in the reference's order without FMA contraction, so
the results are expected to be bit-identical; the tests allow 1e-13 relative (north_star: 1e-12)."""
import numpy as np
import pytest

from helpers import make_grid
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


def rand_state(g, seed):
    d = g.d
    rng = np.random.default_rng(seed)
    u = 0.2 * rng.standard_normal(d.shape3) * g.maskW
    v = 0.2 * rng.standard_normal(d.shape3) * g.maskS
    w = 1e-4 * rng.standard_normal(d.shape3) * g.maskC
    T = (10.0 + rng.standard_normal(d.shape3)) * g.maskC
    return u, v, w, T


GAD_CASES = [
    dict(scheme=2, diffKh=1e3, diffK4=0.0, ab=False),
    dict(scheme=2, diffKh=1e3, diffK4=1e11, ab=True),
    dict(scheme=1, diffKh=0.0, diffK4=0.0, ab=False),
    dict(scheme=20, diffKh=5e2, diffK4=0.0, ab=False),
    dict(scheme=3, diffKh=1e3, diffK4=0.0, ab=True),
    dict(scheme=4, diffKh=1e3, diffK4=0.0, ab=True),
    dict(scheme=30, diffKh=0.0, diffK4=0.0, ab=False),
    dict(scheme=33, diffKh=1e2, diffK4=0.0, ab=False),
    dict(scheme=77, diffKh=1e2, diffK4=0.0, ab=False, kr4=True),
    dict(scheme=2, diffKh=1e3, diffK4=0.0, ab=False, implDiff=True, implAdv=True),
    dict(scheme=7, diffKh=1e2, diffK4=0.0, ab=False),
    dict(scheme=7, diffKh=0.0, diffK4=0.0, ab=True),
    # calcAdvection = F (the call temp_integrate.F makes after GAD_ADVECTION): diffusion on top of a given tendency
    dict(scheme=33, diffKh=1e3, diffK4=1e11, ab=False, calcAdv=0, gT0=1e-6),
]


@pytest.mark.parametrize("case", GAD_CASES, ids=lambda c: "-".join(f"{k}{v}" for k, v in c.items()))
@pytest.mark.parametrize("shape", [dict(sNx=31, sNy=17, OL=4, nSx=2, nSy=2, Nr=6), dict(sNx=70, sNy=40, OL=4, Nr=5)],
                         ids=["tiles2x2", "1tile"])
def test_gad_calc_rhs_matches_oracle(rt, case, shape):
    g = make_grid(**shape, seed=21)
    d = g.d
    o = Oracle(g, dict(implicitDiffusion=int(case.get("implDiff", False))))
    u, v, w, T = rand_state(g, 5)
    rng = np.random.default_rng(9)
    TAB = T + 0.01 * rng.standard_normal(d.shape3)
    rt.init(d)
    rt.set_grid(g)
    rt.set_params(implicitDiffusion=int(case.get("implDiff", False)))
    ns = (d.PY, d.PX)
    dT = np.full(d.Nr, 1200.0)
    kr4 = np.full(d.Nr, 1e-3 if case.get("kr4") else 0.0)
    for bj in range(1, d.nSy + 1):
        for bi in range(1, d.nSx + 1):
            t = (bj - 1, bi - 1)
            gT_o = case.get("gT0", 0.0) * rng.standard_normal((d.Nr,) + ns)
            gT_g = gT_o.copy()
            fV_o, fV_g = np.zeros((2,) + ns), np.zeros((2,) + ns)
            rTrans = np.zeros(ns)
            for k in range(d.Nr, 0, -1):
                kUp, kDown = 1 + (k + 1) % 2, 1 + k % 2
                sl = {n: np.zeros(ns) for n in "xA yA maskUp uFld vFld wFld uTrans vTrans rTransKp1".split()}
                o.calc_adv_flow(bi, bj, k, u, v, w, sl["xA"], sl["yA"], sl["maskUp"], sl["uFld"], sl["vFld"],
                                sl["wFld"], sl["uTrans"], sl["vTrans"], rTrans, sl["rTransKp1"])
                KappaR = 1e-4 * (1 + rng.random(ns))
                args = (bi, bj, 1, d.sNx, 1, d.sNy, k, max(1, k - 1), kUp, kDown, sl["xA"], sl["yA"], sl["maskUp"],
                        sl["uFld"], sl["vFld"], sl["wFld"], sl["uTrans"], sl["vTrans"], rTrans, sl["rTransKp1"],
                        case["diffKh"], case["diffK4"], KappaR, kr4, np.ascontiguousarray(T[t]),
                        np.ascontiguousarray(TAB[t]), dT)
                fZo, fMo, fZg, fMg = (np.zeros(ns) for _ in range(4))
                o.gad_calc_rhs(*args, case["scheme"], case["scheme"], case.get("calcAdv", 1), int(case.get("implAdv", False)),
                               int(case["ab"]), int(bool(case.get("kr4"))), fZo, fMo, fV_o, gT_o)
                rt.gad_calc_rhs(*args, 1, case["scheme"], case["scheme"], case.get("calcAdv", 1), int(case.get("implAdv", False)),
                                int(case["ab"]), int(bool(case.get("kr4"))), 0, 0, 0, fZg, fMg, fV_g, gT_g)
                assert close(fZg, fZo) and close(fMg, fMo), (k, "horizontal fluxes")
                assert close(fV_g[kUp - 1], fV_o[kUp - 1]), (k, "vertical flux")
                assert close(gT_g[k - 1, :-1, :-1], gT_o[k - 1, :-1, :-1]), (k, "tendency")
            assert np.abs(gT_o).max() > 0


def test_gad_rejects_unsupported_options(rt):
    g = make_grid(8, 8, 2, Nr=2, seed=1)
    d = g.d
    rt.init(d)
    rt.set_grid(g)
    z = np.zeros((d.PY, d.PX))
    z3 = np.zeros((d.Nr, d.PY, d.PX))
    base = [1, 1, 1, 8, 1, 8, 1, 1, 1, 2, z, z, z, z, z, z, z, z, z, z, 0.0, 0.0, z, np.zeros(2), z3, z3, np.ones(2), 1]
    with pytest.raises(rt.B200Error):      # GM/Redi
        rt.gad_calc_rhs(*base, 2, 2, 1, 0, 0, 0, 1, 0, 0, z.copy(), z.copy(), np.zeros((2, d.PY, d.PX)), z3.copy())
    with pytest.raises(rt.B200Error):      # not an advection scheme of GAD.h
        rt.gad_calc_rhs(*base, 99, 99, 1, 0, 0, 0, 0, 0, 0, z.copy(), z.copy(), np.zeros((2, d.PY, d.PX)), z3.copy())


MOM_CASES = [
    dict(),                                                                      # config-1 like: no-slip, harmonic
    dict(viscA4D=1e11, viscA4Z=1e11, useBiharmonicVisc=1),
    dict(no_slip_sides=0, no_slip_bottom=0, bottomDragLinear=1e-3),
    dict(selectBotDragQuadr=0, bottomDragQuadratic=2e-3, bottomVisc_pCell=1),
    dict(selectBotDragQuadr=1, bottomDragQuadratic=2e-3),
    dict(selectBotDragQuadr=2, bottomDragQuadratic=2e-3, selectCoriScheme=1),
    dict(selectCoriScheme=2, usingSphericalPolarGrid=1, selectMetricTerms=1),
    dict(selectCoriScheme=3, implicitViscosity=1),
    dict(momAdvection=0),
    dict(momViscosity=0, useCDscheme=1),
    dict(rigidLid=1),
]


@pytest.mark.parametrize("case", MOM_CASES, ids=lambda c: "-".join(f"{k}{v}" for k, v in c.items()) or "default")
@pytest.mark.parametrize("shape", [dict(sNx=31, sNy=17, OL=3, nSx=2, nSy=2, Nr=5), dict(sNx=66, sNy=34, OL=3, Nr=4)],
                         ids=["tiles2x2", "1tile"])
def test_mom_fluxform_matches_oracle(rt, case, shape):
    g = make_grid(**shape, seed=31)
    d = g.d
    rng = np.random.default_rng(2)
    g.a["tanPhiAtU"] = 0.5 * rng.random(d.shape2)
    g.a["tanPhiAtV"] = 0.5 * rng.random(d.shape2)
    g.a["cosFacU"] = 0.5 + 0.5 * rng.random((d.nSy, d.nSx, d.PY))
    g.a["cosFacV"] = 0.5 + 0.5 * rng.random((d.nSy, d.nSx, d.PY))
    params = dict(viscAhD=400.0, viscAhZ=300.0, no_slip_sides=1, no_slip_bottom=1, sideDragFactor=2.0)
    params.update(case)
    o = Oracle(g, params)
    u, v, w, _ = rand_state(g, 8)
    rt.init(d)
    rt.set_grid(g)
    rt.set_params(**params)
    ns = (d.PY, d.PX)
    iMin, iMax, jMin, jMax = 0, d.sNx + 1, 0, d.sNy + 1
    gU_o, gV_o, gU_g, gV_g = (np.zeros(d.shape3) for _ in range(4))
    for bj in range(1, d.nSy + 1):
        for bi in range(1, d.nSx + 1):
            kap = 1e-3 * (1 + rng.random((d.Nr + 1,) + ns))
            kav = 1e-3 * (1 + rng.random((d.Nr + 1,) + ns))
            fU_o, fV_o, fU_g, fV_g = (np.zeros((2,) + ns) for _ in range(4))
            for k in range(1, d.Nr + 1):
                kUp, kDown = 1 + (k + 1) % 2, 1 + k % 2
                gd_o, hd_o, gd_g, hd_g = (np.zeros(ns) for _ in range(4))
                o.mom_fluxform(bi, bj, k, iMin, iMax, jMin, jMax, kap, kav, fU_o[kUp - 1], fV_o[kUp - 1],
                               fU_o[kDown - 1], fV_o[kDown - 1], gd_o, hd_o, u, v, w, gU_o, gV_o)
                rt.mom_fluxform(bi, bj, k, iMin, iMax, jMin, jMax, kap, kav, fU_g[kUp - 1], fV_g[kUp - 1],
                                fU_g[kDown - 1], fV_g[kDown - 1], gd_g, hd_g, u, v, w, gU_g, gV_g)
                t = (bj - 1, bi - 1, k - 1)
                assert close(fU_g[kDown - 1], fU_o[kDown - 1]) and close(fV_g[kDown - 1], fV_o[kDown - 1]), (k, "fVer kp")
                assert close(fU_g[kUp - 1], fU_o[kUp - 1]) and close(fV_g[kUp - 1], fV_o[kUp - 1]), (k, "fVer km")
                assert close(gU_g[t], gU_o[t]) and close(gV_g[t], gV_o[t]), (k, "gU/gV")
                assert close(gd_g, gd_o) and close(hd_g, hd_o), (k, "dissipation")
    if case.get("momAdvection", 1) or not case.get("useCDscheme", 0):
        assert np.abs(gU_o).max() > 0


def test_mom_rejects_bad_range(rt):
    g = make_grid(8, 8, 2, Nr=2, seed=1)
    d = g.d
    rt.init(d)
    rt.set_grid(g)
    z = np.zeros((d.PY, d.PX))
    z3 = np.zeros(d.shape3)
    k3 = np.zeros((3, d.PY, d.PX))
    with pytest.raises(rt.B200Error):
        rt.mom_fluxform(1, 1, 1, 1 - d.OLx, d.sNx + d.OLx, 0, 9, k3, k3, z.copy(), z.copy(), z.copy(), z.copy(),
                        z.copy(), z.copy(), z3, z3, z3, z3.copy(), z3.copy())


def test_advect_xy_ab3_c4_golden_with_the_cuda_kernel(rt):
    """verification/advect_xy input.ab3_c4 (scheme 4 + AB3, 100 steps) with gad_calc_rhs_b200_ in the loop: every
    printed digit of %MON dynstat_theta_* / dynstat_salt_* at steps 10, ..., 100."""
    import json
    import os
    from oracle import advect_xy as ax
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "advect_xy.ab3_c4.json")))
    d, g, _, _ = ax.setup_ab3()
    rt.init(d)
    rt.set_grid(g)
    calls = [0]

    def cuda_rhs(bi, bj, iMin, iMax, jMin, jMax, k, kM1, kUp, kDown, xA, yA, maskUp, uFld, vFld, wFld, uTrans, vTrans,
                 rTrans, rTransKp1, diffKh, diffK4, KappaR, diffKr4, TracerN, TracAB, deltaTLev, advScheme, vertAdvScheme,
                 calcAdvection, implicitAdvection, applyAB_onTracer, trUseDiffKr4, fZon, fMer, fVerT, gTracer):
        calls[0] += 1
        rt.gad_calc_rhs(bi, bj, iMin, iMax, jMin, jMax, k, kM1, kUp, kDown, xA, yA, maskUp, uFld, vFld, wFld, uTrans, vTrans,
                        rTrans, rTransKp1, diffKh, diffK4, KappaR, diffKr4, TracerN, TracAB, deltaTLev, 1, advScheme,
                        vertAdvScheme, calcAdvection, implicitAdvection, applyAB_onTracer, trUseDiffKr4, 0, 0, 0, fZon, fMer,
                        fVerT, gTracer)
    out = ax.run_ab3(100, calc_rhs=cuda_rhs)
    assert calls[0] == 100 * 2 * 2
    for i, (t, s) in enumerate(out):
        if i == 0:
            continue
        for r, fld in ((t, "theta"), (s, "salt")):
            for st in ("max", "min", "mean", "sd"):
                assert f"{r[st]:.13E}" == gold[f"dynstat_{fld}_{st}"][i], (i, fld, st)
