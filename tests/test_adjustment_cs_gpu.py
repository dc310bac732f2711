"""verification/adjustment.cs-32x32x1 (cs32 cubed sphere, 48 tiles, pkg/exch2) with the CUDA kernels in
the loop through the C ABI: MOM_FLUXFORM per tile and CG2D on the exch2 tile graph (push table compiled
from the topology tables), against the experiment's golden output.

(a) CUDA MOM_FLUXFORM, CPU solver: tendencies are bit-identical, so the run must agree with the golden
    exactly as the all-CPU oracle run does (>= 13 digits, identical iteration counts), 24 steps.
(b) CUDA CG2D as well (tolerance 1e-13, dot products in a different order): the reference's pass rule --
    cg2d_init_res to >= 10 digits, iteration counts +-1, monitor statistics 1e-9."""
import json
import os

import numpy as np
import pytest

from helpers import CudaEngine
from oracle import adjustment_cs as ac

pytestmark = pytest.mark.gpu
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "adjustment.cs-32x32x1.json")))


@pytest.fixture()
def rt():
    from mitgcm_b200 import runtime
    yield runtime
    runtime.finalize()


def test_cuda_mom_fluxform_on_the_cube_reproduces_the_golden(rt):
    eng = CudaEngine(rt, use_gad=False, use_cg2d=False)      # CG2D stays on the run's own oracle (exch2 hook)
    norm, out = ac.run(24, engine=eng)
    assert f"{norm:.16E}" == GOLD["cg2dNorm"]
    assert [r["numIters"] for r in out] == GOLD["cg2d_iters"]
    for i, r in enumerate(out):
        assert r["firstResidual"] == pytest.approx(float(GOLD["cg2d_init_res"][i]), rel=1e-13)
        for f in ("eta", "uvel", "vvel", "wvel"):
            for st in ("max", "min", "sd"):
                assert r[f][st] == pytest.approx(float(GOLD[f"dynstat_{f}_{st}"][i + 1]), rel=2e-13, abs=1e-30), (i, f, st)


def test_all_cuda_kernels_on_the_cube_meet_the_reference_pass_rule(rt):
    eng = CudaEngine(rt, use_gad=False)
    _, out = ac.run(24, engine=eng)
    for i, r in enumerate(out):
        assert abs(r["numIters"] - GOLD["cg2d_iters"][i]) <= 1, i
        assert r["firstResidual"] == pytest.approx(float(GOLD["cg2d_init_res"][i]), rel=1e-10), i
        assert r["rhsMax"] == pytest.approx(float(GOLD["sumRHS_rhsMax"][i][1]), rel=1e-10), i
        for f in ("eta", "uvel", "vvel", "wvel"):
            for st in ("max", "min", "sd"):
                assert r[f][st] == pytest.approx(float(GOLD[f"dynstat_{f}_{st}"][i + 1]), rel=1e-9, abs=1e-30), (i, f, st)


def test_resident_cubed_sphere_step_matches_the_golden():
    """The same experiment stepped ENTIRELY on the device (mitgcm_b200_forward_step_ on the exch2 tile graph:
    generic dynamics kernel with the explicit half of the surface pressure gradient, CG2D on the cube,
    correction, exactConserv eta / etaH update, EXCH_UV_XYZ_RL with signs as one gather)."""
    from mitgcm_b200.model import Model, ini_cg2d_tilegraph
    from oracle.baroclinic_gyre import mon_stats
    T, d, g, P, ssh = ac.setup()
    P = dict(P)
    P.update(abEps=0.1, deltaTtracer=900.0, viscAr=0.0, tempStepping=0, cg2dMaxIters=600, momForcing=1,
             momDissip_In_AB=1, exactConserv=1, diffKhT=0.0, diffK4T=0.0, diffKrT=0.0)
    op = ini_cg2d_tilegraph(g, P, T)
    etaN = ac.tile_from_xstack(T, d, ssh)
    ac.eo.exch2_3d(T, etaN[0][:, None], d.OLx)
    z3 = np.zeros(d.shape3)
    state = dict(uVel=z3, vVel=z3, wVel=z3, theta=z3, etaN=etaN, etaH=etaN.copy(), surfForcU=np.zeros(d.shape2),
                 surfForcV=np.zeros(d.shape2))
    m = Model(g, P, state, op, device=0, topo=T)
    maskInC, maskInW, maskInS = g.maskC[:, :, 0], g.maskW[:, :, 0], g.maskS[:, :, 0]
    try:
        for it in range(24):
            r = m.step()
            assert abs(r["numIters"] - GOLD["cg2d_iters"][it]) <= 1, it
            assert r["firstResidual"] == pytest.approx(float(GOLD["cg2d_init_res"][it]), rel=1e-10), it
            st = dict(eta=mon_stats(d, m.get("etaN")[:, :, None], maskInC[:, :, None], maskInC, g.rA, [g.drF[0]]),
                      uvel=mon_stats(d, m.get("uVel"), g.hFacW, maskInW, g.rAw, g.drF),
                      vvel=mon_stats(d, m.get("vVel"), g.hFacS, maskInS, g.rAs, g.drF),
                      wvel=mon_stats(d, m.get("wVel"), g.maskC, maskInC, g.rA, g.drC[:1]))
            for f in ("eta", "uvel", "vvel", "wvel"):
                for s in ("max", "min", "sd"):
                    assert st[f][s] == pytest.approx(float(GOLD[f"dynstat_{f}_{s}"][it + 1]), rel=1e-9, abs=1e-30), (it, f, s)
    finally:
        m.close()
