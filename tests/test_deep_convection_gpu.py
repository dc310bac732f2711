"""verification/tutorial_deep_convection (the non-hydrostatic step, oracle/deep_convection.py, pinned to the golden
output on the CPU) with the CUDA kernels in the loop, through the C ABI with host buffers:

  * cg3d_b200_ alone in an otherwise-oracle step: its inputs on step 1 are bit-identical to the oracle's, so the
    printed solver lines must come out to the digits the dot-product summation order allows -- rhsMax every digit,
    cg3d_init_res >= 12 digits, the residual after 100 iterations >= 6 digits;
  * cg3d_b200_ + cg2d_b200_ (min-residual solution, cg2dUseMinResSol = 1) + gad_calc_rhs_b200_ + mom_fluxform_b200_:
    iteration counts +-1, statistics to the solver tolerance.
"""
import json
import os

import numpy as np
import pytest

from helpers import CudaEngine

pytestmark = pytest.mark.gpu
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "tutorial_deep_convection.json")))


@pytest.fixture()
def rt():
    from mitgcm_b200 import runtime
    yield runtime
    runtime.finalize()


def _oracle_fallback():
    from oracle import deep_convection as dc
    from oracle.pyoracle import Oracle
    d, g, P = dc.setup()
    return Oracle(g, P)


def test_cuda_cg3d_in_the_oracle_step_reproduces_the_golden_solver_lines(rt):
    from oracle import deep_convection as dc
    fb = _oracle_fallback()
    eng = CudaEngine(rt, use_gad=False, use_mom=False, use_cg2d=False, fallback=fb, use_cg3d=True)
    op, op3, out, rec0 = dc.run(3, engine=eng)
    for n, r in enumerate(out):
        c = r["cg3d"]
        assert c["numIters"] == 100
        tol = 1e-12 if n == 0 else 1e-7            # later steps start from the CUDA phi_nh of the step before
        assert c["rhsMax"] == pytest.approx(float(GOLD["cg3d_sumRHS_rhsMax"][n][1]), rel=tol)
        assert c["firstResidual"] == pytest.approx(float(GOLD["cg3d_init_res"][n]), rel=tol)
        assert c["lastResidual"] == pytest.approx(float(GOLD["cg3d_last_res"][n]), rel=1e-6 if n == 0 else 1e-4)
        if n == 0:
            assert f"{c['rhsMax']:.14E}" == GOLD["cg3d_sumRHS_rhsMax"][0][1]
        # the CPU CG2D of this run sees the CUDA phi_nh of the previous step in its right-hand side
        assert abs(r["numIters"] - GOLD["cg2d_iters"][n]) <= 1
        for fld in ("uvel", "vvel", "wvel", "theta", "eta"):
            for st in ("max", "sd"):
                assert r[fld][st] == pytest.approx(float(GOLD[f"dynstat_{fld}_{st}"][n + 1]), rel=1e-7), (n, fld, st)


def test_all_cuda_kernels_in_the_non_hydrostatic_step(rt):
    from oracle import deep_convection as dc
    fb = _oracle_fallback()
    eng = CudaEngine(rt, fallback=fb)
    op, op3, out, rec0 = dc.run(3, engine=eng)
    for n, r in enumerate(out):
        assert abs(r["numIters"] - GOLD["cg2d_iters"][n]) <= 1
        assert abs(r["nIterMin"] - GOLD["cg2d_iters_min"][n]) <= 1
        assert r["firstResidual"] == pytest.approx(float(GOLD["cg2d_init_res"][n]), rel=1e-6)
        assert r["rhsMax"] == pytest.approx(float(GOLD["sumRHS_rhsMax"][n][1]), rel=1e-7)
        c = r["cg3d"]
        assert c["numIters"] == 100
        assert c["firstResidual"] == pytest.approx(float(GOLD["cg3d_init_res"][n]), rel=1e-5)
        assert c["rhsMax"] == pytest.approx(float(GOLD["cg3d_sumRHS_rhsMax"][n][1]), rel=1e-6)
        assert c["lastResidual"] == pytest.approx(float(GOLD["cg3d_last_res"][n]), rel=1e-3)
        for fld in ("uvel", "vvel", "wvel", "theta", "eta"):
            for st in ("max", "sd"):
                assert r[fld][st] == pytest.approx(float(GOLD[f"dynstat_{fld}_{st}"][n + 1]), rel=1e-6), (n, fld, st)
