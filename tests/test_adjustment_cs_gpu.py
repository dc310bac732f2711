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
