"""verification/tutorial_advection_in_gyre (the barotropic gyre restarted from a 10-year spin-up, oracle/advection_in_gyre.py)
with the CUDA MOM_FLUXFORM in the loop through the C ABI (reference argument list, host buffers, 2 x 2 tiles of 30 x 30,
OL = 4): advective terms on a developed flow, harmonic viscosity, no-slip sides and bottom.  The tendencies are
bit-identical to the oracle's, so with the CPU solver the run reproduces every printed digit of the golden output --
including cg2d_init_res = 6.7e-10 and the 1e-14 wvel statistics, which are differences of nearly equal numbers.
(The CUDA solver is not put in this loop: the start residual of an almost steady state is what the previous solve left
behind, i.e. it depends on the summation order at the 1e-10 tolerance.)"""
import json
import os

import pytest

from helpers import CudaEngine
from oracle import advection_in_gyre as ag

pytestmark = pytest.mark.gpu
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "tutorial_advection_in_gyre.json")))


@pytest.fixture()
def rt():
    from mitgcm_b200 import runtime
    yield runtime
    runtime.finalize()


def test_cuda_mom_fluxform_on_the_spun_up_gyre_reproduces_the_golden(rt):
    eng = CudaEngine(rt, use_gad=False, use_cg2d=False)
    norm, first, out = ag.run(4, engine=eng)
    assert f"{norm:.16E}" == GOLD["cg2dNorm"]
    assert [r["numIters"] for r in out] == GOLD["cg2d_iters"]
    for i, r in enumerate(out):
        assert f"{r['firstResidual']:.14E}" == GOLD["cg2d_init_res"][i], i
        assert f"{r['rhsMax']:.14E}" == GOLD["sumRHS_rhsMax"][i][1], i
        for f in ("eta", "uvel", "vvel", "wvel"):
            for st in ("max", "min", "mean", "sd"):
                assert f"{r[f][st]:.13E}" == GOLD[f"dynstat_{f}_{st}"][i + 1], (i, f, st)
