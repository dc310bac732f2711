"""verification/adjustment.128x64x1 (global lat-lon grid from pole to pole, one-layer atmosphere in p coordinates)
with CUDA MOM_FLUXFORM and CUDA CG2D in the loop through the C ABI, against the experiment's golden output:
identical iteration counts for 24 steps, cg2d_init_res and the monitor statistics to 1e-10 (the golden itself is
from an older model version: the all-CPU oracle agrees with it to 4e-12)."""
import json
import os

import pytest

from helpers import CudaEngine
from oracle import adjustment_latlon as al

pytestmark = pytest.mark.gpu
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "adjustment.128x64x1.json")))


@pytest.fixture()
def rt():
    from mitgcm_b200 import runtime
    yield runtime
    runtime.finalize()


def test_cuda_kernels_on_the_lat_lon_sphere_reproduce_the_golden(rt):
    _, out = al.run(24, engine=CudaEngine(rt, use_gad=False))
    assert [r["numIters"] for r in out] == GOLD["cg2d_iters"]
    for n, r in enumerate(out):
        assert r["firstResidual"] == pytest.approx(float(GOLD["cg2d_init_res"][n]), rel=1e-10), n
        for f in ("eta", "uvel", "vvel"):
            for st in ("max", "min", "sd"):
                assert r[f][st] == pytest.approx(float(GOLD[f"dynstat_{f}_{st}"][n + 1]), rel=1e-10, abs=1e-12), (n, f, st)
