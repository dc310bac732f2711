"""CG3D (model/src/cg3d.F) on the GPU through the C ABI (reference argument list, host buffers) against the CPU
oracle restatement (oracle/cg3d_oracle.c; parity UNPINNED beyond the operator normalisation known answer).

Tolerances as for CG2D: point-wise arithmetic is identical, dot products are summed in a different order.  With a
fixed iteration count the normalised RHS is bit-exact and x agrees to 1e-11 * iterations; converged solves agree
in the iteration count +-1 and in x to the solver tolerance."""
import numpy as np
import pytest

from helpers import make_grid
from mitgcm_b200.grid import global_area
from oracle.pyoracle import Oracle

pytestmark = pytest.mark.gpu


@pytest.fixture()
def rt():
    from mitgcm_b200 import runtime
    yield runtime
    runtime.finalize()


@pytest.fixture(params=["fused", "four-sweep"])
def kernel(request, monkeypatch):
    """Both forms of the solver kernel (csrc/cg3d.cu) must pass every test."""
    if request.param == "four-sweep":
        monkeypatch.setenv("MITGCM_B200_CG3D_UNFUSED", "1")
    else:
        monkeypatch.delenv("MITGCM_B200_CG3D_UNFUSED", raising=False)
    return request.param


def problem(shape, seed=3, tol=1e-9, wunit=-1.0):
    g = make_grid(**shape, seed=7)
    d = g.d
    o = Oracle(g, dict(deltaTMom=20.0, deltaTFreeSurf=20.0, globalArea=global_area(g)))
    op = o.ini_cg3d(1.0, tol, wunit)
    rng = np.random.default_rng(seed)
    jj, ii = d.interior()
    b = np.zeros(d.shape3)
    b[..., jj, ii] = rng.standard_normal(b[..., jj, ii].shape)
    b *= g.maskC
    x = 0.1 * rng.standard_normal(d.shape3) * g.maskC
    return g, o, op, b, x


SHAPES = [dict(sNx=31, sNy=17, OL=2, nSx=2, nSy=2, Nr=6, dx=200.0, dz=[20.0] * 6),
          dict(sNx=40, sNy=24, OL=3, Nr=12, dx=100.0, dz=[10.0] * 12),
          dict(sNx=16, sNy=16, OL=2, nSx=3, nSy=1, Nr=1, dx=500.0, dz=[50.0]),
          dict(sNx=64, sNy=8, OL=2, nSx=1, nSy=4, Nr=5, dx=50.0, dz=[5.0, 10.0, 20.0, 30.0, 40.0], land_frac=0.3)]


@pytest.mark.parametrize("shape", SHAPES, ids=["tiles2x2", "1tile-OL3", "Nr1", "strips-land"])
def test_cg3d_fixed_iterations_match_oracle(rt, kernel, shape):
    g, o, op, b, x = problem(shape, tol=0.0)
    d = g.d
    nit = 25
    bo, xo = b.copy(), x.copy()
    ro = o.cg3d(op, bo, xo, nit)
    rt.init(d)
    rt.set_grid(g)
    rt.set_cg3d_operator(op)
    bg, xg = b.copy(), x.copy()
    rg = rt.cg3d(bg, xg, nit)
    assert rg["numIters"] == ro["numIters"] == nit
    assert np.array_equal(bg, bo)                                   # normalised RHS: no reduction involved beyond max
    assert rg["rhsMax"] == ro["rhsMax"]
    assert rg["firstResidual"] == pytest.approx(ro["firstResidual"], rel=1e-13)
    assert rg["sumRHS"] == pytest.approx(ro["sumRHS"], abs=1e-12 * np.abs(bo).sum())
    jj, ii = d.interior()
    scale = np.abs(xo[..., jj, ii]).max()
    assert np.abs(xg[..., jj, ii] - xo[..., jj, ii]).max() <= 1e-11 * nit * scale
    assert rg["lastResidual"] == pytest.approx(ro["lastResidual"], rel=1e-9)
    assert ro["lastResidual"] < ro["firstResidual"]


@pytest.mark.parametrize("shape", SHAPES[:2], ids=["tiles2x2", "1tile-OL3"])
def test_cg3d_converged_solve_matches_oracle(rt, kernel, shape):
    g, o, op, b, x = problem(shape, tol=1e-10)
    d = g.d
    bo, xo = b.copy(), x.copy()
    ro = o.cg3d(op, bo, xo, 2000)
    assert ro["numIters"] < 2000 and ro["lastResidual"] < 1e-10
    rt.init(d)
    rt.set_grid(g)
    rt.set_cg3d_operator(op)
    bg, xg = b.copy(), x.copy()
    rg = rt.cg3d(bg, xg, 2000)
    assert abs(rg["numIters"] - ro["numIters"]) <= 1
    assert rg["lastResidual"] < 1e-10
    jj, ii = d.interior()
    assert np.abs(xg[..., jj, ii] - xo[..., jj, ii]).max() <= 1e-7 * np.abs(xo[..., jj, ii]).max()


def test_cg3d_zero_rhs_and_unnormalised_tolerance(rt, kernel):
    g, o, op, b, x = problem(SHAPES[0], tol=1e-9, wunit=1e-12)      # cg3dTargetResWunit > 0: no RHS normalisation
    assert not op["cg3dNormaliseRHS"]
    d = g.d
    rt.init(d)
    rt.set_grid(g)
    rt.set_cg3d_operator(op)
    bo, xo, bg, xg = b.copy(), x.copy(), b.copy(), x.copy()
    ro, rg = o.cg3d(op, bo, xo, 30), rt.cg3d(bg, xg, 30)
    assert rg["numIters"] == ro["numIters"]
    jj, ii = d.interior()
    assert np.abs(xg[..., jj, ii] - xo[..., jj, ii]).max() <= 1e-9 * np.abs(xo[..., jj, ii]).max()
    z = np.zeros(d.shape3)
    r0 = rt.cg3d(z.copy(), z.copy(), 10)                            # zero RHS, zero first guess: immediate exit
    assert r0["numIters"] == 0 and r0["firstResidual"] == 0.0


@pytest.mark.parametrize("variant", [0, 1, 2, 3])
def test_cg3d_fused_kernel_variants_and_pending_x_update(rt, monkeypatch, variant):
    """The fused kernel applies x += alpha s two iterations at a time; a solve that ends on an odd iteration (1, 3,
    maxIters or convergence) owes x one update.  Every (CTAs per SM, levels in flight) variant, iteration counts
    0..5 and 8: x, r-derived residuals against the oracle; x of the fused and the four-sweep kernel agree to
    round-off of the dot products."""
    g, o, op, b, x = problem(SHAPES[0], tol=0.0)
    d = g.d
    rt.init(d)
    rt.set_grid(g)
    rt.set_cg3d_operator(op)
    jj, ii = d.interior()
    for nit in (0, 1, 2, 3, 4, 5, 8):
        bo, xo = b.copy(), x.copy()
        ro = o.cg3d(op, bo, xo, nit)
        monkeypatch.setenv("MITGCM_B200_CG3D_VARIANT", str(variant))
        monkeypatch.delenv("MITGCM_B200_CG3D_UNFUSED", raising=False)
        bg, xg = b.copy(), x.copy()
        rg = rt.cg3d(bg, xg, nit)
        monkeypatch.setenv("MITGCM_B200_CG3D_UNFUSED", "1")
        b4, x4 = b.copy(), x.copy()
        r4 = rt.cg3d(b4, x4, nit)
        assert rg["numIters"] == r4["numIters"] == ro["numIters"] == nit
        scale = np.abs(xo[..., jj, ii]).max()
        assert np.abs(xg[..., jj, ii] - xo[..., jj, ii]).max() <= 1e-12 * max(nit, 1) * scale, nit
        assert np.abs(xg[..., jj, ii] - x4[..., jj, ii]).max() <= 1e-12 * max(nit, 1) * scale, nit
        assert rg["lastResidual"] == pytest.approx(ro["lastResidual"], rel=1e-10), nit
        assert np.array_equal(bg, bo)


def test_cg3d_fused_converges_on_odd_and_even_iterations(rt, monkeypatch):
    """Convergence decided on the device with an update of x pending: tolerances chosen so that the oracle stops
    after an odd and after an even number of iterations."""
    seen = set()
    for tol in (3e-2, 1e-2, 3e-3, 1e-3, 3e-4, 1e-4, 3e-5, 1e-5):
        g, o, op, b, x = problem(SHAPES[1], tol=tol)
        d = g.d
        bo, xo = b.copy(), x.copy()
        ro = o.cg3d(op, bo, xo, 500)
        rt.init(d)
        rt.set_grid(g)
        rt.set_cg3d_operator(op)
        bg, xg = b.copy(), x.copy()
        rg = rt.cg3d(bg, xg, 500)
        rt.finalize()
        assert rg["numIters"] == ro["numIters"], tol       # far from round-off: the same iteration stops both
        seen.add(ro["numIters"] % 2)
        jj, ii = d.interior()
        assert np.abs(xg[..., jj, ii] - xo[..., jj, ii]).max() <= 1e-11 * ro["numIters"] * np.abs(xo[..., jj, ii]).max(), tol
    assert seen == {0, 1}


def test_cg3d_needs_its_operator(rt):
    g = make_grid(8, 8, 2, Nr=3, seed=1)
    rt.init(g.d)
    rt.set_grid(g)
    z = np.zeros(g.d.shape3)
    with pytest.raises(rt.B200Error):
        rt.cg3d(z.copy(), z.copy(), 5)
