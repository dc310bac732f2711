"""pkg/exch2 on the GPU: the tile-graph exchange as one gather launch, and
CG2D / CG2D_SR on the cubed sphere (config 4 geometry: 12 tiles of 32x16, OL = 4) through the C ABI."""
import numpy as np
import pytest

from helpers import load_cs32
from mitgcm_b200.exch2 import cubed_sphere_topology, set_topology
from mitgcm_b200.grid import Dims
from oracle import exch2_oracle as eo
from oracle.pyoracle import Oracle

pytestmark = pytest.mark.gpu


@pytest.fixture()
def rt():
    from mitgcm_b200 import runtime
    yield runtime
    runtime.finalize()


@pytest.mark.parametrize("nf,sx,sy,OL,Nr", [(32, 32, 16, 4, 3), (8, 4, 4, 2, 2), (6, 6, 6, 3, 1), (16, 8, 4, 2, 5)])
def test_device_exchange_is_bit_identical_to_exch2_3d(rt, nf, sx, sy, OL, Nr):
    T = cubed_sphere_topology(nf, sx, sy)
    d = Dims(sNx=sx, sNy=sy, OLx=OL, OLy=OL, nSx=T.nTiles, nSy=1, Nr=Nr)
    rt.init(d)
    set_topology(T)
    rng = np.random.default_rng(7)
    a3 = rng.standard_normal(d.shape3)
    a2 = rng.standard_normal(d.shape2)
    rt.set_field("theta", a3)
    rt.set_field("etaN", a2)
    rt.exch("theta")
    rt.exch("etaN")
    g3, g2 = rt.get_field("theta", np.zeros(d.shape3)), rt.get_field("etaN", np.zeros(d.shape2))
    eo.exch2_3d(T, a3[0], OL)
    r2 = a2[0][:, None].copy()
    eo.exch2_3d(T, r2, OL)
    assert np.array_equal(g3, a3) and np.array_equal(g2[0], r2[:, 0])


def test_bad_topology_is_rejected(rt):
    T = cubed_sphere_topology(8, 4, 4)
    rt.init(Dims(sNx=4, sNy=4, OLx=2, OLy=2, nSx=12, nSy=2, Nr=1))
    T.oi[0, 0] += 100                       # index map that leaves the tile array
    with pytest.raises(RuntimeError):
        set_topology(T)


@pytest.mark.parametrize("sr", [False, True], ids=["cg2d", "cg2d_sr"])
def test_cg2d_on_the_cubed_sphere_matches_the_oracle(rt, sr):
    """Config-4 operator (real bathymetry, cs32 grid files).  Fixed iteration counts: normalised RHS
    bit-exact, x to 1e-11*nit; converged solve: same count +-1, residuals 1e-9."""
    from mitgcm_b200.model import ini_cg2d_tilegraph
    T, g, P = load_cs32()
    d = g.d
    op = ini_cg2d_tilegraph(g, P, T)
    o = Oracle(g, P)
    hook = eo.Exch2Hook(o, T, d.OLx)
    try:
        rng = np.random.default_rng(5)
        jj, ii = d.interior()
        wet = g.maskC[:, :, 0]
        b = np.zeros(d.shape2)
        b[:, :, jj, ii] = rng.standard_normal((1, 12, 16, 32)) * wet[:, :, jj, ii] * 1e-3
        b[:, :, jj, ii] -= b[:, :, jj, ii].sum() / wet[:, :, jj, ii].sum() * wet[:, :, jj, ii]
        b *= g.rA
        x = 0.01 * rng.standard_normal(d.shape2) * wet
        rt.init(d)
        rt.set_grid(g)
        set_topology(T)
        rt.set_cg2d_operator(op)
        for nit in (1, 2, 7, 25):
            bo, xo, bg, xg = b.copy(), x.copy(), b.copy(), x.copy()
            ro = o.cg2d(op, bo, xo, nit, -1, sr=sr, history=True)
            rg = rt.cg2d(bg, xg, nit, -1, sr=sr, residuals=True)
            assert rg["numIters"] == ro["numIters"] == nit
            assert np.array_equal(bg[:, :, jj, ii], bo[:, :, jj, ii])
            sc = np.abs(xo[:, :, jj, ii]).max()
            assert np.abs(xg[:, :, jj, ii] - xo[:, :, jj, ii]).max() <= 1e-11 * nit * sc
            assert np.allclose(rg["hist"], ro["hist"], rtol=1e-9, atol=0)
        bo, xo, bg, xg = b.copy(), x.copy(), b.copy(), x.copy()
        ro = o.cg2d(op, bo, xo, 200, -1, sr=sr)
        rg = rt.cg2d(bg, xg, 200, -1, sr=sr)
        assert ro["numIters"] < 200 and abs(rg["numIters"] - ro["numIters"]) <= 1
        assert rg["firstResidual"] == pytest.approx(ro["firstResidual"], rel=1e-12)
        assert rg["lastResidual"] < np.sqrt(op["cg2dTolerance_sq"])
    finally:
        hook.close()


def test_cg2d_config3_real_bathymetry_operator(rt):
    """Config 3 (global_ocean.90x40x15): 9x4 tiles of 10x10 with OL = 3 (odd overlap: scalar path of
    the kernel), tolerance 1e-13, operator from the experiment's bathymetry."""
    from test_oracle_golden import config3_grid
    from mitgcm_b200.grid import global_area
    g = config3_grid()
    d = g.d
    P = dict(deltaTMom=1800.0, deltaTFreeSurf=86400.0, cg2dTargetResidual=1e-13, globalArea=global_area(g))
    o = Oracle(g, P)
    op = o.ini_cg2d()
    rng = np.random.default_rng(3)
    jj, ii = d.interior()
    b = np.zeros(d.shape2)
    b[:, :, jj, ii] = rng.standard_normal((d.nSy, d.nSx, d.sNy, d.sNx))
    b *= g.maskC[:, :, 0] * g.rA / 1800.0
    x = np.zeros(d.shape2)
    rt.init(d)
    rt.set_grid(g)
    rt.set_cg2d_operator(op)
    bo, xo, bg, xg = b.copy(), x.copy(), b.copy(), x.copy()
    ro = o.cg2d(op, bo, xo, 500, -1)
    rg = rt.cg2d(bg, xg, 500, -1)
    assert abs(rg["numIters"] - ro["numIters"]) <= 1 and rg["lastResidual"] < 1e-13
    assert rg["firstResidual"] == pytest.approx(ro["firstResidual"], rel=1e-12)
    sc = np.abs(xo[:, :, jj, ii]).max()
    assert np.abs(xg[:, :, jj, ii] - xo[:, :, jj, ii]).max() <= 1e-9 * sc


@pytest.mark.parametrize("withSigns", [True, False])
@pytest.mark.parametrize("nf,sx,sy,OL,Nr", [(32, 32, 16, 4, 3), (8, 4, 4, 2, 2), (6, 6, 6, 3, 1)])
def test_device_vector_exchange_is_bit_identical_to_exch2_uv_3d(rt, nf, sx, sy, OL, Nr, withSigns):
    """EXCH_UV_XYZ_RL on device mirrors (one gather launch for both components) against the literal
    restatement of EXCH2_UV_3D_RX (two EXCH2_RX2_CUBE passes + cube-corner fix-ups)."""
    T = cubed_sphere_topology(nf, sx, sy)
    d = Dims(sNx=sx, sNy=sy, OLx=OL, OLy=OL, nSx=T.nTiles, nSy=1, Nr=Nr)
    rt.init(d)
    set_topology(T)
    rng = np.random.default_rng(11)
    u, v = rng.standard_normal(d.shape3), rng.standard_normal(d.shape3)
    rt.set_field("uVel", u)
    rt.set_field("vVel", v)
    rt.exch_uv("uVel", "vVel", withSigns)
    gu, gv = rt.get_field("uVel", np.zeros(d.shape3)), rt.get_field("vVel", np.zeros(d.shape3))
    eo.exch2_uv_3d(T, u[0], v[0], OL, withSigns)
    assert np.array_equal(gu, u) and np.array_equal(gv, v)


def test_vector_exchange_on_the_periodic_tiling_is_two_scalar_exchanges(rt):
    from mitgcm_b200.grid import exch_xyz
    d = Dims(sNx=12, sNy=8, OLx=3, OLy=3, nSx=2, nSy=2, Nr=2)
    rt.init(d)
    rng = np.random.default_rng(2)
    u, v = rng.standard_normal(d.shape3), rng.standard_normal(d.shape3)
    rt.set_field("uVel", u)
    rt.set_field("vVel", v)
    rt.exch_uv("uVel", "vVel", True)
    assert np.array_equal(rt.get_field("uVel", np.zeros(d.shape3)), exch_xyz(d, u))
    assert np.array_equal(rt.get_field("vVel", np.zeros(d.shape3)), exch_xyz(d, v))
