"""Size-independent properties of the oracle restatements that have no golden of their own (CPU only):
the implicit vertical viscosity solve inverts its tridiagonal operator, CG3D's solution satisfies the 7-point
system it was given, and the vector-invariant tendencies of a resting fluid vanish."""
import numpy as np

from helpers import make_grid
from mitgcm_b200.grid import global_area
from oracle.pyoracle import Oracle


def test_mom_implicit_r_inverts_the_vertical_viscosity_operator():
    g = make_grid(12, 9, 2, nSx=2, nSy=1, Nr=8, seed=3)
    d = g.d
    dt = 900.0
    o = Oracle(g, dict(deltaTMom=dt, implicitViscosity=1))
    rng = np.random.default_rng(1)
    kap = 5e-2 * (1 + rng.random((d.Nr + 1, d.PY, d.PX)))
    rhs = rng.standard_normal(d.shape3)
    y = rhs.copy()
    assert o.mom_implicit_r(1, 1, 0, kap, y) == 0
    # multiply back: (1 - b - d) y_k + b y_{k-1} + d y_{k+1} on i = 1..sNx+1, j = 1..sNy of tile (1,1)
    rdrF, rdrC = g.a["recip_drF"], g.a["recip_drC"]
    rh, mk = g.a["recip_hFacW"][0, 0], g.a["maskW"][0, 0]
    jj, ii = slice(d.OLy, d.OLy + d.sNy), slice(d.OLx, d.OLx + d.sNx + 1)
    back = np.zeros((d.Nr, d.sNy, d.sNx + 1))
    for k in range(d.Nr):
        b = np.zeros((d.sNy, d.sNx + 1)); dd = np.zeros_like(b)
        if k >= 1:
            b = np.where(mk[k - 1][jj, ii] == 1.0, -dt * rh[k][jj, ii] * rdrF[k] * kap[k][jj, ii] * rdrC[k], 0.0)
        if k <= d.Nr - 2:
            dd = np.where(mk[k + 1][jj, ii] == 1.0, -dt * rh[k][jj, ii] * rdrF[k] * kap[k + 1][jj, ii] * rdrC[k + 1], 0.0)
        back[k] = (1.0 - (b + dd)) * y[0, 0, k][jj, ii]
        if k >= 1:
            back[k] += b * y[0, 0, k - 1][jj, ii]
        if k <= d.Nr - 2:
            back[k] += dd * y[0, 0, k + 1][jj, ii]
    assert np.abs(back - rhs[0, 0][:, jj, ii]).max() < 1e-12 * np.abs(rhs).max()
    # untouched outside the routine's range and on the other tile
    assert np.array_equal(y[0, 1], rhs[0, 1])


def test_cg3d_solution_satisfies_the_seven_point_system():
    g = make_grid(20, 14, 2, nSx=1, nSy=2, Nr=6, dx=200.0, dz=[20.0] * 6, seed=7)
    d = g.d
    o = Oracle(g, dict(deltaTMom=20.0, deltaTFreeSurf=20.0, globalArea=global_area(g)))
    op = o.ini_cg3d(1.0, 1e-11, -1.0)
    rng = np.random.default_rng(2)
    jj, ii = d.interior()
    b = np.zeros(d.shape3)
    b[..., jj, ii] = rng.standard_normal(b[..., jj, ii].shape)
    b *= g.maskC
    b0 = b.copy()
    x = np.zeros(d.shape3)
    r = o.cg3d(op, b, x, 3000)
    assert r["numIters"] < 3000 and r["lastResidual"] < 1e-11
    o.exch_xyz(x, d.Nr)
    aW, aS, aV, aC = op["aW3d"], op["aS3d"], op["aV3d"], op["aC3d"]
    Ax = aC * x
    Ax[..., :, 1:] += aW[..., :, 1:] * x[..., :, :-1]
    Ax[..., :, :-1] += aW[..., :, 1:] * x[..., :, 1:]
    Ax[..., 1:, :] += aS[..., 1:, :] * x[..., :-1, :]
    Ax[..., :-1, :] += aS[..., 1:, :] * x[..., 1:, :]
    Ax[:, :, 1:] += aV[:, :, 1:] * x[:, :, :-1]
    Ax[:, :, :-1] += aV[:, :, 1:] * x[:, :, 1:]
    rhs = b0 * op["cg3dNorm"]
    res = (Ax - rhs)[..., jj, ii] * g.maskC[..., jj, ii]
    assert np.abs(res).max() < 1e-9 * np.abs(rhs).max()


def test_mom_vecinv_of_a_resting_fluid_is_zero():
    g = make_grid(16, 12, 3, Nr=4, seed=5)
    d = g.d
    o = Oracle(g, dict(viscAhD=300.0, viscAhZ=300.0, no_slip_sides=1, no_slip_bottom=1, useCoriolis=1))
    z3 = np.zeros(d.shape3)
    ns = (d.PY, d.PX)
    kap = np.full((d.Nr + 1,) + ns, 1e-3)
    gU, gV = np.full(d.shape3, 3.0), np.full(d.shape3, 3.0)
    for k in range(1, d.Nr + 1):
        f = [np.zeros(ns) for _ in range(6)]
        o.mom_vecinv(1, 1, k, 0, d.sNx + 1, 0, d.sNy + 1, kap, kap, f[0], f[1], f[2], f[3], f[4], f[5], z3, z3, z3, gU, gV)
        assert not f[4].any() and not f[5].any()
    sl = (0, 0, slice(None), slice(d.OLy - 1, d.OLy + d.sNy + 1), slice(d.OLx - 1, d.OLx + d.sNx + 1))
    assert not gU[sl].any() and not gV[sl].any()
