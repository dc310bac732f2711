"""Parity at the scale bench.py runs at: the kernels the headline number comes from --
the double2 / resident-strip CG2D with a full cooperative grid (296 CTAs), the TMA-staged MOM_FLUXFORM kernel, the
pipelined GAD_CALC_RHS kernel -- against the CPU oracle on the bench workload itself, not on the small grids of the
other test files.

  * cg2d_b200_ on ONE 2048 x 2048 tile (random bathymetry, 15 % land, partial cells) for 1, 2 and 25 fixed
    iterations, host arrays through the C ABI, against the oracle CG2D: the normalised right-hand side and rhsMax
    bit for bit, x to 1e-12 after 1 and 2 iterations (north-star tolerance), 25 iterations to the conditioning of
    the recurrences;
  * bench parameters (bench.params: linear EOS, c2 advection, harmonic viscosity, AB2, f = f0 + df sin) on a
    512 x 512 x 50 block, 3 steps: the GPU steps one 512 x 512 tile (the bench's tiling), the oracle the same global
    field cut into 4 x 4 tiles on 16 threads (the state is generated per global index, so both see the same numbers);
    tendencies of step 0 bit for bit, fields <= 1e-12 after the first solve, <= 1e-9 after 3 steps, CG2D iteration
    counts +-1.
"""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from helpers import make_grid, cg2d_problem

pytestmark = pytest.mark.gpu


def relerr(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def to_global(d, a):
    """(nSy, nSx, ..., PY, PX) tile array -> (..., Ny, Nx) interior."""
    out = np.zeros(a.shape[2:-2] + (d.Ny, d.Nx))
    for bj in range(d.nSy):
        for bi in range(d.nSx):
            out[..., bj * d.sNy:(bj + 1) * d.sNy, bi * d.sNx:(bi + 1) * d.sNx] = \
                a[bj, bi, ..., d.OLy:d.OLy + d.sNy, d.OLx:d.OLx + d.sNx]
    return out


def test_cg2d_2048_fixed_iterations_match_oracle():
    from mitgcm_b200 import runtime as rt
    g = make_grid(2048, 2048, 2, seed=11)
    o, op, b, x = cg2d_problem(g, tol=1e-30)
    jj, ii = g.d.interior()
    try:
        rt.init(g.d)
        rt.set_cg2d_operator(op)
        for nit in (1, 2, 25):
            bo, xo, bg, xg = b.copy(), x.copy(), b.copy(), x.copy()
            ro = o.cg2d(op, bo, xo, nit, -1, history=True)
            rg = rt.cg2d(bg, xg, nit, -1, residuals=True)
            assert rg["numIters"] == ro["numIters"] == nit
            assert np.array_equal(bg[:, :, jj, ii], bo[:, :, jj, ii])
            assert rg["rhsMax"] == ro["rhsMax"]
            assert rg["firstResidual"] == pytest.approx(ro["firstResidual"], rel=1e-12)
            assert relerr(xg[:, :, jj, ii], xo[:, :, jj, ii]) < (1e-12 if nit <= 2 else 1e-10), nit
            np.testing.assert_allclose(rg["hist"][:nit], ro["hist"], rtol=1e-9)
    finally:
        rt.finalize()


def test_bench_workload_512x512x50_matches_oracle():
    import bench
    from mitgcm_b200.model import make_channel, Model
    from oracle.channel import ChannelOracle
    n, nr, nS = 512, 50, 4
    P = bench.params(nr)
    # the oracle's copy: 4 x 4 tiles on 16 threads; the GPU's copy: one tile, as in bench.py
    go, Po, so = make_channel(n // nS, n // nS, nr, nSx=nS, nSy=nS, block=(n, n), **P)
    co = ChannelOracle(go, Po, so, threads=16)
    gg, Pg, sg = make_channel(n, n, nr, block=(n, n), **P)
    for name in ("uVel", "vVel", "theta", "etaN"):
        assert np.array_equal(to_global(go.d, so[name]), to_global(gg.d, sg[name])), name
    from mitgcm_b200.model import ini_cg2d
    m = Model(gg, Pg, sg, ini_cg2d(gg, Pg))
    del sg
    try:
        for it in range(3):
            ro, rg = co.step(), m.step()
            assert abs(ro["numIters"] - rg["numIters"]) <= 1, (it, ro["numIters"], rg["numIters"])
            assert rg["firstResidual"] == pytest.approx(ro["firstResidual"], rel=1e-9), it
            if it == 0:
                # MOM_FLUXFORM + TIMESTEP and GAD_CALC_RHS + AB2 before the solver feeds back: point-wise code, bit for bit
                for name in ("gU", "gV", "gtNm1"):
                    a, b = to_global(gg.d, m.get(name)), to_global(go.d, co.s[name])
                    assert np.array_equal(a, b), (name, relerr(a, b))
            tol = 1e-12 if it == 0 else 1e-9
            for name in ("uVel", "vVel", "theta", "etaN"):
                a, b = to_global(gg.d, m.get(name)), to_global(go.d, co.s[name])
                assert relerr(a, b) < tol, (it, name, relerr(a, b))
    finally:
        m.close()
