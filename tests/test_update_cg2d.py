"""UPDATE_CG2D (model/src/update_cg2d.F:57-192): the per-step rebuild of the CG2D operator under the non-linear free
surface / r* (configs 3 and 4).  CPU: the oracle restatement against INI_CG2D on the config-3 bathymetry.  GPU: the
device kernels (update_cg2d_b200_) against the oracle, bit for bit, with a surface-following column thickness."""
import numpy as np
import pytest

from helpers import make_grid
from test_oracle_golden import config3_grid


def _perturb_hfac(g, seed=11, amp=0.03):
    """What r* does to the open-water fractions (calc_surf_dr.F / update_surf_dr.F): hFac of wet cells scaled by a
    smooth column factor 1 + eta/H; masks unchanged."""
    rng = np.random.default_rng(seed)
    d = g.d
    f = 1.0 + amp * np.sin(np.linspace(0, 9, d.PX))[None, None, None, None, :] * np.cos(np.linspace(0, 7, d.PY))[None, None, None, :, None]
    f = f + 0.0 * rng.random()
    for n in ("hFacW", "hFacS"):
        g.a[n] = np.ascontiguousarray(g.a[n] * f)


def _params(g, **kw):
    from mitgcm_b200.grid import global_area
    p = dict(deltaTMom=1800.0, deltaTFreeSurf=86400.0, cg2dTargetResidual=1e-13, globalArea=global_area(g))
    p.update(kw)
    return p


def test_oracle_update_cg2d_reproduces_ini_cg2d_when_hfac_is_unchanged():
    """With the hFac of INI_CG2D, UPDATE_CG2D must give the same operator on the ranges it fills (it multiplies by
    the normalisation in another order: a few ulps) and zero aW / aS outside 1..sN+1, which it does not exchange."""
    from oracle.pyoracle import Oracle
    g = config3_grid()
    d = g.d
    o = Oracle(g, _params(g))
    op0 = o.ini_cg2d()
    op = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in op0.items()}
    o.update_cg2d(op, True)
    jj, ii = d.interior()
    for n in ("aW2d", "aS2d", "aC2d", "pC", "pW", "pS"):
        a, b = op[n][:, :, jj, ii], op0[n][:, :, jj, ii]
        assert np.abs(a - b).max() <= 4e-16 * np.abs(b).max(), n
    # halo of the halo: zero after UPDATE_CG2D (update_cg2d.F:66-72), exchanged values after INI_CG2D
    assert np.all(op["aW2d"][:, :, :, :d.OLx] == 0.0) and np.all(op["aS2d"][:, :, :d.OLy, :] == 0.0)
    # ... and the column i = sNx+1 / row j = sNy+1 it does fill equals what the exchange gave INI_CG2D
    e = (slice(None), slice(None), jj, d.OLx + d.sNx)
    assert np.abs(op["aW2d"][e] - op0["aW2d"][e]).max() <= 4e-16 * np.abs(op0["aW2d"]).max()


def test_oracle_update_cg2d_follows_the_column_thickness():
    from oracle.pyoracle import Oracle
    g = config3_grid()
    o = Oracle(g, _params(g))
    op = o.ini_cg2d()
    a0 = op["aW2d"].copy()
    _perturb_hfac(g)
    o2 = Oracle(g, _params(g))
    o2.update_cg2d(op, True)
    jj, ii = g.d.interior()
    wet = a0[:, :, jj, ii] != 0
    rel = np.abs(op["aW2d"][:, :, jj, ii][wet] / a0[:, :, jj, ii][wet] - 1.0)
    assert 1e-4 < rel.max() < 0.04          # the 3 % change of the column went into the operator
    assert np.array_equal(op["aW2d"][:, :, jj, ii] == 0, a0[:, :, jj, ii] == 0)


@pytest.mark.gpu
@pytest.mark.parametrize("which", ["config3", "tiles", "flat"])
def test_update_cg2d_kernels_match_the_oracle_bit_for_bit(which):
    from mitgcm_b200 import runtime as rt
    from oracle.pyoracle import Oracle
    if which == "config3":
        g, p = config3_grid(), {}
    elif which == "tiles":
        g, p = make_grid(37, 19, 3, nSx=3, nSy=2, Nr=6, seed=4), dict(deltaTFreeSurf=1200.0, deltaTMom=1200.0)
    else:
        g, p = make_grid(64, 48, 2, Nr=4, seed=5, land_frac=0.0, partial=False), dict(deltaTFreeSurf=600.0, deltaTMom=1200.0,
                                                                                     implicSurfPress=0.6, implicDiv2DFlow=0.7)
    P = _params(g, **p)
    o = Oracle(g, P)
    op = o.ini_cg2d()
    try:
        rt.init(g.d)
        rt.set_grid(g)
        rt.set_params(**{k: P[k] for k in ("deltaTMom", "deltaTFreeSurf", "implicSurfPress", "implicDiv2DFlow") if k in P})
        rt.set_cg2d_operator(op)
        rt.set_params(nIter0=0)
        for myIter, freq in ((0, 1), (3, 2), (4, 2), (5, 0)):
            _perturb_hfac(g, seed=myIter, amp=0.01 * (myIter + 1))
            o = Oracle(g, P)
            upd = freq != 0 and (myIter == 0 or myIter % freq == 0)      # update_cg2d.F:54-60
            o.update_cg2d(op, upd)
            rt.set_field("hFacW", g.a["hFacW"])
            rt.set_field("hFacS", g.a["hFacS"])
            rt.set_params(cg2dPreCondFreq=freq)
            rt.update_cg2d(myIter)
            for n in ("aW2d", "aS2d", "aC2d", "pC", "pW", "pS"):
                got = rt.get_field(n, np.zeros(g.d.shape2))
                assert np.array_equal(got, op[n]), (which, myIter, n, np.abs(got - op[n]).max())
        # the solver runs on the refreshed mirrors
        rng = np.random.default_rng(2)
        jj, ii = g.d.interior()
        b = np.zeros(g.d.shape2)
        b[:, :, jj, ii] = rng.standard_normal((g.d.nSy, g.d.nSx, g.d.sNy, g.d.sNx))
        b *= g.maskC[:, :, 0] * g.rA / P["deltaTMom"]
        bo, xo, bg, xg = b.copy(), np.zeros_like(b), b.copy(), np.zeros_like(b)
        # the oracle solver reads aW2d(i+1), aS2d(j+1) only up to sN+1: the un-exchanged halo of UPDATE_CG2D is enough
        ro = o.cg2d(op, bo, xo, 30, -1)
        rg = rt.cg2d(bg, xg, 30, -1)
        assert rg["numIters"] == ro["numIters"]
        assert np.abs(xg[:, :, jj, ii] - xo[:, :, jj, ii]).max() <= 1e-10 * np.abs(xo).max()
    finally:
        rt.finalize()
