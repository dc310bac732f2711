"""Shared test set-up: random-bathymetry Cartesian grids and CG2D problems."""
import numpy as np

from mitgcm_b200.grid import Dims, cartesian_grid, masks_from_depth, exch_xyz, global_area
from oracle.pyoracle import Oracle


def make_grid(sNx, sNy, OL, nSx=1, nSy=1, Nr=1, seed=0, land_frac=0.15, partial=True, dx=20e3, dz=None):
    d = Dims(sNx=sNx, sNy=sNy, OLx=OL, OLy=OL, nSx=nSx, nSy=nSy, Nr=Nr)
    rng = np.random.default_rng(seed)
    delX = dx * (1.0 + 0.2 * rng.random(d.Nx))
    delY = dx * (1.0 + 0.2 * rng.random(d.Ny))
    delR = np.asarray(dz if dz is not None else 100.0 * (1.0 + 0.5 * np.arange(Nr)))
    g = cartesian_grid(d, delX, delY, delR, f0=1e-4, beta=1e-11)
    H = delR.sum()
    depth = -H * (0.3 + 0.7 * rng.random((d.Ny, d.Nx))) if partial else -H * np.ones((d.Ny, d.Nx))
    land = rng.random((d.Ny, d.Nx)) < land_frac
    depth[land] = 0.0
    masks_from_depth(g, depth, hFacMin=0.2 if partial else 1.0, hFacMinDr=0.0)
    return g


def cg2d_problem(g, seed=1, tol=1e-9, **params):
    """Operator from INI_CG2D (oracle) on grid g plus a random RHS with zero mean on wet points
    and a random first guess.  Returns (oracle, op, b, x)."""
    d = g.d
    p = dict(deltaTMom=1200.0, deltaTFreeSurf=1200.0, cg2dTargetResidual=tol, globalArea=global_area(g))
    p.update(params)
    o = Oracle(g, p)
    op = o.ini_cg2d()
    rng = np.random.default_rng(seed)
    jj, ii = d.interior()
    wet = g.maskC[:, :, 0]
    b = np.zeros(d.shape2)
    b[:, :, jj, ii] = rng.standard_normal((d.nSy, d.nSx, d.sNy, d.sNx))
    b *= wet * g.rA / 1200.0
    x = 0.1 * rng.standard_normal(d.shape2) * wet
    return o, op, b, x
