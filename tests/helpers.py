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


class CudaEngine:
    """Adapter with the Oracle's method signatures (mom_fluxform, gad_calc_rhs, cg2d) that routes to the
    CUDA library through mitgcm_b200.runtime, i.e. through the C ABI with host buffers and the reference
    argument lists.  Lets the end-to-end drivers under oracle/ run with the CUDA kernels in the loop."""

    def __init__(self, rt, use_gad=True, use_mom=True, use_cg2d=True, fallback=None, use_cg3d=True):
        self.rt, self.fb = rt, fallback
        self.use_gad, self.use_mom, self.use_cg2d, self.use_cg3d = use_gad, use_mom, use_cg2d, use_cg3d

    def setup(self, g, params, op, topo=None, op3=None):
        from mitgcm_b200 import _lib
        rt = self.rt
        rt.init(g.d)
        rt.set_grid(g)
        if topo is not None:            # pkg/exch2 tile graph (cubed sphere)
            from mitgcm_b200.exch2 import set_topology
            set_topology(topo)
        known = {k: v for k, v in params.items() if "MP_" + k.upper() in _lib.ENUMS or "MI_" + k.upper() in _lib.ENUMS}
        rt.set_params(**known)
        rt.set_cg2d_operator(op)
        if op3 is not None:
            rt.set_cg3d_operator(op3)

    def gad_calc_rhs(self, bi, bj, iMin, iMax, jMin, jMax, k, kM1, kUp, kDown, xA, yA, maskUp, uFld, vFld, wFld,
                     uTrans, vTrans, rTrans, rTransKp1, diffKh, diffK4, KappaR, diffKr4, TracerN, TracAB, deltaTLev,
                     advScheme, vertAdvScheme, calcAdvection, implicitAdvection, applyAB_onTracer, trUseDiffKr4,
                     fZon, fMer, fVerT, gTracer):
        a = (bi, bj, iMin, iMax, jMin, jMax, k, kM1, kUp, kDown, xA, yA, maskUp, uFld, vFld, wFld, uTrans, vTrans,
             rTrans, rTransKp1, diffKh, diffK4, KappaR, diffKr4, TracerN, TracAB, deltaTLev)
        if not self.use_gad:
            return self.fb.gad_calc_rhs(*a, advScheme, vertAdvScheme, calcAdvection, implicitAdvection,
                                        applyAB_onTracer, trUseDiffKr4, fZon, fMer, fVerT, gTracer)
        self.rt.gad_calc_rhs(*a, 1, advScheme, vertAdvScheme, calcAdvection, implicitAdvection, applyAB_onTracer,
                             trUseDiffKr4, 0, 0, 0, fZon, fMer, fVerT, gTracer)

    def mom_fluxform(self, *a):
        return (self.rt if self.use_mom else self.fb).mom_fluxform(*a)

    def mom_vecinv(self, *a):
        return (self.rt if self.use_mom else self.fb).mom_vecinv(*a)

    def cg2d(self, op, b, x, numIters, nIterMin=-1, sr=False):
        if not self.use_cg2d:
            return self.fb.cg2d(op, b, x, numIters, nIterMin, sr=sr)
        return self.rt.cg2d(b, x, numIters, nIterMin, sr=sr)

    def cg3d(self, op3, b, x, numIters):
        if not self.use_cg3d:
            return self.fb.cg3d(op3, b, x, numIters)
        return self.rt.cg3d(b, x, numIters)


def load_cs32():
    """Config 4 geometry from the committed fixture (tests/golden/inputs/cs32_grid_bathy.npz, made by
    tests/golden/make_input_fixtures.py from the reference's grid_cs32.face00N.bin and bathy_Hmin50.bin):
    returns (topology, grid with masks, params)."""
    import os
    from mitgcm_b200.grid import cubed_sphere_grid, cube_masks_from_depth
    from mitgcm_b200.exch2 import cubed_sphere_topology
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "inputs", "cs32_grid_bathy.npz"))
    keep = "xC yC rA xG yG dxC dyC dxG dyG rAw rAs".split()
    faces = [{n: z[f"{n}_{f}"] for n in keep} for f in range(6)]
    T = cubed_sphere_topology(32, 32, 16)
    d = Dims(sNx=32, sNy=16, OLx=4, OLy=4, nSx=12, nSy=1, Nr=15)
    delR = [50., 70., 100., 140., 190., 240., 290., 340., 390., 440., 490., 540., 590., 640., 690.]
    g = cubed_sphere_grid(d, T, faces, delR)
    cube_masks_from_depth(g, T, z["bathy_Hmin50"], hFacMin=0.1, hFacMinDr=20.0)
    P = dict(deltaTMom=1200.0, deltaTFreeSurf=86400.0, cg2dTargetResWunit=1e-14, globalArea=global_area(g))
    return T, g, P
