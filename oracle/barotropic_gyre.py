"""End-to-end oracle run of verification/tutorial_barotropic_gyre (config 1).

TEST INFRASTRUCTURE ONLY.  Drives the C restatements (cg2d, mom_fluxform, step
glue) in the order of model/src/forward_step.F (non-staggered) so that the
oracle can be PINNED against the reference's golden output
verification/tutorial_barotropic_gyre/results/output.txt: cg2dNorm (:373),
per-step `cg2d: Sum(rhs),rhsMax`, cg2d_init_res, cg2d_iters, cg2d_last_res
(:1323-1326 ...) and %MON dynstat_{eta,uvel,vvel}_{max,min,mean,sd}.

Inputs are regenerated from the formulas of input/gendata.py (bathy.bin,
windx_cosy.bin: float32); tests/test_oracle_golden.py checks they are
byte-identical to the reference's files when /root/reference is present.
"""
from __future__ import annotations

import numpy as np

from mitgcm_b200.grid import Dims, cartesian_grid, masks_from_depth, exch_xyz, global_area
from .pyoracle import Oracle

NX = NY = 62


def gen_inputs():
    """input/gendata.py: walls of zero depth around a 5000 m flat basin; zonal
    wind stress -tauMax*cos(pi*y) at u-points. Both stored as float32."""
    h = -5000.0 * np.ones((NY, NX))
    h[:, [0, -1]] = 0
    h[[0, -1], :] = 0
    y = (np.arange(NY) - .5) / (NY - 2)
    x = (np.arange(NX) - 1) / (NX - 2)
    Y, _ = np.meshgrid(y, x, indexing='ij')
    tau = -0.1 * np.cos(Y * np.pi)
    return h.astype('>f4'), tau.astype('>f4')


def tile_field(d: Dims, glob: np.ndarray) -> np.ndarray:
    """global (Ny,Nx) -> tiled array with halos filled by EXCH_XY."""
    a = np.zeros(d.shape2)
    for bj in range(d.nSy):
        for bi in range(d.nSx):
            a[bj, bi, d.OLy:d.OLy + d.sNy, d.OLx:d.OLx + d.sNx] = \
                glob[bj * d.sNy:(bj + 1) * d.sNy, bi * d.sNx:(bi + 1) * d.sNx]
    return exch_xyz(d, a)


def mon_stats(d, arr, hfac, mask2, area, dr):
    """pkg/monitor/mon_calc_stats_rl.F (min, max, mean, sd); sequential sums in
    the reference's order (k, j, i inside each tile, tiles in order)."""
    jj, ii = d.interior()
    a = arr[:, :, :, jj, ii]
    m = mask2[:, :, None, jj, ii] * hfac[:, :, :, jj, ii]
    vol = area[:, :, None, jj, ii] * dr[None, None, :, None, None] * m
    sel = m > 0
    mn, mx = a[sel].min(), a[sel].max()

    def tsum(x):
        tot = 0.0
        for bj in range(d.nSy):
            for bi in range(d.nSx):
                s = 0.0
                for v in x[bj, bi][sel[bj, bi]]:
                    s = s + v
                tot = tot + s
        return tot
    theVol = tsum(vol)
    mean = tsum(vol * a) / theVol
    sd = np.sqrt(tsum(vol * (a - mean) * (a - mean)) / theVol)
    return dict(min=mn, max=mx, mean=mean, sd=sd)


def run(nSteps=10, nSx=1, nSy=1, sr=False, cg2d_fn=None, collect_inputs=False):
    """Returns (cg2dNorm, [per-step dict]).  cg2d_fn, if given, replaces the
    oracle's CG2D (signature of Oracle.cg2d) so a test can put the CUDA solver in
    the loop; it must not be used to claim oracle parity."""
    d = Dims(sNx=NX // nSx, sNy=NY // nSy, OLx=2, OLy=2, nSx=nSx, nSy=nSy, Nr=1)
    g = cartesian_grid(d, [20e3] * NX, [20e3] * NY, [5000.0], xgOrigin=-20e3, ygOrigin=-20e3,
                       f0=1e-4, beta=1e-11, gBaro=9.81)
    bathy, wind = gen_inputs()
    masks_from_depth(g, bathy.astype(np.float64), hFacMin=1.0, hFacMinDr=0.0)
    o = Oracle(g, dict(deltaTMom=1200.0, deltaTFreeSurf=1200.0, viscAhD=400.0, viscAhZ=400.0,
                       no_slip_sides=1, sideDragFactor=2.0, no_slip_bottom=1, selectBotDragQuadr=-1,
                       cg2dTargetResidual=1e-7, globalArea=global_area(g)))
    op = o.ini_cg2d()
    fu = tile_field(d, wind.astype(np.float64))
    mass2rUnit = 1.0 / 1000.0
    sfU = fu * mass2rUnit
    sfV = np.zeros(d.shape2)
    z3 = lambda: np.zeros(d.shape3)
    uVel, vVel, wVel, gU, gV, guNm1, gvNm1 = (z3() for _ in range(7))
    etaN = np.zeros(d.shape2)
    kappaR = np.zeros((d.Nr + 1, d.PY, d.PX))
    dPhi = np.zeros((d.PY, d.PX))
    abEps = 0.01
    out = []
    maskInC = g.maskC[:, :, 0]
    maskInW = g.maskW[:, :, 0]
    maskInS = g.maskS[:, :, 0]
    inputs = []
    for it in range(nSteps):
        abFac = 0.0 if it == 0 else 0.5 + abEps
        for bj in range(1, d.nSy + 1):
            for bi in range(1, d.nSx + 1):
                fVerU = np.zeros((2, d.PY, d.PX))
                fVerV = np.zeros((2, d.PY, d.PX))
                for k in range(1, d.Nr + 1):
                    kUp, kDown = 1 + (k + 1) % 2, 1 + k % 2
                    guDiss, gvDiss = np.zeros((d.PY, d.PX)), np.zeros((d.PY, d.PX))
                    o.mom_fluxform(bi, bj, k, 0, d.sNx + 1, 0, d.sNy + 1, kappaR, kappaR,
                                   fVerU[kUp - 1], fVerV[kUp - 1], fVerU[kDown - 1], fVerV[kDown - 1],
                                   guDiss, gvDiss, uVel, vVel, wVel, gU, gV)
                    o.timestep(bi, bj, k, 0, d.sNx + 1, 0, d.sNy + 1, dPhi, dPhi, guDiss, gvDiss, sfU, sfV,
                               1, 1, abFac, uVel, vVel, gU, gV, guNm1, gvNm1)
        b, x = np.zeros(d.shape2), np.zeros(d.shape2)
        for bj in range(1, d.nSy + 1):
            for bi in range(1, d.nSx + 1):
                o.solve_rhs(bi, bj, etaN, gU, gV, b, x)
        if collect_inputs:
            inputs.append((b.copy(), x.copy()))
        res = (cg2d_fn or o.cg2d)(op, b, x, 1000, -1, sr=sr)
        o.exch_xyz(x)
        etaN = g.recip_Bo * x
        for bj in range(1, d.nSy + 1):
            for bi in range(1, d.nSx + 1):
                o.correction_step(bi, bj, etaN, gU, gV, uVel, vVel)
                o.integrate_for_w(bi, bj, uVel, vVel, wVel)
        o.exch_xyz(uVel, d.Nr)
        o.exch_xyz(vVel, d.Nr)
        o.exch_xyz(wVel, d.Nr)
        rec = dict(res)
        rec["eta"] = mon_stats(d, etaN[:, :, None], maskInC[:, :, None], maskInC, g.rA, g.drF)
        rec["uvel"] = mon_stats(d, uVel, g.hFacW, maskInW, g.rAw, g.drF)
        rec["vvel"] = mon_stats(d, vVel, g.hFacS, maskInS, g.rAs, g.drF)
        out.append(rec)
    if collect_inputs:
        return op, out, inputs, o
    return op["cg2dNorm"], out
