"""End-to-end oracle run of verification/advect_cs (theta): multi-dimensional advection GAD_ADVECTION on the cs32
cubed sphere -- the three-pass, facet-dependent form (gad_advection.F:249-272, :339-812) with FILL_CS_CORNER_TR_RL /
FILL_CS_CORNER_UV_RS, scheme 33 (DST3 flux limiter), GAD_MULTIDIM_COMPRESSIBLE build (code/GAD_OPTIONS.h:44).

TEST INFRASTRUCTURE ONLY.  6 tiles of 32x32 (SIZE.h: nSx = 2, nSy = 3: the same tile numbering), OL = 4, Nr = 1,
delR = 1e5, deltaT = 2700, 192 steps, momStepping = F: the flow is the experiment's own INI_VEL (code/ini_vel.F:37-60,
solid-body rotation from the streamfunction psi = fac*fCoriG) and never changes.  Golden: %MON dynstat_theta_* every
8 steps (monitorFreq = 21600).  Per step (temp_integrate.F with tempMultiDimAdvec): GAD_ADVECTION -> gT, (GAD_CALC_RHS
adds nothing: no diffusion), TIMESTEP_TRACER without Adams-Bashforth, _EXCH_XYZ_RL(theta).
`engine`: object with gad_advection(...) as the Oracle's (the CUDA library in tests/test_gad_advection_gpu.py)."""
from __future__ import annotations

import os

import numpy as np

from mitgcm_b200.grid import Dims, cubed_sphere_grid, set_hfac
from mitgcm_b200.exch2 import cubed_sphere_topology
from . import exch2_oracle as eo
from .pyoracle import Oracle
from .baroclinic_gyre import mon_stats
from .adjustment_cs import tile_from_xstack

FIXTURE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "inputs",
                       "cs32_grid_bathy.npz")
KEEP = "xC yC rA xG yG dxC dyC dxG dyG rAw rAs".split()
RSPHERE = 6370.0e3


def tile_edges(T):
    """Per tile: 1 N | 2 S | 4 E | 8 W facet edges it touches (exch2_is{N,S,E,W}edge)."""
    tb = T.tables()
    return (np.asarray(tb["isNedge"]) * 1 + np.asarray(tb["isSedge"]) * 2 + np.asarray(tb["isEedge"]) * 4
            + np.asarray(tb["isWedge"]) * 8).astype(np.int32)


def setup(sNx=32, sNy=32):
    z = np.load(FIXTURE)
    faces = [{n: z[f"{n}_{f}"] for n in KEEP} for f in range(6)]
    T = cubed_sphere_topology(32, sNx, sNy)
    d = Dims(sNx=sNx, sNy=sNy, OLx=4, OLy=4, nSx=T.nTiles, nSy=1, Nr=1)
    g = cubed_sphere_grid(d, T, faces, [1.0e5], rotationPeriod=86164.0)
    one = np.ones(d.shape3)
    set_hfac(g, one, one.copy(), one.copy())
    theta = tile_from_xstack(T, d, z["advcs_T_init"])[:, :, None].copy()
    eo.exch2_3d(T, theta[0], d.OLx)
    # code/ini_vel.F
    omega = g.a["omega"]
    omegaprime = 38.60328935834681 / RSPHERE
    fac = -(RSPHERE * RSPHERE) * omegaprime / (2.0 * omega)
    psi = fac * g.a["fCoriG"]
    ip1 = np.minimum(np.arange(d.PX) + 1, d.PX - 1)
    jp1 = np.minimum(np.arange(d.PY) + 1, d.PY - 1)
    uVel = np.ascontiguousarray((0.0 + (psi - psi[:, :, jp1, :]) * g.a["recip_dyG"])[:, :, None])
    vVel = np.ascontiguousarray((0.0 + (psi[:, :, :, ip1] - psi) * g.a["recip_dxG"])[:, :, None])
    eo.exch2_uv_3d(T, uVel[0], vVel[0], d.OLx, True)
    uVel *= g.maskW
    vVel *= g.maskS
    return T, d, g, theta, uVel, vVel


def run(nSteps=192, every=8, engine=None, sNx=32, sNy=32):
    """Returns the monitor statistics of theta at steps 0, every, 2*every, ..."""
    T, d, g, theta, uVel, vVel = setup(sNx, sNy)
    o = Oracle(g, dict(rkSign=-1.0))
    e = engine or o
    if engine is not None and hasattr(engine, "setup"):
        engine.setup(g, o.params, None, T)
    edges = tile_edges(T)
    wVel = np.zeros(d.shape3)
    for bi in range(1, d.nSx + 1):
        o.integrate_for_w(bi, 1, uVel, vVel, wVel)
    eo.exch2_3d(T, wVel[0], d.OLx)
    dT = np.full(d.Nr, 2700.0)
    maskInC = g.maskC[:, :, 0]
    stat = lambda: mon_stats(d, theta, g.hFacC, maskInC, g.rA, g.drF)
    out = [stat()]
    ns = (d.PY, d.PX)
    for it in range(nSteps):
        new = theta.copy()
        for bi in range(1, d.nSx + 1):
            gT = np.zeros((d.Nr,) + ns)
            e.gad_advection(bi, 1, 33, 33, 0, 1, dT, uVel, vVel, wVel, theta, gT, int(T.myFace[bi - 1]), int(edges[bi - 1]))
            new[0, bi - 1] = theta[0, bi - 1] + dT[:, None, None] * gT
        theta = new
        eo.exch2_3d(T, theta[0], d.OLx)
        if (it + 1) % every == 0:
            out.append(stat())
    return out
