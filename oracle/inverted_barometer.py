"""End-to-end oracle run of verification/inverted_barometer: a closed, stratified f-plane box (60 x 60 cells of 100 km,
4 levels of 500 m, 2 x 2 tiles of 30 x 30, OL = 2) at rest under an atmospheric pressure load (pLoadFile), 40 steps.

TEST INFRASTRUCTURE ONLY.  Pins, against the experiment's golden output (results/output.txt), the surface boundary value of
the hydrostatic potential -- phi0surf = pLoad / rhoConst (external_forcing_surf.F:352-374), which CALC_GRAD_PHI_HYD adds
to phiHyd (calc_grad_phi_hyd.F) and which drives the adjustment eta -> -pLoad / (rhoConst g) -- together with the no-slip
bottom at Nr = 4 (viscAz = 0.01) and MOM_FLUXFORM / GAD_CALC_RHS on a multi-level Cartesian f-plane.
Parameters from input/data: deltaT = 1200, abEps = 0.1, viscAh = 400, viscAz = 0.01, free-slip sides, no-slip bottom,
diffKhT = 400, diffKzT = 0.01, f0 = 1e-4, beta = 0, tAlpha = 2e-4, gravity = 9.81, rhoConst = 999.8,
cg2dTargetResidual = 1e-13; salinity is uniform with sBeta = 0 and is not stepped.  Shares the step loop of
oracle/flt_example.py; `engine` as in baroclinic_gyre.py.
"""
from __future__ import annotations

import os

import numpy as np

from mitgcm_b200.grid import Dims, cartesian_grid, masks_from_depth, global_area
from .barotropic_gyre import tile_field
from .flt_example import step_loop

FIXTURE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "inputs",
                       "inverted_barometer.npz")
TREF = [20.0, 10.0, 8.0, 6.0]


def setup():
    z = np.load(FIXTURE)
    d = Dims(sNx=30, sNy=30, OLx=2, OLy=2, nSx=2, nSy=2, Nr=4)
    g = cartesian_grid(d, [100e3] * 60, [100e3] * 60, [500.0] * 4, f0=1e-4, beta=0.0, gBaro=9.81)
    masks_from_depth(g, z["topog"], hFacMin=1.0, hFacMinDr=0.0)
    P = dict(deltaTMom=1200.0, deltaTFreeSurf=1200.0, viscAhD=400.0, viscAhZ=400.0, no_slip_sides=0, sideDragFactor=2.0,
             no_slip_bottom=1, selectBotDragQuadr=-1, implicitDiffusion=0, cg2dTargetResidual=1e-13,
             globalArea=global_area(g))
    return z, d, g, P


def run(nSteps=40, engine=None):
    """Returns (cg2dNorm, statistics of the start state, [per-step dict])."""
    z, d, g, P = setup()
    rhoConst = 999.8
    return step_loop(d, g, P, nSteps, engine, tRef=TREF, salt0=10.0, rhoConst=rhoConst, tAlpha=2e-4, gravity=9.81, abEps=0.1,
                     viscAr=1e-2, diffKhT=400.0, diffKrT=1e-2, deltaT=1200.0, sfU=np.zeros(d.shape2),
                     phi0surf=tile_field(d, z["pLoad"]) * (1.0 / rhoConst))
