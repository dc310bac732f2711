"""End-to-end oracle run of verification/matrix_example (the ocean underneath the transport-matrix package, which only
records the flow): a barotropic beta-plane gyre of 32 x 32 cells of 50 km restarted from pickup.0000200000, 2 x 4 tiles of
16 x 8, OL = 3, deltaT = 20000 s, 10 steps.  TEST INFRASTRUCTURE ONLY: the step sequence and what it pins are those of
oracle/advection_in_gyre.py (advective terms on a developed flow, no-slip sides and bottom, AB2 continued from the pickup);
this experiment adds another tiling (8 tiles, odd overlap), rhoConst = rhoNil = 1035, viscAh = 5e3 and a 17-times longer time step.
Golden: results/output.txt (checkpoint63e)."""
from __future__ import annotations

import os

from . import advection_in_gyre as ag

CONFIG = dict(fixture=os.path.join(ag.INPUTS, "matrix_example.npz"), n=32, dx=50e3, tiles=(2, 4), OL=3, ygOrigin=-50e3,
              deltaT=20000.0, viscAh=5e3, viscAr=1e-2, abEps=0.1, rhoConst=1035.0, tol=1e-7)


def run(nSteps=10, engine=None):
    return ag.run(nSteps, engine, CONFIG)
