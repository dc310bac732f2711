"""Config 2 (verification/tutorial_baroclinic_gyre) stepped ENTIRELY on the device through
mitgcm_b200_forward_step_: DO_OCEANIC_PHYS (SST relaxation, linear EOS, IVDC), THERMODYNAMICS
(GAD_CALC_RHS + implicit vertical diffusion), DYNAMICS (CALC_PHI_HYD + MOM_FLUXFORM + TIMESTEP),
SOLVE_FOR_PRESSURE (exactConserv) + CG2D, correction step, INTEGR_CONTINUITY, exchanges -- against the
reference's golden output.  Pass rule = the reference's own (verification/testreport:956-987):
cg2d_init_res to >= 10 digits; iteration counts +-1; monitor statistics to the solver tolerance."""
import json
import os

import numpy as np
import pytest

from oracle import baroclinic_gyre as bc
from oracle.barotropic_gyre import tile_field

pytestmark = pytest.mark.gpu
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "tutorial_baroclinic_gyre.json")))


def test_resident_config2_matches_the_reference_golden():
    from mitgcm_b200.model import Model, ini_cg2d
    d, g, P, wind, sst = bc.setup()
    NR = d.Nr
    tRef = np.array(bc.TREF)
    P = dict(P)
    P.update(abEps=0.01, deltaTtracer=1200.0, viscAr=1e-2, diffKhT=1000.0, diffK4T=0.0, diffKrT=1e-5,
             tempAdvScheme=2, tempStepping=1, cg2dMaxIters=1000, momForcing=1, momDissip_In_AB=1,
             exactConserv=1, buoyancyLinear=1, doThetaClimRelax=1, gravity=9.81, tAlpha=2e-4, sBeta=0.0,
             rhoNil=999.8, rhoConst=999.8, ivdc_kappa=1.0)
    op = ini_cg2d(g, P)
    fu = tile_field(d, wind.astype(np.float64))
    state = dict(
        uVel=np.zeros(d.shape3), vVel=np.zeros(d.shape3), wVel=np.zeros(d.shape3),
        theta=np.where(g.maskC != 0.0, tRef[None, None, :, None, None], 0.0),
        salt=np.where(g.maskC != 0.0, 30.0, 0.0), etaN=np.zeros(d.shape2), etaH=np.zeros(d.shape2),
        surfForcU=fu * (1.0 / 999.8), surfForcV=np.zeros(d.shape2),
        SST=tile_field(d, sst.astype(np.float64)),
        lambdaThetaClimRelax=np.where(np.abs(g.yC) <= 180.0, 1.0 / 2592000.0, 0.0),
        tRef=tRef, sRef=np.full(NR, 30.0))
    m = Model(g, P, state, op, device=0)
    maskInC, maskInW, maskInS = g.maskC[:, :, 0], g.maskW[:, :, 0], g.maskS[:, :, 0]
    try:
        for it in range(10):
            r = m.step()
            assert abs(r["numIters"] - GOLD["cg2d_iters"][it]) <= 1, it
            assert r["firstResidual"] == pytest.approx(float(GOLD["cg2d_init_res"][it]), rel=1e-10), it
            st = dict(eta=bc.mon_stats(d, m.get("etaN")[:, :, None], maskInC[:, :, None], maskInC, g.rA, [g.drF[0]]),
                      uvel=bc.mon_stats(d, m.get("uVel"), g.hFacW, maskInW, g.rAw, g.drF),
                      vvel=bc.mon_stats(d, m.get("vVel"), g.hFacS, maskInS, g.rAs, g.drF),
                      wvel=bc.mon_stats(d, m.get("wVel"), g.maskC, maskInC, g.rA, g.drC[:NR]),
                      theta=bc.mon_stats(d, m.get("theta"), g.hFacC, maskInC, g.rA, g.drF))
            for s in ("max", "min", "mean", "sd"):
                assert st["theta"][s] == pytest.approx(float(GOLD[f"dynstat_theta_{s}"][it + 1]), rel=1e-10), (it, s)
            for f in ("eta", "uvel", "vvel", "wvel"):
                for s in ("max", "min", "sd"):
                    assert st[f][s] == pytest.approx(float(GOLD[f"dynstat_{f}_{s}"][it + 1]), rel=2e-6), (it, f, s)
    finally:
        m.close()
