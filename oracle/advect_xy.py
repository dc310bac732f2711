"""Oracle run of verification/advect_xy (salt only): multi-dimensional advection (GAD_ADVECTION,
saltAdvScheme = 33, built with GAD_MULTIDIM_COMPRESSIBLE) of a blob by the uniform flow u = v = 1 m/s on
a doubly periodic 20 x 20 x 1 Cartesian grid (2 tiles of 20 x 10, OL = 3), momStepping = F, 80 steps.

TEST INFRASTRUCTURE ONLY.  Pins GAD_ADVECTION and the DST3 flux-limiter leaves (gad_dst3fl_adv_{x,y}.F,
shared with GAD_CALC_RHS) against the experiment's golden output (results/output.txt:
%MON dynstat_salt_{max,min,mean,sd} at steps 0, 16, ..., 80).  The experiment's temperature uses the
second-order-moment scheme (80), which is outside the path and is not run.
Set-up read from input/data, code/SIZE.h, code/ini_salt.F, code/ini_vel.F: dXspacing = dYspacing = 10 km,
delZ = 10 km, deltaT = 2500 s, salt = 35 + 1 inside the 60 km sphere around (40 km, 40 km, -50 km).
`advect` replaces the oracle's GAD_ADVECTION (same signature) so a GPU test can put the CUDA kernel in."""
from __future__ import annotations

import numpy as np

from mitgcm_b200.grid import Dims, cartesian_grid, masks_from_depth
from .pyoracle import Oracle
from .baroclinic_gyre import mon_stats


def setup(nSx=1, nSy=2):
    d = Dims(sNx=20 // nSx, sNy=20 // nSy, OLx=3, OLy=3, nSx=nSx, nSy=nSy, Nr=1)
    g = cartesian_grid(d, [10e3] * 20, [10e3] * 20, [10e3], f0=0.0, beta=0.0)
    masks_from_depth(g, -10e3 * np.ones((20, 20)), hFacMin=1.0, hFacMinDr=0.0)
    xC, yC, rC = g.a["xC"], g.a["yC"], g.a["rC"]
    rD = np.sqrt((xC - 40e3) ** 2 + (yC - 40e3) ** 2 + (rC[0] + 50e3) ** 2)
    salt = np.where(rD <= 60e3, 36.0, 35.0)[:, :, None].copy()
    return d, g, salt


def run(nSteps=80, scheme=33, compressible=True, advect=None, nSx=1, nSy=2, every=16):
    """Returns the list of monitor statistics of salt at steps 0, every, 2*every, ..."""
    d, g, salt = setup(nSx, nSy)
    o = Oracle(g, {})
    o.exch_xyz(salt, d.Nr)
    uVel = np.ones(d.shape3) * g.maskW
    vVel = np.ones(d.shape3) * g.maskS
    wVel = np.zeros(d.shape3)
    dT = np.full(d.Nr, 2500.0)
    maskInC = g.maskC[:, :, 0]
    fn = advect or o.gad_advection
    out = [mon_stats(d, salt, g.hFacC, maskInC, g.rA, g.drF)]
    for it in range(nSteps):
        new = salt.copy()
        for bj in range(1, d.nSy + 1):
            for bi in range(1, d.nSx + 1):
                gS = np.zeros((d.Nr, d.PY, d.PX))
                rc = fn(bi, bj, scheme, scheme, 0, int(compressible), dT, uVel, vVel, wVel, salt, gS)
                assert not rc
                new[bj - 1, bi - 1] = salt[bj - 1, bi - 1] + dT[:, None, None] * gS     # TIMESTEP_TRACER
        salt = new
        o.exch_xyz(salt, d.Nr)
        if (it + 1) % every == 0:
            out.append(mon_stats(d, salt, g.hFacC, maskInC, g.rA, g.drF))
    return out


# ---------------------------------------------------------------------------------------------------------
# input.ab3_c4 (results/output.ab3_c4.txt): the same flow with tempAdvScheme = saltAdvScheme = 4 (centred 4th order,
# gad_c4_adv_{x,y}.F through GAD_CALC_RHS) stepped with the third-order Adams-Bashforth scheme on the tendencies
# (ALLOW_ADAMSBASHFORTH_3 in code/CPP_OPTIONS.h, doAB_onGtGs = T, alph_AB = 0.5, beta_AB = 0.281105, deltaT = 2750 s,
# 100 steps, monitor every 10).  theta = exp(-(rD / 20 km)^2 / 2) (code/ini_theta.F), salt as above.
def ab3_factors(myIter, nIter0, startAB, alph_AB, beta_AB):
    """adams_bashforth3.F:66-83."""
    if myIter == nIter0 and startAB == 0:
        return 0.0, 0.0, 0.0
    if (myIter == nIter0 and startAB == 1) or (myIter == 1 + nIter0 and startAB == 0):
        return alph_AB, -alph_AB, 0.0
    return alph_AB + beta_AB, -alph_AB - 2.0 * beta_AB, beta_AB


def adams_bashforth3(gT_k, gTrNm_k, myIter, nIter0, startAB, alph_AB, beta_AB):
    """ADAMS_BASHFORTH3 on one level in the tendency form (kArg = k; adams_bashforth3.F:100-113): gTrNm_k is the
    (2, PY, PX) pair of stored tendencies of that level; both arguments are updated in place."""
    m1, m2 = (myIter + 1) % 2, myIter % 2          # 0-based: m1 = 1 + MOD(myIter+1, 2), m2 = 1 + MOD(myIter, 2)
    ab0, ab1, ab2 = ab3_factors(myIter, nIter0, startAB, alph_AB, beta_AB)
    ab = ab0 * gT_k + ab1 * gTrNm_k[m1] + ab2 * gTrNm_k[m2]
    gTrNm_k[m2] = gT_k
    gT_k += ab


def setup_ab3(nSx=1, nSy=2):
    d, g, salt = setup(nSx, nSy)
    xC, yC, rC = g.a["xC"], g.a["yC"], g.a["rC"]
    rD = np.sqrt((xC - 40e3) ** 2 + (yC - 40e3) ** 2 + (rC[0] + 50e3) ** 2)
    theta = np.exp(-0.5 * (rD / 20e3) ** 2)[:, :, None].copy()
    return d, g, theta, salt


def run_ab3(nSteps=100, scheme=4, calc_rhs=None, nSx=1, nSy=2, every=10, alph_AB=0.5, beta_AB=0.281105, deltaT=2750.0):
    """Returns [(theta stats, salt stats)] at steps 0, every, 2*every, ...  `calc_rhs` replaces the oracle's
    GAD_CALC_RHS (same signature) so a GPU test can put the CUDA kernel in the loop."""
    d, g, theta, salt = setup_ab3(nSx, nSy)
    o = Oracle(g, {})
    ns = (d.PY, d.PX)
    trc = {"theta": theta, "salt": salt}
    for a in trc.values():
        o.exch_xyz(a, d.Nr)
    uVel = np.ones(d.shape3) * g.maskW
    vVel = np.ones(d.shape3) * g.maskS
    wVel = np.zeros(d.shape3)
    gNm = {n: np.zeros((d.nSy, d.nSx, d.Nr, 2) + ns) for n in trc}      # gtNm / gsNm (.., k, bi, bj, 2)
    dT = np.full(d.Nr, deltaT)
    zr = np.zeros(d.Nr)
    kap = np.zeros(ns)
    maskInC = g.maskC[:, :, 0]
    fn = calc_rhs or o.gad_calc_rhs
    stats = lambda: tuple(mon_stats(d, trc[n], g.hFacC, maskInC, g.rA, g.drF) for n in ("theta", "salt"))
    out = [stats()]
    for it in range(nSteps):
        for n in ("theta", "salt"):
            new = trc[n].copy()
            for bj in range(1, d.nSy + 1):
                for bi in range(1, d.nSx + 1):
                    ti = (bj - 1, bi - 1)
                    gT = np.zeros((d.Nr,) + ns)
                    fV = np.zeros((2,) + ns)
                    rTrans = np.zeros(ns)
                    sl = {m: np.zeros(ns) for m in "xA yA maskUp uFld vFld wFld uTrans vTrans rTransKp1 fZon fMer".split()}
                    cur = np.ascontiguousarray(trc[n][ti])
                    for k in range(d.Nr, 0, -1):
                        kUp, kDown = 1 + (k + 1) % 2, 1 + k % 2
                        o.calc_adv_flow(bi, bj, k, uVel, vVel, wVel, sl["xA"], sl["yA"], sl["maskUp"], sl["uFld"], sl["vFld"],
                                        sl["wFld"], sl["uTrans"], sl["vTrans"], rTrans, sl["rTransKp1"])
                        fn(bi, bj, 0, d.sNx + 1, 0, d.sNy + 1, k, max(1, k - 1), kUp, kDown, sl["xA"], sl["yA"], sl["maskUp"],
                           sl["uFld"], sl["vFld"], sl["wFld"], sl["uTrans"], sl["vTrans"], rTrans, sl["rTransKp1"], 0.0, 0.0, kap,
                           zr, cur, cur, dT, scheme, scheme, 1, 0, 0, 0, sl["fZon"], sl["fMer"], fV, gT)
                        adams_bashforth3(gT[k - 1], gNm[n][ti][k - 1], it, 0, 0, alph_AB, beta_AB)     # temp_integrate.F:382-388
                    new[ti] = cur + dT[:, None, None] * gT          # TIMESTEP_TRACER + CYCLE_TRACER
            trc[n][...] = new
            o.exch_xyz(trc[n], d.Nr)
        if (it + 1) % every == 0:
            out.append(stats())
    return out
