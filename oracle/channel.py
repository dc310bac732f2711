"""Oracle model step for Nr-level Cartesian configurations with one passive/active tracer.

TEST INFRASTRUCTURE ONLY.  Same sequence as barotropic_gyre.py (which is pinned to the
reference's golden output) extended by TEMP_INTEGRATE (temp_integrate.F:296-540: CALC_ADV_FLOW,
GAD_CALC_RHS, ADAMS_BASHFORTH2, TIMESTEP_TRACER, CYCLE_TRACER) in the forward_step.F order:
THERMODYNAMICS, DYNAMICS, SOLVE_FOR_PRESSURE, MOMENTUM_CORRECTION_STEP, INTEGR_CONTINUITY,
DO_FIELDS_BLOCKING_EXCHANGES.  Used as the checker of the resident CUDA step and as the CPU
baseline of bench.py.  Tiles are processed by a thread pool (the reference's parallelism)."""
from __future__ import annotations

from concurrent.futures import ThreadPoolExecutor

import numpy as np

from .pyoracle import Oracle


class ChannelOracle:
    def __init__(self, grid, params, state, threads=1):
        """params: dict with the Oracle params plus abEps, deltaTtracer, diffKhT, diffK4T, diffKrT,
        viscAr, tempAdvScheme, tempStepping, cg2dMaxIters, momForcing, momDissip_In_AB.
        state: dict of numpy arrays uVel vVel wVel theta etaN surfForcU surfForcV (copied)."""
        self.g, self.d = grid, grid.d
        self.P = dict(abEps=0.01, deltaTtracer=params.get("deltaTMom", 1200.0), diffKhT=0.0, diffK4T=0.0,
                      diffKrT=0.0, viscAr=0.0, tempAdvScheme=2, tempStepping=1, cg2dMaxIters=150,
                      momForcing=1, momDissip_In_AB=1, useSRCGSolver=0, buoyancyLinear=0, gravity=9.81,
                      tAlpha=2e-4, sBeta=0.0, rhoNil=999.8, rhoConst=999.8, ivdc_kappa=0.0,
                      vectorInvariantMomentum=0, multiDimAdvection=1, gad_multidim_compressible=0,
                      saltStepping=0, saltAdvScheme=2, diffKhS=0.0, diffK4S=0.0, diffKrS=0.0)
        extra = {k: params[k] for k in list(params) if k in self.P}
        self.P.update(extra)
        self.o = Oracle(grid, {k: v for k, v in params.items() if k not in self.P})
        self.op = self.o.ini_cg2d()
        d = self.d
        self.s = {k: np.ascontiguousarray(v, dtype=np.float64).copy() for k, v in state.items()}
        for n in ("gU", "gV", "guNm1", "gvNm1", "gtNm1", "gsNm1"):
            self.s[n] = np.zeros(d.shape3)
        if self.P["saltStepping"] and "salt" not in self.s:
            self.s["salt"] = np.zeros(d.shape3)
        self.kapS = np.full((d.PY, d.PX), float(self.P["diffKrS"]))
        self.kapU = np.full((d.Nr + 1, d.PY, d.PX), float(self.P["viscAr"]))
        self.kapT = np.full((d.PY, d.PX), float(self.P["diffKrT"]))
        self.threads = threads
        self.it = 0
        # coupled buoyancy (eosType = 'LINEAR'): FIND_RHO_2D before the thermodynamics, CALC_PHI_HYD in DYNAMICS
        if self.P["buoyancyLinear"]:
            from .pyoracle import Eos
            self.eos = Eos(self.P["rhoNil"], self.P["rhoConst"], self.P["tAlpha"], self.P["sBeta"])
            self.tRef = np.ascontiguousarray(self.s.pop("tRef"), dtype=np.float64)
            self.sRef = np.ascontiguousarray(self.s.pop("sRef", np.zeros(d.Nr)), dtype=np.float64)
            if "salt" not in self.s:
                self.s["salt"] = np.zeros(d.shape3)
            self.rho, self.ivdc = np.zeros(d.shape3), np.zeros(d.shape3)
            self.phi0 = np.zeros(d.shape2)

    def _tiles(self):
        return [(bi, bj) for bj in range(1, self.d.nSy + 1) for bi in range(1, self.d.nSx + 1)]

    def _map(self, fn):
        if self.threads > 1 and len(self._tiles()) > 1:
            with ThreadPoolExecutor(self.threads) as ex:
                list(ex.map(fn, self._tiles()))
        else:
            for t in self._tiles():
                fn(t)

    def step(self):
        d, o, s, P = self.d, self.o, self.s, self.P
        ns = (d.PY, d.PX)
        abFac = 0.0 if self.it == 0 else 0.5 + P["abEps"]
        dT = np.full(d.Nr, float(P["deltaTtracer"]))
        zr = np.zeros(d.Nr)

        def thermo(t, trc="theta", gNm1name="gtNm1", schName="tempAdvScheme", kh="diffKhT", k4="diffK4T", kr="diffKrT", kap2d=None):
            bi, bj = t
            kap2d = self.kapT if kap2d is None else kap2d
            ti = (bj - 1, bi - 1)
            gT = np.zeros((d.Nr,) + ns)
            fV = np.zeros((2,) + ns)
            rTrans = np.zeros(ns)
            sl = {n: np.zeros(ns) for n in "xA yA maskUp uFld vFld wFld uTrans vTrans rTransKp1 fZon fMer".split()}
            theta = np.ascontiguousarray(s[trc][ti])
            # gad_init_fixed.F:100-131: AB on gT and the 1-D advection only for schemes 2, 3, 4
            sch = P[schName]
            abScheme = sch in (2, 3, 4)
            multiDim = bool(P["multiDimAdvection"]) and not abScheme
            if multiDim:       # temp_integrate.F:276-290: GAD_ADVECTION gives the advective tendency
                o.gad_advection(bi, bj, sch, sch, 0, P["gad_multidim_compressible"], dT, s["uVel"], s["vVel"], s["wVel"],
                                s[trc], gT)
            kapK = None
            if P["buoyancyLinear"] and P["ivdc_kappa"] != 0.0:     # CALC_3D_DIFFUSIVITY with the convective flag
                kapK = np.zeros((d.Nr,) + ns)
                o.calc_3d_diffusivity(bi, bj, self.ivdc, float(P["ivdc_kappa"]), zr, np.full(d.Nr, float(P[kr])), kapK)
            for k in range(d.Nr, 0, -1):
                kUp, kDown = 1 + (k + 1) % 2, 1 + k % 2
                o.calc_adv_flow(bi, bj, k, s["uVel"], s["vVel"], s["wVel"], sl["xA"], sl["yA"], sl["maskUp"],
                                sl["uFld"], sl["vFld"], sl["wFld"], sl["uTrans"], sl["vTrans"], rTrans, sl["rTransKp1"])
                o.gad_calc_rhs(bi, bj, 0, d.sNx + 1, 0, d.sNy + 1, k, max(1, k - 1), kUp, kDown, sl["xA"], sl["yA"],
                               sl["maskUp"], sl["uFld"], sl["vFld"], sl["wFld"], sl["uTrans"], sl["vTrans"], rTrans,
                               sl["rTransKp1"], P[kh], P[k4], kap2d if kapK is None else kapK[k - 1], zr, theta, theta, dT,
                               sch, sch, 0 if multiDim else 1, 0, 0, 0, sl["fZon"], sl["fMer"], fV, gT)
                if abScheme:
                    # ADAMS_BASHFORTH2 (adams_bashforth2.F:84-86)
                    gNm1 = s[gNm1name][ti][k - 1]
                    ab = abFac * (gT[k - 1] - gNm1)
                    gNm1[...] = gT[k - 1]
                    gT[k - 1] = gT[k - 1] + ab
            # TIMESTEP_TRACER + CYCLE_TRACER
            s[trc][ti] = theta + dT[:, None, None] * gT

        def thermo_salt(t):
            thermo(t, "salt", "gsNm1", "saltAdvScheme", "diffKhS", "diffK4S", "diffKrS", self.kapS)

        buoy = bool(P["buoyancyLinear"])

        def dyn(t):
            bi, bj = t
            fU, fVv = np.zeros((2,) + ns), np.zeros((2,) + ns)
            z = np.zeros(ns)
            zy = z
            if buoy:
                phiF, phiC, z, zy = (np.zeros(ns) for _ in range(4))
            for k in range(1, d.Nr + 1):
                kUp, kDown = 1 + (k + 1) % 2, 1 + k % 2
                gd, hd = np.zeros(ns), np.zeros(ns)
                if buoy:
                    o.calc_phi_hyd(bi, bj, 0, d.sNx + 1, 0, d.sNy + 1, k, self.rho, self.g.rF, self.g.rC,
                                   P["gravity"], 1.0 / P["rhoConst"], self.phi0, phiF, phiC, z, zy)
                # dynamics.F:517-533: MOM_FLUXFORM or, with vectorInvariantMomentum, MOM_VECINV
                mom = o.mom_vecinv if P["vectorInvariantMomentum"] else o.mom_fluxform
                mom(bi, bj, k, 0, d.sNx + 1, 0, d.sNy + 1, self.kapU, self.kapU, fU[kUp - 1], fVv[kUp - 1],
                    fU[kDown - 1], fVv[kDown - 1], gd, hd, s["uVel"], s["vVel"], s["wVel"], s["gU"], s["gV"])
                o.timestep(bi, bj, k, 0, d.sNx + 1, 0, d.sNy + 1, z, zy, gd, hd, s["surfForcU"], s["surfForcV"],
                           P["momForcing"], P["momDissip_In_AB"], abFac, s["uVel"], s["vVel"], s["gU"], s["gV"],
                           s["guNm1"], s["gvNm1"])
            if o.params["implicitViscosity"] and o.params["momViscosity"]:      # dynamics.F:572-579
                o.mom_implicit_r(bi, bj, 0, self.kapU, s["gU"])
                o.mom_implicit_r(bi, bj, 1, self.kapU, s["gV"])

        if buoy:      # DO_OCEANIC_PHYS: density of theta(n), before the thermodynamics updates theta
            self._map(lambda t: o.density_ivdc(self.eos, t[0], t[1], s["theta"], s["salt"], self.tRef, self.sRef,
                                               self.rho, self.ivdc))
        if P["tempStepping"]:
            self._map(thermo)
        if P["saltStepping"]:
            self._map(thermo_salt)
        self._map(dyn)
        b, x = np.zeros(d.shape2), np.zeros(d.shape2)
        self._map(lambda t: o.solve_rhs(t[0], t[1], s["etaN"], s["gU"], s["gV"], b, x))
        res = o.cg2d(self.op, b, x, int(P["cg2dMaxIters"]), -1, sr=bool(P["useSRCGSolver"]))
        o.exch_xyz(x)
        s["etaN"] = self.g.recip_Bo * x

        def corr(t):
            o.correction_step(t[0], t[1], s["etaN"], s["gU"], s["gV"], s["uVel"], s["vVel"])
            o.integrate_for_w(t[0], t[1], s["uVel"], s["vVel"], s["wVel"])
        self._map(corr)
        for n in ("uVel", "vVel", "wVel"):
            o.exch_xyz(s[n], d.Nr)
        if P["tempStepping"]:
            o.exch_xyz(s["theta"], d.Nr)
        if P["saltStepping"]:
            o.exch_xyz(s["salt"], d.Nr)
        self.it += 1
        return res
