"""ctypes binding of the CPU oracle (oracle/_ref/liboracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs, never by mitgcm_b200.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_ref", "liboracle.so")
P = C.POINTER(C.c_double)


def build(force=False):
    srcs = [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith((".c", ".h"))]
    if force or not os.path.exists(LIB) or any(os.path.getmtime(s) > os.path.getmtime(LIB) for s in srcs):
        subprocess.check_call(["make", "-C", HERE], stdout=subprocess.DEVNULL)
    return LIB


class Dims(C.Structure):
    _fields_ = [(n, C.c_int) for n in "sNx sNy OLx OLy nSx nSy Nr".split()]


GRID_FIELDS = ("dxC dyC dxG dyG dxF dyF dxV dyU rA rAw rAs rAz "
               "recip_dxC recip_dyC recip_dxG recip_dyG recip_dxF recip_dyF recip_dxV recip_dyU "
               "recip_rA recip_rAw recip_rAs recip_rAz fCori fCoriG tanPhiAtU tanPhiAtV "
               "cosFacU cosFacV drF drC recip_drF recip_drC "
               "hFacC hFacW hFacS recip_hFacC recip_hFacW recip_hFacS maskC maskW maskS recip_Bo").split()


class GridS(C.Structure):
    _fields_ = [("d", Dims)] + [(n, P) for n in GRID_FIELDS]


PARAM_D = ("deltaTMom deltaTFreeSurf freeSurfFac implicSurfPress implicDiv2DFlow rkSign "
           "cg2dpcOffDFac cg2dTargetResidual cg2dTargetResWunit globalArea "
           "viscAhD viscAhZ viscA4D viscA4Z sideDragFactor bottomDragLinear bottomDragQuadratic "
           "recip_rSphere afFacMom vfFacMom cfFacMom mtFacMom").split()
PARAM_I = ("momAdvection momViscosity useBiharmonicVisc implicitViscosity no_slip_sides no_slip_bottom "
           "bottomVisc_pCell selectBotDragQuadr selectImplicitDrag useCDscheme selectCoriScheme "
           "selectMetricTerms usingSphericalPolarGrid rigidLid select_rStar selectKEscheme "
           "implicitDiffusion useCoriolis useAbsVorticity selectVortScheme useJamartMomAdv upwindShear "
           "highOrderVorticity upwindVorticity momImplVertAdv").split()


class ParamS(C.Structure):
    _fields_ = [(n, C.c_double) for n in PARAM_D] + [(n, C.c_int) for n in PARAM_I]


class Cg2dOp(C.Structure):
    _fields_ = [(n, P) for n in "aW2d aS2d aC2d pW pS pC".split()] + \
               [("cg2dNorm", C.c_double), ("cg2dTolerance_sq", C.c_double), ("cg2dNormaliseRHS", C.c_int)]


class Cg3dOp(C.Structure):
    _fields_ = [(n, P) for n in "aW3d aS3d aV3d aC3d zMC zML zMU".split()] + \
               [("cg3dNorm", C.c_double), ("cg3dTolerance_sq", C.c_double), ("cg3dNormaliseRHS", C.c_int)]


DEFAULT_PARAMS = dict(
    deltaTMom=1200.0, deltaTFreeSurf=1200.0, freeSurfFac=1.0, implicSurfPress=1.0, implicDiv2DFlow=1.0,
    rkSign=-1.0, cg2dpcOffDFac=0.51, cg2dTargetResidual=1e-7, cg2dTargetResWunit=-1.0, globalArea=0.0,
    viscAhD=0.0, viscAhZ=0.0, viscA4D=0.0, viscA4Z=0.0, sideDragFactor=2.0, bottomDragLinear=0.0,
    bottomDragQuadratic=0.0, recip_rSphere=1.0 / 6370e3, afFacMom=1.0, vfFacMom=1.0, cfFacMom=1.0,
    mtFacMom=1.0, momAdvection=1, momViscosity=1, useBiharmonicVisc=0, implicitViscosity=0,
    no_slip_sides=1, no_slip_bottom=1, bottomVisc_pCell=0, selectBotDragQuadr=-1, selectImplicitDrag=0,
    useCDscheme=0, selectCoriScheme=0, selectMetricTerms=0, usingSphericalPolarGrid=0, rigidLid=0,
    select_rStar=0, selectKEscheme=0, implicitDiffusion=0, useCoriolis=1, useAbsVorticity=0, selectVortScheme=1,
    useJamartMomAdv=0, upwindShear=0, highOrderVorticity=0, upwindVorticity=0, momImplVertAdv=0)


def ptr(a):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"], "oracle needs contiguous float64"
    return a.ctypes.data_as(P)


class Oracle:
    """Holds the loaded library plus a grid/params pair in C form."""

    def __init__(self, grid, params=None):
        self.lib = C.CDLL(build())
        self.grid = grid
        d = grid.d
        self.d = Dims(d.sNx, d.sNy, d.OLx, d.OLy, d.nSx, d.nSy, d.Nr)
        self.g = GridS()
        self.g.d = self.d
        self._keep = []
        for n in GRID_FIELDS:
            a = grid.a.get(n)
            if a is None:
                a = np.zeros(d.shape3 if n.startswith(("hFac", "recip_hFac", "mask")) else d.shape2)
            a = np.ascontiguousarray(a, dtype=np.float64)
            grid.a[n] = a
            setattr(self.g, n, ptr(a))
        self.set_params(**(params or {}))
        L = self.lib
        L.og_global_sum_tile.restype = C.c_double

    def set_params(self, **kw):
        self.params = dict(DEFAULT_PARAMS)
        self.params.update(kw)
        self.p = ParamS(**self.params)

    # ---- eesupp ----
    def exch_xyz(self, a, nz=1):
        self.lib.og_exch_xyz(C.byref(self.d), ptr(a), C.c_int(nz))
        return a

    # ---- cg2d ----
    def ini_cg2d(self):
        d = self.grid.d
        arrs = {n: np.zeros(d.shape2) for n in "aW2d aS2d aC2d pW pS pC".split()}
        op = Cg2dOp(**{n: ptr(a) for n, a in arrs.items()})
        self.lib.og_ini_cg2d(C.byref(self.g), C.byref(self.p), C.byref(op))
        arrs.update(cg2dNorm=op.cg2dNorm, cg2dTolerance_sq=op.cg2dTolerance_sq,
                    cg2dNormaliseRHS=bool(op.cg2dNormaliseRHS))
        return arrs

    def update_cg2d(self, op, updatePreCond=True):
        """UPDATE_CG2D (update_cg2d.F:57-192) on the grid's CURRENT hFacW / hFacS: rewrites aW2d, aS2d, aC2d (and the
        preconditioner when updatePreCond) of `op` in place; cg2dNorm and the tolerance stay."""
        cop = self._op(op)
        self.lib.og_update_cg2d(C.byref(self.g), C.byref(self.p), C.byref(cop), C.c_int(int(updatePreCond)))
        return op

    def _op(self, op):
        return Cg2dOp(**{n: ptr(op[n]) for n in "aW2d aS2d aC2d pW pS pC".split()},
                      cg2dNorm=op["cg2dNorm"], cg2dTolerance_sq=op["cg2dTolerance_sq"],
                      cg2dNormaliseRHS=int(op["cg2dNormaliseRHS"]))

    def cg2d(self, op, b, x, numIters, nIterMin=-1, sr=False, history=False):
        """Returns dict(firstResidual, minResidualSq, lastResidual, numIters, nIterMin, sumRHS, rhsMax[, hist]);
        b and x are updated in place exactly as CG2D does."""
        cop = self._op(op)
        f, m, l, s, r = (C.c_double() for _ in range(5))
        ni, nm = C.c_int(numIters), C.c_int(nIterMin)
        hist = np.zeros(max(numIters, 1)) if history else None
        fn = self.lib.og_cg2d_sr if sr else self.lib.og_cg2d
        fn(C.byref(self.d), C.byref(cop), ptr(b), ptr(x), C.byref(f), C.byref(m), C.byref(l),
           C.byref(ni), C.byref(nm), C.byref(s), C.byref(r), ptr(hist) if history else None)
        out = dict(firstResidual=f.value, minResidualSq=m.value, lastResidual=l.value,
                   numIters=ni.value, nIterMin=nm.value, sumRHS=s.value, rhsMax=r.value)
        if history:
            out["hist"] = hist[:ni.value]
        return out

    # ---- cg3d ----
    def ini_cg3d(self, vertFac=1.0, cg3dTargetResidual=1e-7, cg3dTargetResWunit=-1.0):
        """INI_CG3D: operators aW3d..aC3d and the preconditioner zMC, zML, zMU (tile3d) + cg3dNorm, tolerance."""
        d = self.grid.d
        arrs = {n: np.zeros(d.shape3) for n in "aW3d aS3d aV3d aC3d zMC zML zMU".split()}
        op = Cg3dOp(**{n: ptr(a) for n, a in arrs.items()})
        self.lib.og_ini_cg3d(C.byref(self.g), C.byref(self.p), C.c_double(vertFac), C.c_double(cg3dTargetResidual),
                             C.c_double(cg3dTargetResWunit), C.byref(op))
        arrs.update(cg3dNorm=op.cg3dNorm, cg3dTolerance_sq=op.cg3dTolerance_sq, cg3dNormaliseRHS=bool(op.cg3dNormaliseRHS))
        return arrs

    def cg3d(self, op, b, x, numIters):
        """CG3D; b and x are updated in place as the reference does.  Returns dict(firstResidual, lastResidual,
        numIters, sumRHS, rhsMax)."""
        cop = Cg3dOp(**{n: ptr(op[n]) for n in "aW3d aS3d aV3d aC3d zMC zML zMU".split()}, cg3dNorm=op["cg3dNorm"],
                     cg3dTolerance_sq=op["cg3dTolerance_sq"], cg3dNormaliseRHS=int(op["cg3dNormaliseRHS"]))
        f, l, s, r = (C.c_double() for _ in range(4))
        ni = C.c_int(numIters)
        self.lib.og_cg3d(C.byref(self.d), C.byref(cop), ptr(self.grid.a["maskC"]), ptr(b), ptr(x), C.byref(f), C.byref(l),
                         C.byref(ni), C.byref(s), C.byref(r))
        return dict(firstResidual=f.value, lastResidual=l.value, numIters=ni.value, sumRHS=s.value, rhsMax=r.value)

    # ---- momentum ----
    def mom_fluxform(self, bi, bj, k, iMin, iMax, jMin, jMax, kappaRU, kappaRV, fVerUkm, fVerVkm,
                     fVerUkp, fVerVkp, guDiss, gvDiss, uVel, vVel, wVel, gU, gV):
        self.lib.og_mom_fluxform(C.byref(self.g), C.byref(self.p), bi, bj, k, iMin, iMax, jMin, jMax,
                                 ptr(kappaRU), ptr(kappaRV), ptr(fVerUkm), ptr(fVerVkm), ptr(fVerUkp),
                                 ptr(fVerVkp), ptr(guDiss), ptr(gvDiss), ptr(uVel), ptr(vVel), ptr(wVel),
                                 ptr(gU), ptr(gV))

    def mom_vecinv(self, bi, bj, k, iMin, iMax, jMin, jMax, kappaRU, kappaRV, fVerUkm, fVerVkm,
                   fVerUkp, fVerVkp, guDiss, gvDiss, uVel, vVel, wVel, gU, gV, csCorners=0, myFace=0):
        rc = self.lib.og_mom_vecinv(C.byref(self.g), C.byref(self.p), bi, bj, k, iMin, iMax, jMin, jMax,
                                    ptr(kappaRU), ptr(kappaRV), ptr(fVerUkm), ptr(fVerVkm), ptr(fVerUkp),
                                    ptr(fVerVkp), ptr(guDiss), ptr(gvDiss), ptr(uVel), ptr(vVel), ptr(wVel),
                                    ptr(gU), ptr(gV), int(csCorners), int(myFace))
        if rc:
            raise ValueError(f"og_mom_vecinv: option not restated (rc={rc})")

    # ---- tracers ----
    def calc_adv_flow(self, bi, bj, k, uVel, vVel, wVel, xA, yA, maskUp, uFld, vFld, wFld,
                      uTrans, vTrans, rTrans, rTransKp1):
        self.lib.og_calc_adv_flow(C.byref(self.g), bi, bj, k, ptr(uVel), ptr(vVel), ptr(wVel), ptr(xA), ptr(yA),
                                  ptr(maskUp), ptr(uFld), ptr(vFld), ptr(wFld), ptr(uTrans), ptr(vTrans),
                                  ptr(rTrans), ptr(rTransKp1))

    def gad_calc_rhs(self, bi, bj, iMin, iMax, jMin, jMax, k, kM1, kUp, kDown, xA, yA, maskUp, uFld, vFld,
                     wFld, uTrans, vTrans, rTrans, rTransKp1, diffKh, diffK4, KappaR, diffKr4, TracerN,
                     TracAB, deltaTLev, advScheme, vertAdvScheme, calcAdvection, implicitAdvection,
                     applyAB_onTracer, trUseDiffKr4, fZon, fMer, fVerT, gTracer):
        self.lib.og_gad_calc_rhs(
            C.byref(self.g), C.byref(self.p), bi, bj, iMin, iMax, jMin, jMax, k, kM1, kUp, kDown,
            ptr(xA), ptr(yA), ptr(maskUp), ptr(uFld), ptr(vFld), ptr(wFld), ptr(uTrans), ptr(vTrans),
            ptr(rTrans), ptr(rTransKp1), C.c_double(diffKh), C.c_double(diffK4), ptr(KappaR), ptr(diffKr4),
            ptr(TracerN), ptr(TracAB), ptr(deltaTLev), int(advScheme), int(vertAdvScheme),
            int(calcAdvection), int(implicitAdvection), int(applyAB_onTracer), int(trUseDiffKr4),
            ptr(fZon), ptr(fMer), ptr(fVerT), ptr(gTracer))

    def gad_advection(self, bi, bj, advScheme, vertAdvScheme, implicitAdvection, compressible, deltaTLev,
                      uFld, vFld, wFld, tracer, gTracer, nCFace=0, edges=0):
        """GAD_ADVECTION (multi-dimensional advection) for one tile; gTracer is (Nr, PY, PX).
        Cubed sphere: nCFace = facet number, edges = 1 N | 2 S | 4 E | 8 W facet edges the tile touches."""
        return self.lib.og_gad_advection(C.byref(self.g), C.byref(self.p), bi, bj, int(advScheme), int(vertAdvScheme),
                                         int(implicitAdvection), int(compressible), ptr(deltaTLev), ptr(uFld), ptr(vFld),
                                         ptr(wFld), ptr(tracer), ptr(gTracer), int(nCFace), int(edges))

    # ---- glue ----
    def timestep(self, bi, bj, k, iMin, iMax, jMin, jMax, dPhiHydX, dPhiHydY, guDiss, gvDiss, sfU, sfV,
                 momForcing, momDissip_In_AB, abFac, uVel, vVel, gU, gV, guNm1, gvNm1, phiSurfX=None, phiSurfY=None):
        self.lib.og_timestep(C.byref(self.g), C.byref(self.p), bi, bj, k, iMin, iMax, jMin, jMax,
                             ptr(dPhiHydX), ptr(dPhiHydY), ptr(guDiss), ptr(gvDiss), ptr(sfU), ptr(sfV),
                             int(momForcing), int(momDissip_In_AB), C.c_double(abFac), ptr(uVel), ptr(vVel),
                             ptr(gU), ptr(gV), ptr(guNm1), ptr(gvNm1),
                             ptr(phiSurfX) if phiSurfX is not None else None,
                             ptr(phiSurfY) if phiSurfY is not None else None)

    def calc_grad_phi_surf(self, bi, bj, iMin, iMax, jMin, jMax, etaFld, phiSurfX, phiSurfY):
        self.lib.og_calc_grad_phi_surf(C.byref(self.g), bi, bj, iMin, iMax, jMin, jMax,
                                       ptr(self.grid.a["Bo_surf"]), ptr(etaFld), ptr(phiSurfX), ptr(phiSurfY))

    def solve_rhs(self, bi, bj, etaN, gU, gV, b, x, etaH=None):
        """etaH given = exactConserv (solve_for_pressure.F:213-222)."""
        self.lib.og_solve_rhs(C.byref(self.g), C.byref(self.p), bi, bj, ptr(self.grid.a["Bo_surf"]), ptr(etaN),
                              ptr(etaN if etaH is None else etaH), ptr(gU), ptr(gV), ptr(b), ptr(x))

    def correction_step(self, bi, bj, etaN, gU, gV, uVel, vVel):
        self.lib.og_correction_step(C.byref(self.g), C.byref(self.p), bi, bj, ptr(self.grid.a["Bo_surf"]),
                                    ptr(etaN), ptr(gU), ptr(gV), ptr(uVel), ptr(vVel))

    def integrate_for_w(self, bi, bj, uVel, vVel, wVel):
        self.lib.og_integrate_for_w(C.byref(self.g), C.byref(self.p), bi, bj, ptr(uVel), ptr(vVel), ptr(wVel))

    # ---- non-hydrostatic step around CG3D (nh_oracle.c) ----
    def calc_gw(self, bi, bj, R_low, Ro_surf, rLowW, rSurfW, rLowS, rSurfS, rC, kappaRU, kappaRV, viscAhW, viscA4W,
                momDissip_In_AB, abFac, uVel, vVel, wVel, gW, gwNm1):
        """CALC_GW + ADAMS_BASHFORTH2 on gW (calc_gw.F:156-640); returns non-zero for options that are not restated."""
        return self.lib.og_calc_gw(C.byref(self.g), C.byref(self.p), bi, bj, ptr(R_low), ptr(Ro_surf), ptr(rLowW), ptr(rSurfW),
                                   ptr(rLowS), ptr(rSurfS), ptr(rC), ptr(kappaRU), ptr(kappaRV), C.c_double(viscAhW),
                                   C.c_double(viscA4W), int(momDissip_In_AB), C.c_double(abFac), ptr(uVel), ptr(vVel),
                                   ptr(wVel), ptr(gW), ptr(gwNm1))

    def timestep_wvel(self, bi, bj, gW, wVel):
        self.lib.og_timestep_wvel(C.byref(self.g), C.byref(self.p), bi, bj, ptr(gW), ptr(wVel))

    def solve_rhs_nh(self, bi, bj, etaN, phi_nh, gU, gV, b, x, b3):
        self.lib.og_solve_rhs_nh(C.byref(self.g), C.byref(self.p), bi, bj, ptr(self.grid.a["Bo_surf"]), ptr(etaN),
                                 ptr(phi_nh), ptr(gU), ptr(gV), ptr(b), ptr(x), ptr(b3))

    def pre_cg3d(self, bi, bj, cg2d_x, etaN, wVel, b3):
        self.lib.og_pre_cg3d(C.byref(self.g), C.byref(self.p), bi, bj, ptr(cg2d_x), ptr(etaN), ptr(wVel), ptr(b3))

    def correction_step_nh(self, bi, bj, etaN, phi_nh, gU, gV, uVel, vVel):
        self.lib.og_correction_step_nh(C.byref(self.g), C.byref(self.p), bi, bj, ptr(self.grid.a["Bo_surf"]), ptr(etaN),
                                       ptr(phi_nh), ptr(gU), ptr(gV), ptr(uVel), ptr(vVel))

    # ---- physics glue of config 2 (phys_oracle.c) ----
    def density_ivdc(self, eos, bi, bj, theta, salt, tRef, sRef, rhoInSitu, IVDConvCount):
        self.lib.og_density_ivdc(C.byref(self.g), C.byref(self.p), C.byref(eos), bi, bj, ptr(theta), ptr(salt),
                                 ptr(tRef), ptr(sRef), ptr(rhoInSitu), ptr(IVDConvCount))

    def forcing_surf_relax_T(self, bi, bj, theta, SST, lam, recip_Cp, mass2rUnit, sfT):
        self.lib.og_forcing_surf_relax_T(C.byref(self.g), bi, bj, ptr(theta), ptr(SST), ptr(lam),
                                         C.c_double(recip_Cp), C.c_double(mass2rUnit), ptr(sfT))

    def apply_forcing_T(self, bi, bj, k, sfT, gtForc):
        self.lib.og_apply_forcing_T(C.byref(self.g), bi, bj, k, ptr(sfT), ptr(gtForc))

    def calc_3d_diffusivity(self, bi, bj, IVDConvCount, ivdc_kappa, KbryanLewis79, diffKrNrT, kappaRk):
        self.lib.og_calc_3d_diffusivity(C.byref(self.g), bi, bj, ptr(IVDConvCount), C.c_double(ivdc_kappa),
                                        ptr(KbryanLewis79), ptr(diffKrNrT), ptr(kappaRk))

    def mom_implicit_r(self, bi, bj, isV, kappaR, gFld):
        """MOM_U_IMPLICIT_R (isV = 0, on gU) / MOM_V_IMPLICIT_R (isV = 1, on gV), implicitViscosity only."""
        return self.lib.og_mom_implicit_r(C.byref(self.g), C.byref(self.p), bi, bj, int(isV), ptr(kappaR), ptr(gFld))

    def gad_implicit_r(self, bi, bj, iMin, iMax, jMin, jMax, deltaTLev, kappaRX, recip_hFac, gTracer):
        return self.lib.og_gad_implicit_r(C.byref(self.g), bi, bj, iMin, iMax, jMin, jMax, ptr(deltaTLev),
                                          ptr(kappaRX), ptr(recip_hFac), ptr(gTracer))

    def calc_phi_hyd(self, bi, bj, iMin, iMax, jMin, jMax, k, rhoInSitu, rF, rC, gravity, recip_rhoConst,
                     phi0surf, phiHydF, phiHydC, dPhiHydX, dPhiHydY):
        self.lib.og_calc_phi_hyd(C.byref(self.g), bi, bj, iMin, iMax, jMin, jMax, k, ptr(rhoInSitu), ptr(rF),
                                 ptr(rC), C.c_double(gravity), C.c_double(recip_rhoConst), ptr(phi0surf),
                                 ptr(phiHydF), ptr(phiHydC), ptr(dPhiHydX), ptr(dPhiHydY))

    def integr_continuity_ec(self, bi, bj, uVel, vVel, etaH, dEtaHdt, etaN, updateEtaN=True):
        self.lib.og_integr_continuity_ec(C.byref(self.g), C.byref(self.p), bi, bj, ptr(uVel), ptr(vVel),
                                         ptr(etaH), ptr(dEtaHdt), ptr(etaN), int(updateEtaN))


class Eos(C.Structure):
    _fields_ = [(n, C.c_double) for n in "rhoNil rhoConst tAlpha sBeta".split()]
