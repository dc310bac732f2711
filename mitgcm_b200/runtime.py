"""Host-side mirror of the reference interfaces on top of the C ABI.

Function names and argument meaning follow the Fortran routines they replace
(CG2D, CG2D_SR, GAD_CALC_RHS, MOM_FLUXFORM); arrays are numpy arrays (host,
reference layout) or torch CUDA tensors (resident on the GPU, no copies)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .grid import Dims, Grid

E = _lib.ENUMS


class B200Error(RuntimeError):
    pass


def _addr(a):
    """void* of a numpy array (host) or a torch tensor (host or device)."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"], "need contiguous float64"
        return C.c_void_p(a.ctypes.data)
    # torch tensor
    import torch
    assert isinstance(a, torch.Tensor) and a.dtype == torch.float64 and a.is_contiguous()
    if a.is_cuda:
        # the library works on its own stream: whatever torch still has in flight for this tensor must be done
        # before the library reads it (bench.py once uploaded an eta that torch had not finished computing)
        torch.cuda.current_stream(a.device).synchronize()
    return C.c_void_p(a.data_ptr())


def _check(ierr=None):
    L = _lib.lib()
    code = L.mitgcm_b200_last_error_() if ierr is None else (ierr.value or L.mitgcm_b200_last_error_() if ierr.value else 0)
    if code:
        raise B200Error(f"libmitgcm_b200 error {code}: {L.mitgcm_b200_last_error_string().decode()}")


def init(d: Dims, device: int = -1):
    L = _lib.lib()
    dims = (C.c_int * 11)(d.sNx, d.sNy, d.OLx, d.OLy, d.nSx, d.nSy, d.Nr, d.nPx, d.nPy, d.myPx, d.myPy)
    ierr = C.c_int(0)
    L.mitgcm_b200_init_(dims, C.byref(C.c_int(device)), C.byref(ierr))
    _check(ierr)


def finalize():
    _lib.lib().mitgcm_b200_finalize_()


def sync():
    _lib.lib().mitgcm_b200_sync_()


def set_params(**kw):
    """set_params(deltaTMom=1200., momAdvection=True, ...): names as in PARAMS.h."""
    L = _lib.lib()
    ierr = C.c_int(0)
    for k, v in kw.items():
        kd, ki = "MP_" + k.upper(), "MI_" + k.upper()
        if kd in E:
            L.mitgcm_b200_set_param_d_(C.byref(C.c_int(E[kd])), C.byref(C.c_double(float(v))), C.byref(ierr))
        elif ki in E:
            L.mitgcm_b200_set_param_i_(C.byref(C.c_int(E[ki])), C.byref(C.c_int(int(v))), C.byref(ierr))
        else:
            raise KeyError(f"unknown parameter {k}")
        _check(ierr)


def field_id(name: str) -> int:
    return E["MG_" + name.upper()]


def set_field(name: str, a):
    ierr = C.c_int(0)
    _lib.lib().mitgcm_b200_set_field_(C.byref(C.c_int(field_id(name))), _addr(a), C.byref(ierr))
    _check(ierr)


def get_field(name: str, out):
    ierr = C.c_int(0)
    _lib.lib().mitgcm_b200_get_field_(C.byref(C.c_int(field_id(name))), _addr(out), C.byref(ierr))
    _check(ierr)
    return out


GRID_FIELD_NAMES = ("dxC dyC dxG dyG dxF dyF dxV dyU rA rAw rAs rAz recip_dxC recip_dyC recip_dxG recip_dyG "
                    "recip_dxF recip_dyF recip_dxV recip_dyU recip_rA recip_rAw recip_rAs recip_rAz fCori fCoriG "
                    "tanPhiAtU tanPhiAtV recip_Bo Bo_surf hFacC hFacW hFacS recip_hFacC recip_hFacW recip_hFacS "
                    "maskC maskW maskS cosFacU cosFacV").split()


def set_grid(g: Grid):
    """Uploads the GRID.h arrays (what INI_GRID / INI_MASKS_ETC left in COMMON)."""
    for n in GRID_FIELD_NAMES:
        if n in g.a:
            set_field(n, np.ascontiguousarray(g.a[n], dtype=np.float64))
    Nr = g.d.Nr
    for n in ("drF", "drC", "recip_drF", "recip_drC"):
        v = np.zeros(Nr + 1)
        v[:len(g.a[n])] = g.a[n]
        set_field(n, v)


def set_cg2d_operator(op: dict):
    """Uploads COMMON /CG2D_I_RS/ + cg2dNorm, cg2dTolerance_sq, cg2dNormaliseRHS
    (model/inc/CG2D.h) as produced by INI_CG2D / UPDATE_CG2D."""
    for n in "aW2d aS2d aC2d pW pS pC".split():
        set_field(n, op[n])
    set_params(cg2dNorm=op["cg2dNorm"], cg2dTolerance_sq=op["cg2dTolerance_sq"],
               cg2dNormaliseRHS=bool(op["cg2dNormaliseRHS"]))


def pin_host(a: np.ndarray):
    """mitgcm_b200_pin_host_: page-lock a host array that will be passed to the per-call entry points again and again
    (what the shims do for the COMMON-block arrays); keep the array alive until finalize()."""
    ierr = C.c_int(0)
    _lib.lib().mitgcm_b200_pin_host_(_addr(a), C.byref(C.c_longlong(a.size)), C.byref(ierr))
    _check(ierr)


def update_cg2d(myIter: int, myTime: float = 0.0, myThid: int = 1):
    """CALL UPDATE_CG2D( myTime, myIter, myThid ) -- model/src/update_cg2d.F:7: operator (and preconditioner) from
    the current hFacW / hFacS mirrors."""
    _lib.lib().update_cg2d_b200_(_d(myTime), _i(myIter), _i(myThid))
    _check()


def cg2d(cg2d_b, cg2d_x, numIters: int, nIterMin: int = -1, sr: bool = False, residuals: bool = False):
    """CALL CG2D( cg2d_b, cg2d_x, firstResidual, minResidualSq, lastResidual, numIters,
    nIterMin, myThid ) -- model/src/cg2d.F:13-17 (CG2D_SR when sr).  b and x are updated in
    place like the Fortran dummies; the scalar outputs come back as a dict, together with the
    `cg2d: Sum(rhs),rhsMax` values the reference prints inside the solver."""
    L = _lib.lib()
    f, m, l, s, r = (C.c_double() for _ in range(5))
    ni, nm = C.c_int(numIters), C.c_int(nIterMin)
    fn = L.cg2d_sr_b200_ if sr else L.cg2d_b200_
    fn(_addr(cg2d_b), _addr(cg2d_x), C.byref(f), C.byref(m), C.byref(l), C.byref(ni), C.byref(nm),
       C.byref(C.c_int(1)))
    _check()
    L.mitgcm_b200_cg2d_stats_(C.byref(s), C.byref(r))
    out = dict(firstResidual=f.value, minResidualSq=m.value, lastResidual=l.value, numIters=ni.value,
               nIterMin=nm.value, sumRHS=s.value, rhsMax=r.value)
    if residuals:
        h = np.zeros(max(ni.value, 1))
        L.mitgcm_b200_cg2d_residuals_(_addr(h), C.byref(C.c_int(ni.value)))
        out["hist"] = h[:ni.value]
    return out


def _i(v):
    return C.byref(C.c_int(int(v)))


def _d(v):
    return C.byref(C.c_double(float(v)))


def gad_calc_rhs(bi, bj, iMin, iMax, jMin, jMax, k, kM1, kUp, kDown, xA, yA, maskUp, uFld, vFld, wFld,
                 uTrans, vTrans, rTrans, rTransKp1, diffKh, diffK4, KappaR, diffKr4, TracerN, TracAB,
                 deltaTLev, trIdentity, advectionSchArg, vertAdvecSchArg, calcAdvection, implicitAdvection,
                 applyAB_onTracer, trUseDiffKr4, trUseGMRedi, trUseKPP, trUseSmolHack, fZon, fMer, fVerT,
                 gTracer, myTime=0.0, myIter=0, myThid=1):
    """CALL GAD_CALC_RHS(...) -- pkg/generic_advdiff/gad_calc_rhs.F:10-21, same argument order.
    diffKr4 and deltaTLev are host numpy vectors of length Nr."""
    L = _lib.lib()
    L.gad_calc_rhs_b200_(
        _i(bi), _i(bj), _i(iMin), _i(iMax), _i(jMin), _i(jMax), _i(k), _i(kM1), _i(kUp), _i(kDown),
        _addr(xA), _addr(yA), _addr(maskUp), _addr(uFld), _addr(vFld), _addr(wFld), _addr(uTrans),
        _addr(vTrans), _addr(rTrans), _addr(rTransKp1), _d(diffKh), _d(diffK4), _addr(KappaR),
        _addr(np.ascontiguousarray(diffKr4, dtype=np.float64)), _addr(TracerN), _addr(TracAB),
        _addr(np.ascontiguousarray(deltaTLev, dtype=np.float64)), _i(trIdentity), _i(advectionSchArg),
        _i(vertAdvecSchArg), _i(calcAdvection), _i(implicitAdvection), _i(applyAB_onTracer), _i(trUseDiffKr4),
        _i(trUseGMRedi), _i(trUseKPP), _i(trUseSmolHack), _addr(fZon), _addr(fMer), _addr(fVerT),
        _addr(gTracer), _d(myTime), _i(myIter), _i(myThid))
    _check()


def gad_advection(implicitAdvection, advectionSchArg, vertAdvecSchArg, trIdentity, deltaTLev, uFld, vFld, wFld,
                  tracer, gTracer, bi, bj, myTime=0.0, myIter=0, myThid=1):
    """CALL GAD_ADVECTION(...) -- pkg/generic_advdiff/gad_advection.F:11-17, same argument order.
    uFld, vFld, wFld, gTracer: (Nr, PY, PX) arrays of tile (bi, bj); tracer: the full tiled array."""
    L = _lib.lib()
    L.gad_advection_b200_(_i(implicitAdvection), _i(advectionSchArg), _i(vertAdvecSchArg), _i(trIdentity),
                          _addr(np.ascontiguousarray(deltaTLev, dtype=np.float64)), _addr(uFld), _addr(vFld),
                          _addr(wFld), _addr(tracer), _addr(gTracer), _i(bi), _i(bj), _d(myTime), _i(myIter),
                          _i(myThid))
    _check()


def mom_fluxform(bi, bj, k, iMin, iMax, jMin, jMax, kappaRU, kappaRV, fVerUkm, fVerVkm, fVerUkp, fVerVkp,
                 guDiss, gvDiss, uVel, vVel, wVel, gU, gV, myTime=0.0, myIter=0, myThid=1):
    """CALL MOM_FLUXFORM(...) -- pkg/mom_fluxform/mom_fluxform.F:42-48, followed by the COMMON
    /DYNVARS_R/ arrays uVel, vVel, wVel (in) and gU, gV (out) the routine works on."""
    L = _lib.lib()
    L.mom_fluxform_b200_(_i(bi), _i(bj), _i(k), _i(iMin), _i(iMax), _i(jMin), _i(jMax), _addr(kappaRU),
                         _addr(kappaRV), _addr(fVerUkm), _addr(fVerVkm), _addr(fVerUkp), _addr(fVerVkp),
                         _addr(guDiss), _addr(gvDiss), _d(myTime), _i(myIter), _i(myThid), _addr(uVel),
                         _addr(vVel), _addr(wVel), _addr(gU), _addr(gV))
    _check()


def mom_vecinv(bi, bj, k, iMin, iMax, jMin, jMax, kappaRU, kappaRV, fVerUkm, fVerVkm, fVerUkp, fVerVkp,
               guDiss, gvDiss, uVel, vVel, wVel, gU, gV, csCorners=0, myFace=0, myTime=0.0, myIter=0, myThid=1):
    """CALL MOM_VECINV(...) -- pkg/mom_vecinv/mom_vecinv.F:10-16, followed by the COMMON /DYNVARS_R/
    arrays and the tile's facet-corner mask / facet number (cubed sphere; 0, 0 otherwise)."""
    L = _lib.lib()
    L.mom_vecinv_b200_(_i(bi), _i(bj), _i(k), _i(iMin), _i(iMax), _i(jMin), _i(jMax), _addr(kappaRU),
                       _addr(kappaRV), _addr(fVerUkm), _addr(fVerVkm), _addr(fVerUkp), _addr(fVerVkp),
                       _addr(guDiss), _addr(gvDiss), _d(myTime), _i(myIter), _i(myThid), _addr(uVel),
                       _addr(vVel), _addr(wVel), _addr(gU), _addr(gV), _i(csCorners), _i(myFace))
    _check()


def set_cg3d_operator(op: dict):
    """Uploads INI_CG3D's output: aW3d, aS3d, aV3d, aC3d, zMC, zML, zMU (full-halo tile3d) and cg3dNorm,
    cg3dTolerance_sq, cg3dNormaliseRHS."""
    for n in "aW3d aS3d aV3d aC3d zMC zML zMU".split():
        set_field(n, np.ascontiguousarray(op[n]))
    set_params(cg3dNorm=op["cg3dNorm"], cg3dTolerance_sq=op["cg3dTolerance_sq"], cg3dNormaliseRHS=int(op["cg3dNormaliseRHS"]))


def cg3d(b, x, numIters, myIter=0, myThid=1):
    """CALL CG3D(cg3d_b, cg3d_x, firstResidual, lastResidual, numIters, myIter, myThid) -- model/src/cg3d.F:13-17.
    Returns dict(firstResidual, lastResidual, numIters, sumRHS, rhsMax); b and x are updated in place."""
    L = _lib.lib()
    f, l = C.c_double(), C.c_double()
    ni = C.c_int(numIters)
    L.cg3d_b200_(_addr(b), _addr(x), C.byref(f), C.byref(l), C.byref(ni), _i(myIter), _i(myThid))
    _check()
    s, r = C.c_double(), C.c_double()
    L.mitgcm_b200_cg3d_rhs_stats_(C.byref(s), C.byref(r))
    return dict(firstResidual=f.value, lastResidual=l.value, numIters=ni.value, sumRHS=s.value, rhsMax=r.value)


def mom_implicit_r(kappaR, bi, bj, gFld, isV=False, myTime=0.0, myIter=0, myThid=1):
    """CALL MOM_U_IMPLICIT_R / MOM_V_IMPLICIT_R(kappaR?, bi, bj, ...) -- pkg/mom_common/mom_{u,v}_implicit_r.F:6-8,
    followed by the COMMON array solved in place (gU / gV)."""
    L = _lib.lib()
    fn = L.mom_v_implicit_r_b200_ if isV else L.mom_u_implicit_r_b200_
    fn(_addr(kappaR), _i(bi), _i(bj), _d(myTime), _i(myIter), _i(myThid), _addr(gFld))
    _check()


def fill_field(name: str, value: float):
    ierr = C.c_int(0)
    _lib.lib().mitgcm_b200_fill_field_(C.byref(C.c_int(field_id(name))), _d(value), C.byref(ierr))
    _check(ierr)


def exch(name: str):
    """_EXCH_XY_RL / _EXCH_XYZ_RL on a device mirror."""
    ierr = C.c_int(0)
    _lib.lib().mitgcm_b200_exch_(C.byref(C.c_int(field_id(name))), C.byref(ierr))
    _check(ierr)


def exch_uv(nameU: str, nameV: str, withSigns: bool = True):
    """_EXCH_UV_XY_RL / _EXCH_UV_XYZ_RL on a pair of device mirrors."""
    ierr = C.c_int(0)
    _lib.lib().mitgcm_b200_exch_uv_(C.byref(C.c_int(field_id(nameU))), C.byref(C.c_int(field_id(nameV))),
                                    C.byref(C.c_int(int(withSigns))), C.byref(ierr))
    _check(ierr)


def forward_step(myIter: int):
    """One FORWARD_STEP on the resident state; returns the three numbers SOLVE_FOR_PRESSURE
    prints (cg2d_init_res, cg2d_iters, cg2d_last_res)."""
    f, l, n, ierr = C.c_double(), C.c_double(), C.c_int(), C.c_int(0)
    _lib.lib().mitgcm_b200_forward_step_(_i(myIter), C.byref(f), C.byref(n), C.byref(l), C.byref(ierr))
    _check(ierr)
    return dict(firstResidual=f.value, numIters=n.value, lastResidual=l.value)
