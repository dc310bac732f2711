"""Parity of the resident CUDA model step (mitgcm_b200_forward_step_) against the oracle step
(oracle/channel.py, the Nr-level extension of the golden-pinned barotropic-gyre driver).

Tolerance: the fused kernels are point-wise identical to the per-level routines; differences come
only from CG2D's dot-product summation order (~1e-12 on eta per step).  Fields are compared
relative to their own max after several steps with 1e-9; iteration counts +-1."""
import numpy as np
import pytest

from mitgcm_b200.model import make_channel, Model
from oracle.channel import ChannelOracle

pytestmark = pytest.mark.gpu


def rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.mark.parametrize("cfg", [
    dict(sNx=24, sNy=16, Nr=4, nSx=2, nSy=2, land_frac=0.2),
    dict(sNx=62, sNy=62, Nr=1, land_frac=0.0, tempStepping=0),
    dict(sNx=40, sNy=36, Nr=6, land_frac=0.1, OL=3, tempAdvScheme=33, viscA4D=1e10, viscA4Z=1e10, useBiharmonicVisc=1),
    dict(sNx=32, sNy=32, Nr=5, land_frac=0.0, momDissip_In_AB=0, selectCoriScheme=1, rigidLid=0),
    # eosType = 'LINEAR' coupled to the momentum equations (CALC_PHI_HYD): pipelined dyn kernel, on-the-fly EOS
    dict(sNx=64, sNy=40, Nr=7, land_frac=0.0, buoyancyLinear=1),
    # same with land / partial cells and 2x2 tiles (generic dyn kernel is not used: fast path handles masks)
    dict(sNx=24, sNy=16, Nr=4, nSx=2, nSy=2, land_frac=0.2, buoyancyLinear=1),
    # IVDC on: density array + convective diffusivity feed the (explicit) vertical diffusion
    dict(sNx=32, sNy=24, Nr=6, land_frac=0.1, buoyancyLinear=1, ivdc_kappa=1.0, selectCoriScheme=1),
    # MOM_VECINV in the dynamics: pipelined kernel (defaults), with land / partial cells / tiles / buoyancy,
    # absolute vorticity + Jamart + upwind shear + free slip, and the generic kernel (vorticity scheme 3, KE scheme 1)
    dict(sNx=64, sNy=40, Nr=7, land_frac=0.0, vectorInvariantMomentum=1),
    dict(sNx=24, sNy=16, Nr=4, nSx=2, nSy=2, land_frac=0.2, vectorInvariantMomentum=1, buoyancyLinear=1),
    dict(sNx=40, sNy=36, Nr=6, land_frac=0.1, vectorInvariantMomentum=1, useAbsVorticity=1, useJamartMomAdv=1, upwindShear=1,
         selectVortScheme=2, selectKEscheme=2, selectCoriScheme=3, no_slip_sides=0, no_slip_bottom=0, bottomDragLinear=1e-3),
    dict(sNx=32, sNy=24, Nr=5, land_frac=0.1, vectorInvariantMomentum=1, selectVortScheme=3, selectKEscheme=1, momDissip_In_AB=0),
    dict(sNx=32, sNy=24, Nr=4, nSx=2, nSy=1, land_frac=0.1, vectorInvariantMomentum=1, selectVortScheme=2, highOrderVorticity=1),
    # non-AB advection schemes: GAD_ADVECTION (multi-dimensional, reference default) or the 1-D form, forward in time
    dict(sNx=40, sNy=36, Nr=6, land_frac=0.1, OL=4, tempAdvScheme=7),
    dict(sNx=24, sNy=16, Nr=4, nSx=2, nSy=2, land_frac=0.2, OL=3, tempAdvScheme=77, gad_multidim_compressible=1, buoyancyLinear=1),
    dict(sNx=40, sNy=36, Nr=6, land_frac=0.1, OL=3, tempAdvScheme=33, multiDimAdvection=0),
    # SALT_INTEGRATE beside TEMP_INTEGRATE: its own scheme (multi-dimensional DST3), diffusivities, and sBeta in the EOS
    dict(sNx=24, sNy=16, Nr=5, nSx=2, nSy=2, land_frac=0.2, OL=3, saltStepping=1, saltAdvScheme=33, diffKhS=5e2, diffKrS=2e-5,
         buoyancyLinear=1, sBeta=7.4e-4, salt=True),
    dict(sNx=40, sNy=24, Nr=6, land_frac=0.1, saltStepping=1, diffKhS=1e3, diffKrS=1e-5, buoyancyLinear=1, sBeta=7.4e-4,
         ivdc_kappa=1.0, salt=True),
    # implicitViscosity: MOM_{U,V}_IMPLICIT_R on u*, v* after the explicit tendencies (flux form and vector invariant)
    dict(sNx=24, sNy=16, Nr=6, nSx=2, nSy=2, land_frac=0.2, implicitViscosity=1, viscAr=5e-2),
    dict(sNx=40, sNy=24, Nr=5, land_frac=0.1, implicitViscosity=1, viscAr=5e-2, vectorInvariantMomentum=1, buoyancyLinear=1),
], ids=["tiles-land", "barotropic", "dst3fl-biharm", "flat", "buoyancy-flat", "buoyancy-land", "buoyancy-ivdc",
        "vecinv-flat", "vecinv-land-tiles-buoyancy", "vecinv-absvort", "vecinv-generic", "vecinv-c4", "multidim-os7mp",
        "multidim-dst3-compressible-tiles", "dst3fl-1d", "salt-dst3-tiles", "salt-c2-ivdc", "implvisc-fluxform", "implvisc-vecinv"])
def test_forward_step_matches_oracle(cfg):
    cfg = dict(cfg)
    with_salt = cfg.pop("salt", False)
    g, P, s = make_channel(**cfg)
    if with_salt:      # salinity: smooth stratification + noise, wet points only
        rng = np.random.default_rng(77)
        s["salt"] = (35.0 + np.linspace(-0.5, 0.5, g.d.Nr)[None, None, :, None, None] + 0.05 * rng.standard_normal(g.d.shape3)) * g.maskC
        s["sRef"] = np.full(g.d.Nr, 35.0)
    co = ChannelOracle(g, P, s)
    m = Model(g, P, s, co.op)
    try:
        jj, ii = g.d.interior()
        for it in range(4):
            ro = co.step()
            rg = m.step()
            assert abs(rg["numIters"] - ro["numIters"]) <= 1, it
            assert rg["firstResidual"] == pytest.approx(ro["firstResidual"], rel=1e-9), it
            for n in ("uVel", "vVel", "wVel", "etaN") + (("theta",) if P["tempStepping"] else ()) + (("salt",) if with_salt else ()):
                a, b = m.get(n), co.s[n]
                assert rel(a[..., jj, ii], b[..., jj, ii]) < 1e-9, (it, n, "interior")
                assert rel(a, b) < 1e-9, (it, n, "halo")
            if it == 0:   # before CG2D feeds back: tendencies are point-wise identical
                a, b = m.get("gU"), co.s["gU"]
                sl = (Ellipsis, slice(g.d.OLy - 1, g.d.OLy + g.d.sNy + 1), slice(g.d.OLx - 1, g.d.OLx + g.d.sNx + 1))
                assert np.array_equal(a[sl], b[sl]), rel(a[sl], b[sl])      # point-wise code: bit-identical by design
                if P["tempStepping"] and P.get("tempAdvScheme", 2) in (2, 3, 4):
                    assert np.array_equal(m.get("gtNm1")[..., jj, ii], co.s["gtNm1"][..., jj, ii])
                # after one solve the fields carry the solver's summation-order noise only: north-star tolerance
                for n in ("uVel", "vVel", "etaN") + (("theta",) if P["tempStepping"] else ()):
                    assert rel(m.get(n)[..., jj, ii], co.s[n][..., jj, ii]) < 1e-12, (n, "step 0")
    finally:
        m.close()


def test_theta_field_ptr_after_steps():
    """include/mitgcm_b200.h: CYCLE_TRACER is a pointer swap, so the device address of MG_THETA alternates between
    two buffers and must be queried again after every step; the freshly queried address holds the current theta,
    the one held from the step before holds the previous theta."""
    import torch
    from mitgcm_b200 import _lib, runtime as rt
    g, P, s = make_channel(sNx=32, sNy=24, Nr=5, land_frac=0.1)
    co = ChannelOracle(g, P, s)
    m = Model(g, P, s, co.op)

    def read_raw(ptr, n):          # what a resident caller does with the address: a device-to-device copy
        out = torch.empty(n, dtype=torch.float64, device="cuda")
        from cuda.bindings import runtime as cudart
        err, = cudart.cudaMemcpy(out.data_ptr(), ptr, n * 8, cudart.cudaMemcpyKind.cudaMemcpyDeviceToDevice)
        assert err == cudart.cudaError_t.cudaSuccess, err
        torch.cuda.synchronize()
        return out.cpu().numpy()
    try:
        L = _lib.lib()
        n3 = int(np.prod(g.d.shape3))
        seen, prev_theta = [], m.get("theta")
        for it in range(3):
            held = L.mitgcm_b200_field_ptr(rt.field_id("theta"))
            m.step()
            rt.sync()
            p = L.mitgcm_b200_field_ptr(rt.field_id("theta"))
            seen.append(p)
            cur = m.get("theta")
            assert np.array_equal(read_raw(p, n3).reshape(g.d.shape3), cur)
            jj, ii = g.d.interior()
            assert np.array_equal(read_raw(held, n3).reshape(g.d.shape3)[..., jj, ii], prev_theta[..., jj, ii])
            assert not np.array_equal(cur, prev_theta)
            prev_theta = cur
        assert seen[0] != seen[1] and seen[0] == seen[2], "theta / theta2 alternate with every step"
    finally:
        m.close()


def test_vecinv_pipelined_kernel_is_bit_identical_to_the_generic_kernels(monkeypatch):
    """The same 3 steps with the cp.async pipelined MOM_VECINV kernel, the shared-memory patch kernel and the
    re-evaluating kernel: the fields must be bit-identical (same expression order everywhere)."""
    outs = []
    for env in ({}, {"MITGCM_B200_VI_NOPIPE": "1"}, {"MITGCM_B200_VI_NOPIPE": "1", "MITGCM_B200_VI_NOTILE": "1"}):
        for k in ("MITGCM_B200_VI_NOPIPE", "MITGCM_B200_VI_NOTILE"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        g, P, s = make_channel(sNx=48, sNy=40, Nr=6, nSx=2, nSy=1, land_frac=0.15, vectorInvariantMomentum=1, buoyancyLinear=1,
                               selectVortScheme=0, selectCoriScheme=1)
        co = ChannelOracle(g, P, s)
        m = Model(g, P, s, co.op)
        try:
            for _ in range(3):
                m.step()
            outs.append({n: m.get(n) for n in ("uVel", "vVel", "wVel", "etaN", "gU", "gV", "guNm1", "gvNm1")})
        finally:
            m.close()
    for o in outs[1:]:
        for n in o:
            assert np.array_equal(o[n], outs[0][n]), n


@pytest.mark.parametrize("opts", [
    dict(buoyancyLinear=1),
    dict(buoyancyLinear=0, selectCoriScheme=2, momDissip_In_AB=0, no_slip_sides=0, bottomDragLinear=1e-3),
    dict(buoyancyLinear=1, rigidLid=1, no_slip_bottom=0, implicitViscosity=1, viscAr=5e-2),
], ids=["headline", "cori2-freeslip-botdrag", "rigidlid-implvisc"])
def test_dyn_tma_kernel_is_identical_to_the_cp_async_and_generic_kernels(monkeypatch, opts):
    """dyn_tma_kernel (operands staged by TMA + mbarrier ring) vs dyn_pipe_kernel (cp.async ring) vs the generic
    dyn_kernel<0>: the same fields after 3 steps on a partial-cell grid with land and 2 x 1 tiles (np.array_equal:
    the TMA kernel skips the identically-zero biharmonic terms, which can only flip the sign of a zero)."""
    outs = []
    # default: the role-split TMA kernel (dyn_tma_uv_kernel: U and V on two thread groups); then the 256-thread one
    NC = {"MITGCM_B200_NO_COLGEOM": "1"}      # the 3-D-array form of every kernel (default: geometry per column, colgeom.cu)
    for env in ({}, {"MITGCM_B200_NO_PHIFUSE": "1"}, {"MITGCM_B200_DYN_TMA_STAGES": "2"}, {"MITGCM_B200_DYN_TMA_MINB": "1"}, {"MITGCM_B200_DYN_TMA_STAGES": "4"},
                {"MITGCM_B200_DYN_TMA_STAGES": "5"}, NC, dict(NC, MITGCM_B200_DYN_TMA_STAGES="2"),
                dict(NC, MITGCM_B200_DYN_TMA_STAGES="4"), dict(NC, MITGCM_B200_DYN_TMA_STAGES="5"),
                dict(NC, MITGCM_B200_DYN_TMA_NOSPLIT="1"), dict(NC, MITGCM_B200_DYN_NOTMA="1"), {"MITGCM_B200_GENERIC_STEP": "1"}):
        for k in ("MITGCM_B200_DYN_NOTMA", "MITGCM_B200_GENERIC_STEP", "MITGCM_B200_DYN_TMA_NOSPLIT", "MITGCM_B200_DYN_TMA_STAGES",
                  "MITGCM_B200_NO_COLGEOM", "MITGCM_B200_DYN_TMA_MINB", "MITGCM_B200_NO_PHIFUSE"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        g, P, s = make_channel(sNx=70, sNy=44, Nr=7, nSx=2, nSy=1, land_frac=0.15, **opts)
        co = ChannelOracle(g, P, s)
        m = Model(g, P, s, co.op)
        try:
            for _ in range(3):
                m.step()
            outs.append({n: m.get(n) for n in ("uVel", "vVel", "wVel", "etaN", "theta", "gU", "gV", "guNm1", "gvNm1")})
        finally:
            m.close()
    for o in outs[1:]:
        for n in o:
            assert np.array_equal(o[n], outs[0][n]), n


@pytest.mark.parametrize("opts", [dict(buoyancyLinear=1), dict(buoyancyLinear=0, rigidLid=1),
                                  dict(buoyancyLinear=1, vectorInvariantMomentum=1)], ids=["free-surface", "rigid-lid", "vecinv"])
def test_column_geometry_kernels_are_bit_identical_to_the_3d_array_kernels(monkeypatch, opts):
    """csrc/colgeom.cu: with z-level geometry (land, partial bottom cells) the step kernels rebuild hFac / mask /
    recip_hFac per column from (kLow, hLow) instead of reading the nine 3-D arrays; the fields after 3 steps must be
    bit-identical to the run that reads the arrays (MITGCM_B200_NO_COLGEOM=1)."""
    from mitgcm_b200 import _lib
    outs, states = [], []
    for env in ({}, {"MITGCM_B200_NO_COLGEOM": "1"}):
        monkeypatch.delenv("MITGCM_B200_NO_COLGEOM", raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        if opts.get("vectorInvariantMomentum"):      # off by default for MOM_VECINV (measured slower): switch it on here
            monkeypatch.setenv("MITGCM_B200_VI_COLGEOM", "1")
        g, P, s = make_channel(sNx=70, sNy=44, Nr=7, nSx=2, nSy=1, land_frac=0.15, **opts)
        co = ChannelOracle(g, P, s)
        m = Model(g, P, s, co.op)
        try:
            for _ in range(3):
                m.step()
            states.append(_lib.lib().mitgcm_b200_col_geom_state_())
            outs.append({n: m.get(n) for n in ("uVel", "vVel", "wVel", "etaN", "theta", "gU", "gV", "gtNm1", "cg2d_b")})
        finally:
            m.close()
    assert states == [1, 0]          # compressed form in use / never checked
    for n in outs[0]:
        assert np.array_equal(outs[0][n], outs[1][n]), n


def test_column_geometry_is_refused_for_surface_following_thickness():
    """hFac scaled by a smooth column factor (what r* does): not the z-level form -> the general kernels run and the
    step still matches the oracle."""
    from mitgcm_b200 import _lib, runtime as rt
    g, P, s = make_channel(sNx=40, sNy=24, Nr=5, land_frac=0.1)
    f = 1.0 + 0.02 * np.sin(np.linspace(0, 6, g.d.PX))[None, None, None, None, :]
    from mitgcm_b200.grid import set_hfac
    set_hfac(g, g.a["hFacC"] * f, g.a["hFacW"] * f, g.a["hFacS"] * f)
    co = ChannelOracle(g, P, s)
    m = Model(g, P, s, co.op)
    try:
        jj, ii = g.d.interior()
        for it in range(4):          # the check is retried three times, then given up (state -1)
            ro, rg = co.step(), m.step()
            assert abs(ro["numIters"] - rg["numIters"]) <= 1
        assert _lib.lib().mitgcm_b200_col_geom_state_() == -1
        for n in ("uVel", "vVel", "theta", "etaN"):
            assert rel(m.get(n)[..., jj, ii], co.s[n][..., jj, ii]) < 1e-9, n
    finally:
        m.close()


@pytest.mark.parametrize("scheme", [2, 33])
def test_thermo_pipelined_kernel_is_bit_identical_to_the_staged_kernel(monkeypatch, scheme):
    """thermo_pipe_kernel (cp.async ring) vs thermo_fast_kernel vs the generic thermo_kernel: same fields after 3 steps
    (scheme 33: the diffusion step on top of GAD_ADVECTION's tendency, calcAdvection = F)."""
    outs = []
    for env in ({}, {"MITGCM_B200_THERMO_NOPIPE": "1"}, {"MITGCM_B200_GENERIC_STEP": "1"}):
        for k in ("MITGCM_B200_THERMO_NOPIPE", "MITGCM_B200_GENERIC_STEP"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        g, P, s = make_channel(sNx=48, sNy=40, Nr=6, nSx=2, nSy=1, land_frac=0.15, buoyancyLinear=1, OL=3, tempAdvScheme=scheme)
        co = ChannelOracle(g, P, s)
        m = Model(g, P, s, co.op)
        try:
            for _ in range(3):
                m.step()
            outs.append({n: m.get(n) for n in ("theta", "gtNm1", "uVel", "etaN")})
        finally:
            m.close()
    for o in outs[1:]:
        for n in o:
            assert np.array_equal(o[n], outs[0][n]), n


def test_mom_implicit_r_matches_oracle():
    """mom_{u,v}_implicit_r_b200_ through the C ABI (reference argument list + gU / gV, host buffers) against the
    oracle on a masked partial-cell grid: bit-identical (same recurrences, no reductions)."""
    from mitgcm_b200 import runtime as rt
    from oracle.pyoracle import Oracle
    from helpers import make_grid
    g = make_grid(31, 17, 3, nSx=2, nSy=2, Nr=7, seed=5)
    d = g.d
    P = dict(deltaTMom=900.0, implicitViscosity=1)
    o = Oracle(g, P)
    rng = np.random.default_rng(3)
    rt.init(d)
    try:
        rt.set_grid(g)
        rt.set_params(**P)
        for isV in (0, 1):
            fld = rng.standard_normal(d.shape3)
            a, b = fld.copy(), fld.copy()
            for bj in range(1, d.nSy + 1):
                for bi in range(1, d.nSx + 1):
                    kap = 5e-2 * (1 + rng.random((d.Nr + 1, d.PY, d.PX)))
                    assert o.mom_implicit_r(bi, bj, isV, kap, a) == 0
                    rt.mom_implicit_r(kap, bi, bj, b, isV=bool(isV))
            assert np.array_equal(a, b)
            assert not np.array_equal(a, fld)
        rt.set_params(implicitViscosity=0)
        with pytest.raises(rt.B200Error):
            rt.mom_implicit_r(kap, 1, 1, b)
    finally:
        rt.finalize()


@pytest.mark.parametrize("shape", [(10, 7, 3, 3, 2, 4), (33, 18, 4, 1, 1, 2), (8, 8, 2, 2, 3, 1)])
def test_exchange_matches_oracle(shape):
    """EXCH_XYZ_RL: full-width halo incl. corners (exch1_rx.template:170-201)."""
    from mitgcm_b200 import runtime as rt
    from mitgcm_b200.grid import Dims
    from oracle.pyoracle import Oracle
    from helpers import make_grid
    sNx, sNy, OL, nSx, nSy, Nr = shape
    g = make_grid(sNx, sNy, OL, nSx=nSx, nSy=nSy, Nr=Nr, seed=4)
    o = Oracle(g)
    rng = np.random.default_rng(0)
    a3, a2 = rng.standard_normal(g.d.shape3), rng.standard_normal(g.d.shape2)
    rt.init(g.d)
    try:
        rt.set_field("theta", a3)
        rt.set_field("etaN", a2)
        rt.exch("theta")
        rt.exch("etaN")
        b3, b2 = a3.copy(), a2.copy()
        o.exch_xyz(b3, Nr)
        o.exch_xyz(b2, 1)
        assert np.array_equal(rt.get_field("theta", np.zeros_like(a3)), b3)
        assert np.array_equal(rt.get_field("etaN", np.zeros_like(a2)), b2)
    finally:
        rt.finalize()


def test_two_rank_step_matches_single_rank():
    """N > 1 path on real GPUs: 2 ranks (CUDA IPC peer pushes + mailbox all-reduce in CG2D, NCCL
    halo exchange) must reproduce the single-rank run of the same global domain.  Needs 2 GPUs."""
    import os, subprocess, sys, torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run scripts/dist_check.py under torchrun on a multi-GPU box)")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533",
                        os.path.join(root, "scripts", "dist_check.py"), "64", "48", "4", "3"],
                       capture_output=True, text=True, timeout=600)
    assert "DIST_CHECK PASS" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
