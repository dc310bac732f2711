#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native MITgcm hot path.

Workload (BASELINE.json configs[4]): synthetic doubly-periodic channel, 2048 x 2048 x 50 cells PER GPU (weak
scaling; --scaling strong splits ONE 2048 x 2048 x 50 domain over the ranks), FP64, flat bottom, wind-driven,
c2 advection of theta, linear equation of state (theta feeds back on the flow through the hydrostatic pressure
gradient), harmonic viscosity, implicit free surface solved by CG2D to 1e-7 every step.  The Coriolis parameter
f = 1e-4 + 6.5e-5 sin(2 pi y / L) and the smooth part of the initial flow are periodic over one 2048-cell block,
so the weak-scaled domain at N ranks is the exact periodic tiling of the N = 1 problem (same noise on every
block: same flow and same CG2D iteration counts at every N, whatever the summation shape): max f dt = 0.198, the stratification (6 -> 4 degC over 5000 m) keeps the internal-wave CFL of the explicit AB2
scheme at 0.17; tests/test_bench_workload_cpu.py steps it on the CPU oracle (bounded flow over hundreds of steps).
A "step" is one FORWARD_STEP on the resident state: THERMODYNAMICS (GAD_CALC_RHS) + DYNAMICS (MOM_FLUXFORM) +
SOLVE_FOR_PRESSURE (CG2D) + correction/continuity + halo exchanges (mitgcm_b200_forward_step_).

  value : timesteps/s with the state resident in HBM (device timed with CUDA events, max over ranks)
  e2e   : the SAME steps (state restored from a snapshot, same step numbers) through the C ABI with HOST buffers
          every step: surface forcing (2 tile2d fields) host->device from pinned memory, eta (1 tile2d field)
          + solver scalars device->host; e2e_dropin: the stock drop-in call cg2d_b200_ with host cg2d_b / cg2d_x
  roofline : dominant kernel of the step, algorithmic bytes / CUDA-event time, against
             MEASURED_PEAKS.json hbm_gbs
  cpu_baseline / --impl reference : the CPU oracle (C restatement of the reference loops; the Fortran reference
          cannot be built in this image: no Fortran compiler) stepping the SAME workload at full size on the host
          cores (one tile per thread), when the host has the memory for it; otherwise the largest sample that fits.
The run FAILS (exit 1, no value) when the state is non-finite on any rank or CG2D hits cg2dMaxIters.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# algorithmic bytes of SURVEY.md section 8(d) (the yardstick `roofline` and `kernel_hbm_gbs` use)
CG2D_BYTES_PER_POINT_ITER = 136.0     # 17 words, the minimum of the three-sync CG2D structure
DYN_BYTES_PER_CELL = 136.0            # MOM_FLUXFORM behind the reference argument list: 17 words
THERMO_BYTES_PER_CELL = 88.0          # GAD_CALC_RHS + AB2 on the tendency: 11 words per cell and tracer
# (what the kernels actually move is reported beside it from the ncu capture in profiles/r02_traffic.json: with the column
# geometry of csrc/colgeom.cu the 3-D kernels move LESS than the section-8(d) figures, which count the geometry arrays)
from mitgcm_b200.model import BENCH_FSIN_AMP as FSIN_AMP      # f = 1e-4 + 6.5e-5 sin(2 pi y / block length)
T_TOP, T_BOT, T_NOISE = 6.0, 4.0, 0.01


def measured_peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.p, self.index = [], None, index

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].startswith("Active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def params(nr):
    from mitgcm_b200.model import DEFAULTS
    P = dict(DEFAULTS)
    P.update(deltaTMom=1200.0, deltaTFreeSurf=1200.0, deltaTtracer=1200.0, cg2dMaxIters=1000,
             cg2dTargetResidual=1e-7, viscAhD=400.0, viscAhZ=400.0, viscAr=1e-2, diffKhT=1e3, diffKrT=1e-5,
             # eosType = 'LINEAR' (SURVEY.md section 8(d)): theta drives the flow through CALC_PHI_HYD
             buoyancyLinear=1, gravity=9.81, tAlpha=2e-4, sBeta=0.0, rhoNil=1000.0, rhoConst=1000.0,
             # stratification of the initial state (popped by make_channel)
             tTop=T_TOP, tBot=T_BOT, tNoise=T_NOISE)
    return P


def fail_run(rank, why, extra=None):
    """A run that produced a non-finite state or an unconverged solver has no throughput: say why, exit 1."""
    if rank == 0:
        print(json.dumps({"error": why, **(extra or {})}), file=sys.stderr, flush=True)
    sys.exit(1)


# ------------------------------------------------------------------------------------------------
def setup_workload(args, rank, world, local):
    """Initialises the library for this rank's block of the channel and fills the mirrors (grid, operator, state,
    forcing).  world == 1: one block on one GPU; world > 1: needs the NCCL process group (peer wiring)."""
    import types
    import torch
    from mitgcm_b200 import runtime as rt
    from mitgcm_b200.grid import Dims, cartesian_grid
    from mitgcm_b200.model import ini_cg2d, LIB_PARAMS
    from mitgcm_b200.parallel import process_grid
    nPx, nPy = process_grid(world) if world > 1 else (1, 1)
    if world > 1:
        import torch.distributed as dist
        from mitgcm_b200 import distributed
    strong = args.scaling == "strong"
    if strong and (args.nx % nPx or args.ny % nPy):
        raise SystemExit("bench.py: --scaling strong needs nx, ny divisible by the process grid")
    NX, NY, NR = (args.nx // nPx, args.ny // nPy, args.nr) if strong else (args.nx, args.ny, args.nr)
    d = Dims(sNx=NX, sNy=NY, OLx=2, OLy=2, nSx=1, nSy=1, Nr=NR, nPx=nPx, nPy=nPy, myPx=rank % nPx, myPy=rank // nPx)
    P = params(NR)
    if args.momentum == "vecinv":
        P["vectorInvariantMomentum"] = 1
    t_setup = time.time()
    dx = 20e3
    g = cartesian_grid(d, [dx] * d.Nx, [dx] * d.Ny, [100.0] * NR, f0=1e-4, gBaro=9.81, fsin=(FSIN_AMP, args.ny * dx))
    fdt_max = float(np.abs(g.a["fCori"]).max() * P["deltaTMom"])
    P["globalArea"] = float(d.Nx * d.Ny) * dx * dx
    rt.init(d, local)
    rt.set_grid(g)                                   # 2-D metrics; flat bottom: 3-D factors are 1
    for n in "hFacC hFacW hFacS recip_hFacC recip_hFacW recip_hFacS maskC maskW maskS".split():
        rt.fill_field(n, 1.0)
    rt.set_params(**{k: P[k] for k in LIB_PARAMS if k in P})
    rt.set_params(deltaTtracer=P["deltaTtracer"], tempAdvScheme=args.temp_adv_scheme, tempVertAdvScheme=args.temp_adv_scheme,
                  nIter0=0, profile=1)
    if world > 1:
        distributed.setup(d)

    def wrap(dd, a):
        # uniform flat-bottom grid: every rank's operator is identical and doubly periodic, so the
        # halo of the operator arrays equals the rank-local periodic wrap
        ox, oy, sx, sy = dd.OLx, dd.OLy, dd.sNx, dd.sNy
        a[..., oy:oy + sy, 0:ox] = a[..., oy:oy + sy, sx:sx + ox]
        a[..., oy:oy + sy, ox + sx:] = a[..., oy:oy + sy, ox:2 * ox]
        a[..., 0:oy, :] = a[..., sy:sy + oy, :]
        a[..., oy + sy:, :] = a[..., oy:2 * oy, :]
        return a
    rt.set_cg2d_operator(ini_cg2d(g, P, hfac_flat=1.0, exch=wrap))
    # state generated on the device per global index; smooth part periodic over one args.nx x args.ny block
    dev = torch.device("cuda", local)
    gen = torch.Generator(device=dev)
    gen.manual_seed(20261018 + (rank if strong else 0))      # weak scaling: every block carries the same noise
    jg = (torch.arange(d.PY, device=dev, dtype=torch.float64) - d.OLy + d.myPy * NY + 0.5) / args.ny
    ig = (torch.arange(d.PX, device=dev, dtype=torch.float64) - d.OLx + d.myPx * NX + 0.5) / args.nx
    YY, XX = torch.meshgrid(jg, ig, indexing="ij")
    two_pi = 2 * np.pi
    psi_u = 0.1 * torch.sin(two_pi * XX) * torch.cos(2 * two_pi * YY) + 0.03 * torch.cos(3 * two_pi * XX) * torch.sin(two_pi * YY)
    psi_v = 0.1 * torch.cos(two_pi * XX) * torch.sin(2 * two_pi * YY) - 0.03 * torch.sin(3 * two_pi * XX) * torch.cos(two_pi * YY)
    prof = torch.linspace(1.0, 0.2, NR, device=dev, dtype=torch.float64)[:, None, None]
    tref = torch.linspace(T_TOP, T_BOT, NR, device=dev, dtype=torch.float64)[:, None, None]
    for name, base, amp in (("uVel", psi_u[None] * prof, 1e-3), ("vVel", psi_v[None] * prof, 1e-3),
                            ("theta", tref.expand(NR, d.PY, d.PX), T_NOISE)):
        f = (base + amp * torch.randn((NR, d.PY, d.PX), device=dev, dtype=torch.float64, generator=gen)).contiguous()
        torch.cuda.synchronize()
        rt.set_field(name, f)
        del f
    rt.set_field("etaN", (0.1 * torch.sin(two_pi * XX) * torch.cos(two_pi * YY)).contiguous())
    for n in ("wVel", "gU", "gV", "guNm1", "gvNm1", "gtNm1", "theta2", "cg2d_b", "cg2d_x"):
        rt.fill_field(n, 0.0)
    rt.fill_field("kappaRU", P["viscAr"])
    rt.fill_field("kappaRV", P["viscAr"])
    rt.fill_field("kappaRT", P["diffKrT"])
    # linear equation of state: reference profile and the vertical grid CALC_PHI_HYD integrates over
    vert = lambda a: np.concatenate([np.asarray(a, dtype=np.float64), np.zeros(NR + 1 - len(a))])
    rt.set_field("tRef", vert(np.linspace(T_TOP, T_BOT, NR)))
    rt.set_field("sRef", vert(np.zeros(NR)))
    rt.set_field("rF", vert(g.a["rF"]))
    rt.set_field("rC", vert(g.a["rC"]))
    tau = (-0.1 * torch.cos(two_pi * YY) * (1.0 / 1000.0)).contiguous()
    sfU_host = tau.cpu().pin_memory()
    sfV_host = torch.zeros_like(sfU_host).pin_memory()
    eta_host = torch.empty_like(sfU_host).pin_memory()
    del psi_u, psi_v, XX, YY, tau
    rt.set_field("surfForcU", sfU_host)
    rt.set_field("surfForcV", sfV_host)
    if world > 1:
        dist.barrier()
        distributed.exchange("uVel", "vVel", "theta", "etaN")
    else:
        for n in ("uVel", "vVel", "theta", "etaN"):
            rt.exch(n)
    t_setup = time.time() - t_setup
    return types.SimpleNamespace(**{k: v for k, v in locals().items() if k not in ("types",)})


# ------------------------------------------------------------------------------------------------
def run_cuda(args, rank, world):
    import ctypes as C
    import torch
    from mitgcm_b200 import _lib, runtime as rt
    from mitgcm_b200.grid import Dims, cartesian_grid
    from mitgcm_b200.model import ini_cg2d, LIB_PARAMS
    from mitgcm_b200.parallel import process_grid

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 path has no CPU fallback")
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    nPx, nPy = process_grid(world) if world > 1 else (1, 1)
    dist = None
    multi_check = None
    if world > 1:
        import torch.distributed as dist
        from mitgcm_b200 import distributed
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        if not args.no_selfcheck:
            # N-rank vs 1-rank parity of the step on a small global domain (both exchanges, CG2D across ranks)
            multi_check = distributed.selfcheck(64, 48, 4, 10)
            if not multi_check["ok"]:
                fail_run(rank, "multi-rank self-check failed (N ranks vs 1 rank of the same global domain)", multi_check)
    W = setup_workload(args, rank, world, local)
    strong, NX, NY, NR, d, P, dev, fdt_max, t_setup = W.strong, W.NX, W.NY, W.NR, W.d, W.P, W.dev, W.fdt_max, W.t_setup
    sfU_host, sfV_host, eta_host = W.sfU_host, W.sfV_host, W.eta_host

    def barrier():
        rt.sync()
        if world > 1:
            dist.barrier()
        rt.sync()

    step = distributed.forward_step if world > 1 else rt.forward_step
    STATE = ("uVel", "vVel", "wVel", "theta", "guNm1", "gvNm1", "gtNm1", "etaN")
    shape = lambda n: (d.PY, d.PX) if n == "etaN" else (NR, d.PY, d.PX)

    def health():
        """max |u|, |v|, |eta| and finiteness over ALL ranks."""
        buf3 = torch.empty((NR, d.PY, d.PX), device=dev, dtype=torch.float64)
        buf2 = torch.empty((d.PY, d.PX), device=dev, dtype=torch.float64)
        v, mism = [], torch.zeros((), device=dev, dtype=torch.float64)
        oy, ox = d.OLy, d.OLx
        for n in ("uVel", "vVel", "theta", "etaN"):
            b = buf2 if n == "etaN" else buf3
            rt.get_field(n, b)           # copy on the library's stream (synchronised); the reduction runs on torch's
            v.append(b.abs().max() if n != "theta" else (b - 5.0).abs().max())
            if not strong:
                # weak scaling: every block holds the same data bit for bit (same arithmetic on the same numbers, the
                # rank-ordered global sums are identical everywhere), so a correct exchange -- whatever the transport
                # and the rank count -- leaves in my halo exactly my own opposite edge
                w = torch.maximum((b[..., oy:oy + NY, :ox] - b[..., oy:oy + NY, NX:NX + ox]).abs().max(),
                                  (b[..., oy:oy + NY, ox + NX:] - b[..., oy:oy + NY, ox:2 * ox]).abs().max())
                w = torch.maximum(w, (b[..., :oy, ox:ox + NX] - b[..., NY:NY + oy, ox:ox + NX]).abs().max())
                w = torch.maximum(w, (b[..., oy + NY:, ox:ox + NX] - b[..., oy:2 * oy, ox:ox + NX]).abs().max())
                w = torch.maximum(w, (b[..., :oy, :ox] - b[..., NY:NY + oy, NX:NX + ox]).abs().max())          # SW corner
                w = torch.maximum(w, (b[..., oy + NY:, ox + NX:] - b[..., oy:2 * oy, ox:2 * ox]).abs().max())  # NE corner
                mism = torch.maximum(mism, torch.nan_to_num(w, nan=1e300))
            torch.cuda.synchronize()     # ... and must finish before the buffer is overwritten by the next field
        t = torch.stack(v)
        bad = (~torch.isfinite(t)).any().to(torch.float64).reshape(1)
        t = torch.cat([torch.nan_to_num(t, nan=1e300, posinf=1e300), bad, mism.reshape(1)])
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t = t.cpu().numpy()
        del buf3, buf2
        return {"max_abs_u": float(t[0]), "max_abs_v": float(t[1]), "max_abs_theta_minus_5": float(t[2]),
                "max_abs_eta": float(t[3]), "finite": bool(t[4] == 0.0),
                "halo_vs_own_periodic_edge": (float(t[5]) if not strong else None)}

    it = 0
    for _ in range(args.warmup):
        step(it)
        it += 1
    h0 = health()
    # snapshot: the e2e region replays exactly the steps of the resident region
    snap = {}
    for n in STATE:
        snap[n] = torch.empty(shape(n), device=dev, dtype=torch.float64)
        rt.get_field(n, snap[n])
    it_snap = it
    # ---- timed region 1: resident state ---------------------------------------------------------
    L = _lib.lib()
    L.mitgcm_b200_launch_count_.restype = C.c_longlong
    clk = ClockSampler(local)
    barrier()
    n0 = L.mitgcm_b200_launch_count_()
    clk.start()
    L.mitgcm_b200_event_record_(C.byref(C.c_int(0)))
    iters, phase = [], np.zeros(7)
    ms7 = (C.c_double * 7)()
    for _ in range(args.steps):
        r = step(it)
        it += 1
        iters.append(r["numIters"])
        L.mitgcm_b200_step_timings_(ms7)
        phase += np.array(list(ms7))
    L.mitgcm_b200_event_record_(C.byref(C.c_int(1)))
    ms = C.c_double()
    L.mitgcm_b200_event_elapsed_ms_(C.byref(C.c_int(0)), C.byref(C.c_int(1)), C.byref(ms))
    barrier()
    clocks = clk.stop()
    launches = int(L.mitgcm_b200_launch_count_() - n0)
    dev_ms = ms.value
    h1 = health()
    # ---- timed region 2: end to end through the C ABI with host buffers, same steps ---------------
    for n in STATE:
        rt.set_field(n, snap[n])
    if world > 1:
        dist.barrier()
    it = it_snap
    iters_e2e = []
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        rt.set_field("surfForcU", sfU_host)           # H2D from pinned memory
        rt.set_field("surfForcV", sfV_host)
        r = step(it)
        it += 1
        iters_e2e.append(r["numIters"])
        rt.get_field("etaN", eta_host)                # D2H of the step's result
    rt.sync()
    e2e_s = time.perf_counter() - t0
    # ---- stock drop-in call: cg2d_b200_ with HOST cg2d_b / cg2d_x (solve_for_pressure.F:292) -------
    dropin = None
    if world == 1 and not args.no_dropin:
        b_host = np.zeros(d.shape2)
        rt.get_field("cg2d_b", b_host)                # the normalised right-hand side of the last step
        # the caller's arrays have fixed addresses (COMMON /SOLVE_FOR_PRESSURE/ cg2d_b, cg2d_x) and are page-locked once
        bb, xx = np.empty(d.shape2), np.empty(d.shape2)
        rt.pin_host(bb)
        rt.pin_host(xx)
        bb[...] = b_host
        xx[...] = 0.0
        rt.cg2d(bb, xx, int(P["cg2dMaxIters"]))       # warm-up (staging buffers)
        nrep, its_d, dt = 3, [], 0.0
        for _ in range(nrep):
            bb[...] = b_host                          # SOLVE_FOR_PRESSURE refills them (not timed: CPU work of the caller)
            xx[...] = 0.0
            t0 = time.perf_counter()
            its_d.append(rt.cg2d(bb, xx, int(P["cg2dMaxIters"]))["numIters"])
            dt += time.perf_counter() - t0
        dt /= nrep
        dropin = {"call": "cg2d_b200_(cg2d_b, cg2d_x, ...) with page-locked host arrays, zero first guess", "ms_per_solve": dt * 1e3,
                  "iters": its_d[0], "iters_per_s": its_d[0] / dt, "h2d_bytes_per_solve": int(2 * b_host.nbytes),
                  "d2h_bytes_per_solve": int(2 * b_host.nbytes)}
    t = torch.tensor([dev_ms, e2e_s, float(max(iters + iters_e2e))], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_s, worst_iters = float(t[0]), float(t[1]), int(t[2])
    finite = h0["finite"] and h1["finite"] and bool(np.isfinite(eta_host.numpy()).all())
    if not finite:
        fail_run(rank, "non-finite model state", {"before": h0, "after": h1})
    # (the smooth part and f are evaluated at global coordinates: blocks agree to round-off, hence 1e-9, not 0)
    if not strong and max(h0["halo_vs_own_periodic_edge"], h1["halo_vs_own_periodic_edge"]) > 1e-9:
        fail_run(rank, "halo exchange left halos that differ from the periodic image of the block", {"before": h0, "after": h1})
    if worst_iters >= int(P["cg2dMaxIters"]):
        fail_run(rank, "CG2D hit cg2dMaxIters", {"iters": iters, "iters_e2e": iters_e2e})
    if rank != 0:
        if world > 1:
            distributed.teardown()
        rt.finalize()
        return
    K = args.steps
    ms_per_step = dev_ms / K
    steps_per_s = K / (dev_ms / 1e3)
    peak, peak_kind = measured_peak()
    cells = NX * NY * NR
    tot_iters = int(sum(iters))
    names = ["thermo", "dyn", "rhs", "cg2d", "eta", "corr", "exch"]
    shares = {n: float(phase[i] / max(phase.sum(), 1e-9)) for i, n in enumerate(names)}
    cand = {
        "cg2d_kernel": (CG2D_BYTES_PER_POINT_ITER * NX * NY * tot_iters / K, phase[3] / K),
        "dyn_kernel": (DYN_BYTES_PER_CELL * cells, phase[1] / K),
        "thermo_kernel": (THERMO_BYTES_PER_CELL * cells, phase[0] / K),
    }
    gbs = lambda nbytes, ms_: float(nbytes / (max(ms_, 1e-9) * 1e-3) / 1e9)
    dom = max(cand, key=lambda k: cand[k][1])
    ach = gbs(*cand[dom])
    traffic, traffic_src, moved = None, None, {}
    try:        # DRAM bytes per launch from the committed ncu capture of this workload (not measured in this run)
        tj = json.load(open(os.path.join(ROOT, "profiles", TRAFFIC_FILE)))
        if tj.get("workload") == f"{NX}x{NY}x{NR}" and not os.environ.get("MITGCM_B200_NO_COLGEOM"):
            if dom in tj:
                traffic = tj[dom]["dram_bytes_per_launch"]
                traffic_src = f"profiles/{TRAFFIC_FILE} (ncu --set full capture of this workload, not measured in this run)"
            # fraction of the peak on the bytes ncu saw move (CG2D: scaled to this run's iteration count)
            for k_, ms_ in (("cg2d_kernel", phase[3] / K), ("dyn_kernel", phase[1] / K), ("thermo_kernel", phase[0] / K)):
                if k_ in tj:
                    nb = tj[k_]["dram_bytes_per_launch"] * ((tot_iters / K) / 150.0 if k_ == "cg2d_kernel" else 1.0)
                    moved[k_] = gbs(nb, ms_) / peak
    except Exception:
        pass
    mult = 1 if strong else world
    unit = "timesteps/s" if strong else "block-timesteps/s"
    out = {
        "metric": "timesteps/s at 2048x2048x50" + (" (one global domain split over the ranks)" if strong else
                                                    " (per-GPU block of the weak-scaled channel; aggregate = ranks x steps/s)"),
        "value": steps_per_s * mult, "unit": unit, "n_gpus": world, "steps": K, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"synthetic doubly-periodic channel {NX}x{NY}x{NR} per GPU"
                               + (f" (strong scaling of {args.nx}x{args.ny}x{NR})" if strong else "")
                               + ", FP64, flat bottom, c2 advection + harmonic viscosity, linear EOS, cg2dTargetResidual 1e-7, "
                               f"f = 1e-4 + {FSIN_AMP:g} sin(2 pi y / {args.ny} cells) (max f dt {fdt_max:.3f})"
                               + (", MOM_VECINV dynamics" if args.momentum == "vecinv" else "")
                               + (f", tempAdvScheme {args.temp_adv_scheme}" if args.temp_adv_scheme != 2 else ""),
                   "process_grid": f"{nPx}x{nPy}", "l2_policy": "working set 30 GB per GPU >> 126 MB L2, no flush needed",
                   "cells_per_gpu": cells,
                   "transport": (os.environ.get("MITGCM_B200_TRANSPORT", "peer") if world > 1 else "none")},
        "cg2d": {"iters_per_step": tot_iters / K, "iters_per_step_e2e": float(sum(iters_e2e)) / K,
                 "iters_per_s": tot_iters / max(phase[3] * 1e-3, 1e-12),
                 "us_per_iter": phase[3] * 1e3 / max(tot_iters, 1),
                 "hbm_gbs": gbs(CG2D_BYTES_PER_POINT_ITER * NX * NY * tot_iters, phase[3])},
        "phase_share": shares, "phase_ms_per_step": {n: float(phase[i] / K) for i, n in enumerate(names)},
        "kernel_hbm_gbs": {k: gbs(*v) for k, v in cand.items()},
        "kernel_frac_of_peak": {k: gbs(*v) / peak for k, v in cand.items()},
        "kernel_frac_of_peak_moved_bytes": moved or None,
        "step_hbm_frac": gbs((DYN_BYTES_PER_CELL + THERMO_BYTES_PER_CELL) * cells
                             + CG2D_BYTES_PER_POINT_ITER * NX * NY * tot_iters / K, ms_per_step) / peak,
        "roofline": {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                     "peak_source": peak_kind, "traffic": traffic, "traffic_source": traffic_src,
                     "algorithmic_bytes_per_launch": cand[dom][0]},
        "e2e": {"value": mult * K / e2e_s, "unit": unit, "h2d_bytes_per_step": int(2 * sfU_host.numel() * 8),
                "d2h_bytes_per_step": int(eta_host.numel() * 8 + 24), "same_steps_as_value": True},
        "gpu_launches": launches, "clocks": clocks, "setup_s": t_setup, "finite": finite,
        "health": {"after_warmup": h0, "after_timed": h1, "max_f_dt": fdt_max},
    }
    if dropin:
        out["e2e_dropin"] = dropin
    if multi_check:
        out["multi_rank_check"] = multi_check
    if world > 1:
        distributed.teardown()
    rt.finalize()
    if world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_arm(args, steps=2, warmup=1)
    print(json.dumps(out), flush=True)


TRAFFIC_FILE = "r02_traffic.json"


# ------------------------------------------------------------------------------------------------
def cpu_arm(args, steps, warmup, budget_s=240.0):
    """The CPU oracle (gcc -O2 C restatement of the reference loops, one tile per host thread) stepping the bench
    workload.  Full size (args.nx x args.ny x args.nr) when the host has the memory (about 45 GB at 2048^2 x 50);
    otherwise the largest power-of-two fraction that fits, scaled by cell count and said so."""
    from mitgcm_b200.model import make_channel
    from oracle.channel import ChannelOracle
    cores = os.cpu_count() or 1
    nt = 1
    while nt * 2 <= cores:
        nt *= 2
    nSx = 1
    while nSx * nSx < nt:
        nSx *= 2
    nSy = nt // nSx
    try:
        import psutil
        avail = psutil.virtual_memory().available
    except Exception:
        avail = 32e9
    sx, sy = args.nx, args.ny
    need = lambda x, y: 30.0 * 8.0 * (x + 4 * nSx) * (y + 4 * nSy) * args.nr
    while need(sx, sy) > 0.8 * avail and sx > 64 and sy > 64:
        if sx >= sy:
            sx //= 2
        else:
            sy //= 2
    sNx, sNy = max(8, sx // nSx), max(8, sy // nSy)
    P = params(args.nr)
    t_setup = time.perf_counter()
    g, P2, s = make_channel(sNx, sNy, args.nr, nSx=nSx, nSy=nSy, block=(args.nx, args.ny), **P)
    co = ChannelOracle(g, P2, s, threads=nt)
    t_setup = time.perf_counter() - t_setup
    del s
    t0 = time.perf_counter()
    for _ in range(warmup):
        co.step()
    t_warm = time.perf_counter() - t0
    est = t_warm / max(warmup, 1) if warmup else None
    n_run = steps
    if est is not None and est * steps > budget_s:          # keep the arm inside the driver's time limit
        n_run = max(1, int(budget_s / est))
    its = []
    t0 = time.perf_counter()
    for _ in range(n_run):
        its.append(co.step()["numIters"])
    dt = time.perf_counter() - t0
    umax = float(max(np.abs(co.s["uVel"]).max(), np.abs(co.s["vVel"]).max()))
    full = (sNx * nSx, sNy * nSy) == (args.nx, args.ny)
    cells = sNx * nSx * sNy * nSy * args.nr
    scale = cells / float(args.nx * args.ny * args.nr)
    return {"value": n_run / dt * scale, "unit": "block-timesteps/s", "cores": nt, "kind": "port",
            "sample": f"oracle (C restatement of the reference loops, gcc -O2, {nt} tiles on {nt} threads) stepping "
                      f"{sNx * nSx}x{sNy * nSy}x{args.nr} for {n_run} steps after {warmup} warm-up ({dt:.1f} s, "
                      f"{np.mean(its):.0f} CG iters/step)"
                      + ("" if full else f", scaled to {args.nx}x{args.ny}x{args.nr} by cell count ({scale:.5f}): "
                                         f"host memory {avail / 1e9:.0f} GB available < {need(args.nx, args.ny) / 1e9:.0f} GB needed"),
            "same_config": bool(full), "steps_run": n_run, "ms_per_step": dt / n_run * 1e3, "setup_s": t_setup,
            "max_abs_uv": umax, "finite": bool(np.isfinite(umax))}


def run_reference(args, rank, world):
    if rank != 0:
        return
    t0 = time.perf_counter()
    cb = cpu_arm(args, steps=max(1, args.steps), warmup=max(1, min(args.warmup, 2)))
    out = {"impl": "reference",
           "metric": "timesteps/s at 2048x2048x50 (per-GPU block of the weak-scaled channel; aggregate = ranks x steps/s)",
           "value": cb["value"], "unit": "block-timesteps/s", "n_gpus": world, "steps": cb["steps_run"], "warmup": args.warmup,
           "ms_per_step": cb["ms_per_step"] if cb["same_config"] else 1e3 / cb["value"], "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": f"synthetic doubly-periodic channel {args.nx}x{args.ny}x{args.nr}, ONE block on the host "
                                  f"cores (CPU oracle; the Fortran reference cannot be built here: no Fortran compiler). "
                                  f"At N > 1 this is still one block on one host: compare with the N = 1 GPU line only"},
           "cpu_baseline": cb,
           "e2e": {"value": cb["value"], "unit": "block-timesteps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "wall_s": time.perf_counter() - t0}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--nx", type=int, default=2048)
    ap.add_argument("--ny", type=int, default=2048)
    ap.add_argument("--nr", type=int, default=50)
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: nx x ny x nr per GPU; strong: ONE nx x ny x nr domain split over the process grid")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-dropin", action="store_true")
    ap.add_argument("--no-selfcheck", action="store_true")
    ap.add_argument("--temp-adv-scheme", type=int, default=2,
                    help="tempAdvScheme (33, 77, 7, ...: GAD_ADVECTION multi-dimensional advection; not the headline workload)")
    ap.add_argument("--momentum", default="fluxform", choices=["fluxform", "vecinv"],
                    help="vecinv: MOM_VECINV instead of MOM_FLUXFORM in the dynamics (not the headline workload)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_cuda(args, rank, world)


if __name__ == "__main__":
    main()
