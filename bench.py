#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native MITgcm hot path.

Workload (BASELINE.json configs[4]): synthetic doubly-periodic channel, 2048 x 2048 x 50 cells PER
GPU (weak scaling), FP64, flat bottom, wind-driven, c2 advection of theta, linear equation of state
(theta feeds back on the flow through the hydrostatic pressure gradient), harmonic viscosity,
implicit free surface solved by CG2D to 1e-7 every step.  A "step" is one FORWARD_STEP on the
resident state: THERMODYNAMICS (GAD_CALC_RHS) + DYNAMICS (MOM_FLUXFORM) + SOLVE_FOR_PRESSURE (CG2D)
+ correction/continuity + halo exchanges (mitgcm_b200_forward_step_).

  value : timesteps/s with the state resident in HBM (device timed with CUDA events)
  e2e   : same metric through the C ABI with HOST buffers every step: surface forcing (2 tile2d
          fields) host->device from pinned memory, eta (1 tile2d field) + solver scalars device->host
  roofline : dominant kernel of the step, algorithmic bytes / CUDA-event time, against
             MEASURED_PEAKS.json hbm_gbs
  cpu_baseline / --impl reference : the CPU oracle (C restatement of the reference loops; the
          Fortran reference cannot be built in this image) stepping a bounded sample of the same
          workload on the host cores, scaled by cell count.
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

CG2D_BYTES_PER_POINT_ITER = 136.0     # 17 words, the minimum of the three-sync CG2D structure
                                      # (this implementation moves 14-16 words: DESIGN.md section 5)
DYN_BYTES_PER_CELL = 160.0            # 20 words (MOM_FLUXFORM + TIMESTEP fused: 19, + phiHyd of CALC_GRAD_PHI_HYD)
THERMO_BYTES_PER_CELL = 112.0         # thermo phase: 12 words (GAD_CALC_RHS + AB2 + TIMESTEP_TRACER fused) + 2 words
                                      # (phihyd_kernel: R theta, W phiHyd; linear EOS evaluated on the fly)


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
             buoyancyLinear=1, gravity=9.81, tAlpha=2e-4, sBeta=0.0, rhoNil=1000.0, rhoConst=1000.0)
    return P


# ------------------------------------------------------------------------------------------------
def run_cuda(args, rank, world):
    import torch
    from mitgcm_b200 import runtime as rt
    from mitgcm_b200.grid import Dims, cartesian_grid
    from mitgcm_b200.model import ini_cg2d, LIB_PARAMS
    from mitgcm_b200.parallel import process_grid

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 path has no CPU fallback")
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    nPx, nPy = process_grid(world)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    NX, NY, NR = args.nx, args.ny, args.nr
    d = Dims(sNx=NX, sNy=NY, OLx=2, OLy=2, nSx=1, nSy=1, Nr=NR, nPx=nPx, nPy=nPy, myPx=rank % nPx, myPy=rank // nPx)
    P = params(NR)
    if args.momentum == "vecinv":
        P["vectorInvariantMomentum"] = 1
    t_setup = time.time()
    g = cartesian_grid(d, [20e3] * d.Nx, [20e3] * d.Ny, [100.0] * NR, f0=1e-4, beta=1e-11, gBaro=9.81)
    P["globalArea"] = float(NX * NY * world) * 20e3 * 20e3
    rt.init(d, local)
    rt.set_grid(g)                                   # 2-D metrics; flat bottom: 3-D factors are 1
    for n in "hFacC hFacW hFacS recip_hFacC recip_hFacW recip_hFacS maskC maskW maskS".split():
        rt.fill_field(n, 1.0)
    rt.set_params(**{k: P[k] for k in LIB_PARAMS if k in P})
    rt.set_params(deltaTtracer=P["deltaTtracer"], tempAdvScheme=args.temp_adv_scheme, tempVertAdvScheme=args.temp_adv_scheme,
                  nIter0=0, profile=1)
    if world > 1:
        from mitgcm_b200 import distributed
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
    # state generated on the device per global index (same global field for any rank count)
    dev = torch.device("cuda", local)
    gen = torch.Generator(device=dev)
    gen.manual_seed(20261018 + rank)
    jg = (torch.arange(d.PY, device=dev, dtype=torch.float64) - d.OLy + d.myPy * NY + 0.5) / d.Ny
    ig = (torch.arange(d.PX, device=dev, dtype=torch.float64) - d.OLx + d.myPx * NX + 0.5) / d.Nx
    YY, XX = torch.meshgrid(jg, ig, indexing="ij")
    two_pi = 2 * np.pi
    psi_u = 0.1 * torch.sin(two_pi * XX) * torch.cos(2 * two_pi * YY) + 0.03 * torch.cos(3 * two_pi * XX) * torch.sin(two_pi * YY)
    psi_v = 0.1 * torch.cos(two_pi * XX) * torch.sin(2 * two_pi * YY) - 0.03 * torch.sin(3 * two_pi * XX) * torch.cos(two_pi * YY)
    prof = torch.linspace(1.0, 0.2, NR, device=dev, dtype=torch.float64)[:, None, None]
    tref = torch.linspace(20.0, 2.0, NR, device=dev, dtype=torch.float64)[:, None, None]
    for name, base, amp in (("uVel", psi_u[None] * prof, 1e-3), ("vVel", psi_v[None] * prof, 1e-3),
                            ("theta", tref.expand(NR, d.PY, d.PX), 0.1)):
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
    rt.set_field("tRef", vert(np.linspace(20.0, 2.0, NR)))
    rt.set_field("sRef", vert(np.zeros(NR)))
    rt.set_field("rF", vert(g.a["rF"]))
    rt.set_field("rC", vert(g.a["rC"]))
    tau = (-0.1 * torch.cos(two_pi * YY) * (1.0 / 1000.0)).contiguous()
    sfU_host = tau.cpu().pin_memory()
    sfV_host = torch.zeros_like(sfU_host).pin_memory()
    eta_host = torch.empty_like(sfU_host).pin_memory()
    rt.set_field("surfForcU", sfU_host)
    rt.set_field("surfForcV", sfV_host)
    halo = distributed.exchange if world > 1 else rt.exch
    for n in ("uVel", "vVel", "theta", "etaN"):
        halo(n)
    t_setup = time.time() - t_setup

    def barrier():
        rt.sync()
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        rt.sync()

    step = distributed.forward_step if world > 1 else rt.forward_step
    it = 0
    for _ in range(args.warmup):
        step(it)
        it += 1
    # ---- timed region 1: resident state ---------------------------------------------------------
    import ctypes as C
    from mitgcm_b200 import _lib
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
    # ---- timed region 2: end to end through the C ABI with host buffers --------------------------
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        rt.set_field("surfForcU", sfU_host)           # H2D from pinned memory
        rt.set_field("surfForcV", sfV_host)
        r = step(it)
        it += 1
        rt.get_field("etaN", eta_host)                # D2H of the step's result
    rt.sync()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([dev_ms, e2e_s], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, e2e_s = float(t[0]), float(t[1])
    finite = bool(np.isfinite(eta_host.numpy()).all())
    if rank != 0:
        rt.finalize()
        return
    K = args.steps
    ms_per_step = dev_ms / K
    value = K / (dev_ms / 1e3)
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
    dom = max(cand, key=lambda k: cand[k][1])
    ach = cand[dom][0] / (cand[dom][1] * 1e-3) / 1e9
    traffic = None
    try:        # DRAM bytes per launch of the dominant kernel from the committed ncu capture of this workload
        tj = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
        if tj.get("workload") == f"{NX}x{NY}x{NR}" and dom in tj:
            traffic = tj[dom]["dram_bytes_per_launch"]
    except Exception:
        pass
    out = {
        "metric": "timesteps/s at 2048x2048x50 (per-GPU block of the weak-scaled channel; aggregate = ranks x steps/s)",
        "value": value * world, "unit": "block-timesteps/s", "n_gpus": world, "steps": K, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"synthetic doubly-periodic channel {NX}x{NY}x{NR} per GPU, FP64, flat bottom, "
                               f"c2 advection + harmonic viscosity, cg2dTargetResidual 1e-7"
                               + (", MOM_VECINV dynamics" if args.momentum == "vecinv" else "")
                               + (f", tempAdvScheme {args.temp_adv_scheme}" if args.temp_adv_scheme != 2 else ""),
                   "process_grid": f"{nPx}x{nPy}", "l2_policy": "working set 30 GB per GPU >> 126 MB L2, no flush needed",
                   "cells_per_gpu": cells},
        "cg2d": {"iters_per_step": tot_iters / K, "iters_per_s": tot_iters / max(phase[3] * 1e-3, 1e-12),
                 "us_per_iter": phase[3] * 1e3 / max(tot_iters, 1),
                 "hbm_gbs": CG2D_BYTES_PER_POINT_ITER * NX * NY * tot_iters / max(phase[3] * 1e-3, 1e-12) / 1e9},
        "phase_share": shares, "phase_ms_per_step": {n: float(phase[i] / K) for i, n in enumerate(names)},
        "kernel_hbm_gbs": {k: float(v[0] / (v[1] * 1e-3) / 1e9) for k, v in cand.items()},
        "roofline": {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                     "peak_source": peak_kind, "traffic": traffic,
                     "algorithmic_bytes_per_launch": cand[dom][0]},
        "e2e": {"value": world * K / e2e_s, "unit": "block-timesteps/s", "h2d_bytes_per_step": int(2 * sfU_host.numel() * 8),
                "d2h_bytes_per_step": int(eta_host.numel() * 8 + 24)},
        "gpu_launches": launches, "clocks": clocks, "setup_s": t_setup, "finite": finite,
    }
    if world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(args, quick=False, steps=4)   # 512x512x50: 10-20 s of host work
    print(json.dumps(out))
    rt.finalize()


# ------------------------------------------------------------------------------------------------
def cpu_baseline(args, quick=False, steps=None, warmup=1):
    """Oracle step on a bounded sample: a (nx/8 x ny/8) ... sized sub-domain with the full Nr, tiled
    one tile per host core, scaled to the full grid by cell count."""
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
    sx, sy = (256, 256) if quick else (512, 512)
    sNx, sNy = max(8, sx // nSx), max(8, sy // nSy)
    P = params(args.nr)
    g, P2, s = make_channel(sNx, sNy, args.nr, nSx=nSx, nSy=nSy, **{k: P[k] for k in P})
    co = ChannelOracle(g, P2, s, threads=nt)
    for _ in range(warmup):
        co.step()
    n_min = steps or (2 if quick else 3)
    t0 = time.perf_counter()
    its = []
    # at least n_min steps, and (full-size sample) on until ~10 s of host work, bounded at 24 steps
    while len(its) < n_min or (not quick and steps is not None and time.perf_counter() - t0 < 10.0 and len(its) < 24):
        its.append(co.step()["numIters"])
    n = len(its)
    dt = time.perf_counter() - t0
    cells = sNx * nSx * sNy * nSy * args.nr
    scale = cells / float(args.nx * args.ny * args.nr)
    return {"value": n / dt * scale, "unit": "block-timesteps/s", "cores": nt, "kind": "port",
            "sample": f"oracle (C restatement of the reference loops, gcc -O2, {nt} tiles on {nt} threads) stepping "
                      f"{sNx * nSx}x{sNy * nSy}x{args.nr} for {n} steps ({dt:.1f} s, {np.mean(its):.0f} CG iters/step), "
                      f"scaled to {args.nx}x{args.ny}x{args.nr} by cell count ({scale:.5f})",
            "sample_steps_per_s": n / dt}


def run_reference(args, rank, world):
    if rank != 0:
        return
    t0 = time.perf_counter()
    cb = cpu_baseline(args, quick=False, steps=None if args.steps < 3 else args.steps, warmup=max(1, min(args.warmup, 2)))
    out = {"impl": "reference",
           "metric": "timesteps/s at 2048x2048x50 (per-GPU block of the weak-scaled channel; aggregate = ranks x steps/s)",
           "value": cb["value"], "unit": "block-timesteps/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": 1e3 / cb["value"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f64", "data": "synthetic",
           "config": {"workload": f"synthetic doubly-periodic channel {args.nx}x{args.ny}x{args.nr} per GPU (CPU oracle on a bounded sample)"},
           "cpu_baseline": cb,
           "e2e": {"value": cb["value"], "unit": "block-timesteps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "wall_s": time.perf_counter() - t0}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--nx", type=int, default=2048)
    ap.add_argument("--ny", type=int, default=2048)
    ap.add_argument("--nr", type=int, default=50)
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
