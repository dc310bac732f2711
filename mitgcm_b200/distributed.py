"""Multi-GPU driver: one process per GPU (torchrun), tiles sharded over an nPx x nPy process grid.

Default transport = peer memory (csrc/halo.cu, csrc/cg2d.cu): every rank exports ONE CUDA IPC handle (its peer
arena) and maps its peers'; torch.distributed is used once, to all-gather the 72-byte handles.  After that
* CG2D talks to its peers from inside the persistent kernel (edge pushes into the neighbours' halos, mailbox
  all-reduces),
* the per-step halos (EXCH_XY_RL / EXCH_XYZ_RL, do_fields_blocking_exchanges.F:54-66) are stores of edge strips
  and corner blocks straight into the 8 neighbours' halo cells, ordered by peer-memory flags on the stream, and
* mitgcm_b200_forward_step_ runs the whole step across ranks without a host synchronisation besides the solver's
  own read-back of its iteration count (theta's halo travels on a side stream under DYNAMICS).

transport="nccl" keeps the round-1 path as the baseline to compare against: device pack -> NCCL send/recv
(batch_isend_irecv) -> device unpack; X phase first, then Y phase over the full width so corners propagate
(exch1_rx.template:172-200); host-synchronous.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch
import torch.distributed as dist

from . import _lib, runtime as rt
from .grid import Dims
from .parallel import neighbours

_S = {}


def setup(d: Dims, transport: str | None = None):
    """Call after rt.init(d): maps the peers' arenas (CG2D workspace + exchanged fields)."""
    L = _lib.lib()
    world, rank = dist.get_world_size(), dist.get_rank()
    assert world == d.nPx * d.nPy and rank == d.myPx + d.nPx * d.myPy
    # one tile per rank on the periodic process grid, or several tiles of an exch2 tile graph (exch2.set_topology with tileProc)
    transport = transport or os.environ.get("MITGCM_B200_TRANSPORT", "peer")
    assert transport in ("peer", "nccl")
    assert transport == "peer" or (d.nSx == 1 and d.nSy == 1), "the NCCL strip exchange handles one tile per rank"
    h = (C.c_ubyte * 72)()
    ierr = C.c_int(0)
    L.mitgcm_b200_comm_handle_(h, C.byref(ierr))
    rt._check(ierr)
    dev = torch.device("cuda", torch.cuda.current_device())
    mine = torch.tensor(list(h), dtype=torch.uint8, device=dev)
    allh = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(allh, mine)
    flat = torch.cat(allh).cpu().numpy().tobytes()
    buf = (C.c_ubyte * (72 * world)).from_buffer_copy(flat)
    L.mitgcm_b200_comm_connect_(C.byref(C.c_int(world)), C.byref(C.c_int(rank)), buf, C.byref(ierr))
    rt._check(ierr)
    dist.barrier()          # every rank has zeroed its flags before anybody raises one
    _S.clear()
    _S.update(d=d, nbr=neighbours(rank, d.nPx, d.nPy), dev=dev, bufs={}, transport=transport)


def teardown():
    """Unmap the peers, barrier, so that every rank may then free its arena (rt.finalize)."""
    rt.sync()
    _lib.lib().mitgcm_b200_comm_disconnect_()
    dist.barrier()
    _S.clear()


def _buf(key, n):
    b = _S["bufs"].get(key)
    if b is None or b.numel() < n:
        b = torch.empty(n, dtype=torch.float64, device=_S["dev"])
        _S["bufs"][key] = b
    return b


def _pack(fid, direction, ptr, unpack):
    ierr = C.c_int(0)
    _lib.lib().mitgcm_b200_pack_(C.byref(C.c_int(fid)), C.byref(C.c_int(direction)), C.c_void_p(ptr),
                                 C.byref(C.c_int(unpack)), C.byref(ierr))
    rt._check(ierr)


def exchange(*names: str):
    """EXCH_XY(Z)_RL of one or several mirrors across ranks, all fields in one exchange."""
    if _S["transport"] == "peer":
        ids = (C.c_int * len(names))(*[rt.field_id(n) for n in names])
        ierr = C.c_int(0)
        _lib.lib().mitgcm_b200_halo_exchange_(C.byref(C.c_int(len(names))), ids, C.byref(ierr))
        rt._check(ierr)
        return
    _exchange_nccl(*names)


def _exchange_nccl(*names: str):
    d, nbr = _S["d"], _S["nbr"]
    L = _lib.lib()
    fids = [rt.field_id(n) for n in names]
    nzs = [1 if f < 100 else (d.Nr if f < 200 else d.Nr + 1) for f in fids]
    ierr = C.c_int(0)
    for ydir, (lo, hi, nP, w, h) in enumerate(((("W", "E", d.nPx, d.OLx, d.sNy)), ("S", "N", d.nPy, d.PX, d.OLy))):
        if nP == 1:
            for fid in fids:
                L.mitgcm_b200_exch_dir_(C.byref(C.c_int(fid)), C.byref(C.c_int(ydir)), C.byref(ierr))
                rt._check(ierr)
            continue
        sizes = [w * h * nz for nz in nzs]
        n = sum(sizes)
        s_lo, s_hi = _buf(("s", ydir, 0), n), _buf(("s", ydir, 1), n)
        r_lo, r_hi = _buf(("r", ydir, 0), n), _buf(("r", ydir, 1), n)
        off = 0
        for fid, sz in zip(fids, sizes):
            _pack(fid, 2 * ydir, s_lo.data_ptr() + 8 * off, 0)       # strip at my low edge -> low neighbour's high halo
            _pack(fid, 2 * ydir + 1, s_hi.data_ptr() + 8 * off, 0)
            off += sz
        rt.sync()
        ops = [dist.P2POp(dist.isend, s_lo[:n], nbr[lo]), dist.P2POp(dist.isend, s_hi[:n], nbr[hi]),
               dist.P2POp(dist.irecv, r_hi[:n], nbr[hi]), dist.P2POp(dist.irecv, r_lo[:n], nbr[lo])]
        for req in dist.batch_isend_irecv(ops):
            req.wait()
        torch.cuda.synchronize()
        off = 0
        for fid, sz in zip(fids, sizes):
            _pack(fid, 2 * ydir, r_lo.data_ptr() + 8 * off, 1)
            _pack(fid, 2 * ydir + 1, r_hi.data_ptr() + 8 * off, 1)
            off += sz
    rt.sync()


_salt_stepping = False


def set_salt_stepping(on: bool):
    """The blocking exchange at the end of the step includes salt when saltStepping is on (NCCL transport;
    the peer transport reads MI_SALTSTEPPING itself)."""
    global _salt_stepping
    _salt_stepping = bool(on)


def forward_step(myIter: int):
    """FORWARD_STEP across ranks.  peer transport: one library call.  nccl transport: part 0 (thermodynamics,
    dynamics, CG2D with in-kernel peer communication), halo of cg2d_x, part 1 (eta, correction step, continuity),
    blocking exchanges."""
    if _S["transport"] == "peer":
        return rt.forward_step(myIter)
    L = _lib.lib()
    f, l, n, ierr = C.c_double(), C.c_double(), C.c_int(), C.c_int(0)
    L.mitgcm_b200_step_part_(C.byref(C.c_int(0)), C.byref(C.c_int(myIter)), C.byref(f), C.byref(n), C.byref(l), C.byref(ierr))
    rt._check(ierr)
    _exchange_nccl("cg2d_x")
    L.mitgcm_b200_step_part_(C.byref(C.c_int(1)), C.byref(C.c_int(myIter)), C.byref(f), C.byref(n), C.byref(l), C.byref(ierr))
    rt._check(ierr)
    _exchange_nccl(*(("uVel", "vVel", "wVel", "theta") + (("salt",) if _salt_stepping else ())))
    return dict(firstResidual=f.value, numIters=n.value, lastResidual=l.value)


def selfcheck(NXg=64, NYg=48, NR=4, nsteps=10, wide=False, buoyancy=True, land_frac=0.15, transport=None,
              verbose=False, tol=1e-9):
    """N-rank vs 1-rank parity of the resident step: the same global problem is stepped (a) on all ranks and (b)
    on rank 0 alone as one process holding nPx x nPy tiles; fields must agree (CG2D sums are rank-ordered, so only
    the summation shape differs) and the CG2D iteration counts must agree to +-1.  Needs an initialised NCCL
    process group with one GPU per rank; the library context is re-initialised (call it before the real set-up).
    Returns dict(ok, max_rel_err, iters, ref_iters, ...) on every rank.
    wide: MOM_VECINV + SALT_INTEGRATE + both tracers through GAD_ADVECTION (DST3) at overlap 3."""
    from .model import LIB_PARAMS, Model, ini_cg2d, make_channel
    from .grid import exch_xyz
    from .parallel import process_grid
    rank, world = dist.get_rank(), dist.get_world_size()
    local = torch.cuda.current_device()
    nPx, nPy = process_grid(world)
    sNx, sNy = NXg // nPx, NYg // nPy
    OL = 3 if wide else 2
    extra = dict(vectorInvariantMomentum=1, saltStepping=1, tempAdvScheme=33, saltAdvScheme=33, diffKhS=5e2,
                 diffKrS=2e-5, sBeta=7.4e-4) if wide else {}
    gG, P, sG = make_channel(sNx, sNy, NR, nSx=nPx, nSy=nPy, OL=OL, land_frac=land_frac,
                             buoyancyLinear=int(buoyancy), **extra)
    if wide:
        rng = np.random.default_rng(7)
        sG["salt"] = exch_xyz(gG.d, (35.0 + np.linspace(-0.5, 0.5, NR)[None, None, :, None, None]
                                     + 0.05 * rng.standard_normal(gG.d.shape3)) * gG.maskC)
    opG = ini_cg2d(gG, P)
    px, py = rank % nPx, rank // nPx
    d = Dims(sNx=sNx, sNy=sNy, OLx=OL, OLy=OL, Nr=NR, nPx=nPx, nPy=nPy, myPx=px, myPy=py)

    def mine(a):            # (nSy,nSx,...) global tiling -> my tile as a 1x1 tiling
        return np.ascontiguousarray(a[py:py + 1, px:px + 1])

    rt.init(d, local)
    for n in rt.GRID_FIELD_NAMES:
        if n in gG.a:
            rt.set_field(n, mine(gG.a[n]))
    for n in ("drF", "drC", "recip_drF", "recip_drC"):
        v = np.zeros(NR + 1)
        v[:len(gG.a[n])] = gG.a[n]
        rt.set_field(n, v)
    rt.set_params(**{k: P[k] for k in LIB_PARAMS if k in P})
    adv = 33 if wide else 2
    rt.set_params(deltaTtracer=P["deltaTtracer"], tempAdvScheme=adv, tempVertAdvScheme=adv, saltAdvScheme=adv,
                  saltVertAdvScheme=adv, nIter0=0)
    setup(d, transport)
    set_salt_stepping(bool(wide))
    rt.set_cg2d_operator({k: (mine(v) if isinstance(v, np.ndarray) else v) for k, v in opG.items()})
    for n in ("uVel", "vVel", "wVel", "theta", "etaN", "surfForcU", "surfForcV"):
        rt.set_field(n, mine(sG[n]))
    rt.fill_field("kappaRU", P["viscAr"])
    rt.fill_field("kappaRV", P["viscAr"])
    rt.fill_field("kappaRT", P["diffKrT"])
    for n in ("tRef", "sRef", "rF", "rC"):      # linear EOS + CALC_PHI_HYD inputs
        src = sG.get(n, gG.a.get(n))
        v = np.zeros(NR + 1)
        v[:len(src)] = src
        rt.set_field(n, v)
    for n in ("gU", "gV", "guNm1", "gvNm1", "gtNm1", "theta2", "cg2d_b", "cg2d_x"):
        rt.fill_field(n, 0.0)
    fields = ("uVel", "vVel", "wVel", "theta", "etaN") + (("salt",) if wide else ())
    if wide:
        rt.set_field("salt", mine(sG["salt"]))
        rt.fill_field("kappaRS", P["diffKrS"])
        rt.fill_field("gsNm1", 0.0)
        rt.fill_field("salt2", 0.0)
    dist.barrier()
    res = [forward_step(it) for it in range(nsteps)]
    out = {n: rt.get_field(n, np.zeros(d.shape2 if n == "etaN" else d.shape3)) for n in fields}
    dist.barrier()
    tr = _S.get("transport")
    teardown()
    rt.finalize()
    dist.barrier()
    # every rank checks its tile against the single-process run of the same global domain (rank 0 computes it)
    if rank == 0:
        m = Model(gG, P, sG, opG, device=local)
        ref_res = [m.step() for _ in range(nsteps)]
        ref = {n: m.get(n) for n in out}
        m.close()
        payload = [ref, ref_res]
    else:
        payload = [None, None]
    dist.broadcast_object_list(payload, src=0)
    ref, ref_res = payload
    worst, ok = 0.0, True
    for n, a in out.items():
        b = ref[n][py:py + 1, px:px + 1]
        err = float(np.abs(a - b).max() / max(np.abs(ref[n]).max(), 1e-300))
        if not np.isfinite(err) or err > tol:
            ok = False
        worst = max(worst, err) if np.isfinite(err) else float("inf")
        if verbose:
            at = np.unravel_index(np.abs(a - b).argmax(), a.shape)[2:]
            print(f"rank {rank} {n}: rel err {err:.2e} at (k,j,i)={tuple(int(x) for x in at)} tile sum {np.abs(a).sum():.6e}",
                  flush=True)
    its, rits = [r["numIters"] for r in res], [r["numIters"] for r in ref_res]
    if any(abs(a - b) > 1 for a, b in zip(its, rits)):
        ok = False
    t = torch.tensor([0.0 if ok else 1.0, worst if np.isfinite(worst) else 1e300], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return dict(ok=bool(t[0].item() == 0.0), max_rel_err=float(t[1].item()), iters=its, ref_iters=rits,
                global_grid=f"{NXg}x{NYg}x{NR}", process_grid=f"{nPx}x{nPy}", steps=nsteps,
                transport=tr, tol=tol)
