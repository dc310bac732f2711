"""Multi-GPU driver: one process per GPU (torchrun), tiles sharded over an nPx x nPy process grid.

* CG2D: the persistent kernel talks to its peers directly (CUDA IPC mappings over NVLink): edge
  pushes into the neighbours' halos and mailbox all-reduces, see csrc/cg2d.cu.  torch.distributed
  is used once, to all-gather the 64-byte IPC handles.
* Per-step halos (EXCH_XY_RL / EXCH_XYZ_RL, do_fields_blocking_exchanges.F:54-66): device pack ->
  NCCL send/recv (batch_isend_irecv) -> device unpack; X phase first, then Y phase over the full
  width so corners propagate (exch1_rx.template:172-200).  A periodic direction with a single
  rank is done locally.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.distributed as dist

from . import _lib, runtime as rt
from .grid import Dims
from .parallel import neighbours

_S = {}


def setup(d: Dims):
    """Call after rt.init(d): maps the peers' CG2D workspaces and prepares the exchange buffers."""
    L = _lib.lib()
    world, rank = dist.get_world_size(), dist.get_rank()
    assert world == d.nPx * d.nPy and rank == d.myPx + d.nPx * d.myPy
    assert d.nSx == 1 and d.nSy == 1, "multi-rank runs use one tile per rank"
    h = (C.c_ubyte * 64)()
    ierr = C.c_int(0)
    L.mitgcm_b200_comm_handle_(h, C.byref(ierr))
    rt._check(ierr)
    dev = torch.device("cuda", torch.cuda.current_device())
    mine = torch.tensor(list(h), dtype=torch.uint8, device=dev)
    allh = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(allh, mine)
    flat = torch.cat(allh).cpu().numpy().tobytes()
    buf = (C.c_ubyte * (64 * world)).from_buffer_copy(flat)
    L.mitgcm_b200_comm_connect_(C.byref(C.c_int(world)), C.byref(C.c_int(rank)), buf, C.byref(ierr))
    rt._check(ierr)
    dist.barrier()
    _S.update(d=d, nbr=neighbours(rank, d.nPx, d.nPy), dev=dev, bufs={})


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
    """EXCH_XY(Z)_RL of one or several mirrors: the strips of all fields travel in ONE message per
    neighbour and phase."""
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
    """The blocking exchange at the end of the step includes salt when saltStepping is on."""
    global _salt_stepping
    _salt_stepping = bool(on)


def forward_step(myIter: int):
    """FORWARD_STEP across ranks: part 0 (thermodynamics, dynamics, CG2D with in-kernel peer
    communication), halo of cg2d_x, part 1 (eta, correction step, continuity), blocking exchanges."""
    L = _lib.lib()
    f, l, n, ierr = C.c_double(), C.c_double(), C.c_int(), C.c_int(0)
    L.mitgcm_b200_step_part_(C.byref(C.c_int(0)), C.byref(C.c_int(myIter)), C.byref(f), C.byref(n), C.byref(l), C.byref(ierr))
    rt._check(ierr)
    exchange("cg2d_x")
    L.mitgcm_b200_step_part_(C.byref(C.c_int(1)), C.byref(C.c_int(myIter)), C.byref(f), C.byref(n), C.byref(l), C.byref(ierr))
    rt._check(ierr)
    exchange(*(("uVel", "vVel", "wVel", "theta") + (("salt",) if _salt_stepping else ())))
    return dict(firstResidual=f.value, numIters=n.value, lastResidual=l.value)
