"""One process per GPU: process grid, neighbour ranks and ordered global sums
(eesupp/src/ini_procs.F:145-260, eesupp/src/global_sum_tile.F:161-191) over torch.distributed."""
from __future__ import annotations

import torch
import torch.distributed as dist


def process_grid(world: int) -> tuple[int, int]:
    """nPx x nPy for 1, 2, 4, 8 ranks: split y first (y-edge strips are contiguous rows).
    MITGCM_B200_PGRID="2x1" overrides (tests of the x-split on two GPUs)."""
    import os
    o = os.environ.get("MITGCM_B200_PGRID")
    if o:
        px, py = (int(v) for v in o.lower().split("x"))
        assert px * py == world, "MITGCM_B200_PGRID does not match the world size"
        return px, py
    return {1: (1, 1), 2: (1, 2), 4: (2, 2), 8: (2, 4)}.get(world, (1, world))


def neighbours(rank: int, nPx: int, nPy: int) -> dict:
    """Periodic Cartesian neighbours, rank = px + nPx*py (MPI_CART_CREATE, ini_procs.F:145)."""
    px, py = rank % nPx, rank // nPx
    r = lambda x, y: (x % nPx) + nPx * (y % nPy)
    return dict(W=r(px - 1, py), E=r(px + 1, py), S=r(px, py - 1), N=r(px, py + 1), px=px, py=py)


def ordered_global_sum(partials: torch.Tensor) -> torch.Tensor:
    """GLOBAL_SUM_TILE_RL with GLOBAL_SUM_ORDER_TILES: gather every rank's partial(s) and add
    them in rank order on every rank, so all ranks hold bit-identical totals."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return partials.clone()
    world = dist.get_world_size()
    buf = [torch.empty_like(partials) for _ in range(world)]
    dist.all_gather(buf, partials)
    tot = torch.zeros_like(partials)
    for b in buf:
        tot = tot + b
    return tot
