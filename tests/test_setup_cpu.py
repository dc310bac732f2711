"""CPU checks of the host-side set-up code against the oracle (no GPU needed)."""
import numpy as np

from helpers import make_grid
from mitgcm_b200.grid import global_area
from mitgcm_b200.model import ini_cg2d, make_channel
from oracle.pyoracle import Oracle


def test_numpy_ini_cg2d_is_bit_identical_to_the_oracle():
    for kw in (dict(sNx=31, sNy=17, OL=3, nSx=2, nSy=2, Nr=4), dict(sNx=20, sNy=12, OL=2, Nr=1)):
        g = make_grid(**kw, seed=2)
        P = dict(deltaTMom=900.0, deltaTFreeSurf=1200.0, cg2dTargetResidual=1e-8, globalArea=global_area(g))
        ref = Oracle(g, P).ini_cg2d()
        got = ini_cg2d(g, P)
        for n in "aW2d aS2d aC2d pW pS pC".split():
            assert np.array_equal(got[n], ref[n]), n
        assert got["cg2dNorm"] == ref["cg2dNorm"] and got["cg2dTolerance_sq"] == ref["cg2dTolerance_sq"]


def test_flat_operator_shortcut_matches_full_arrays():
    g, P, _ = make_channel(16, 12, 5)
    a, b = ini_cg2d(g, P), ini_cg2d(g, P, hfac_flat=1.0)
    for n in "aW2d aS2d aC2d pW pS pC".split():
        assert np.array_equal(a[n], b[n])


def test_gloo_world2_tile_sum_is_rank_ordered():
    """Host logic of the N > 1 path (no GPU): per-rank partial sums gathered and added in rank
    order give the same value on every rank (GLOBAL_SUM_TILE_RL semantics, global_sum_tile.F:161-191)."""
    import torch.multiprocessing as mp
    mp.spawn(_gloo_worker, args=(2,), nprocs=2, join=True)


def _gloo_worker(rank, world):
    import os
    import torch
    import torch.distributed as dist
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29517")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mitgcm_b200.parallel import ordered_global_sum
    part = torch.tensor([0.1 * (rank + 1), 1e-17 * (rank + 3)], dtype=torch.float64)
    tot = ordered_global_sum(part)
    ref = torch.tensor([0.1 + 0.2, 1e-17 * 3 + 1e-17 * 4], dtype=torch.float64)
    assert torch.equal(tot, ref)
    dist.destroy_process_group()
