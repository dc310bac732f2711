"""pkg/exch2 tile graph spread over ranks, host side (no GPU): the lists every rank is given
(mitgcm_b200_exch2_dist_lists_ = what mitgcm_b200_set_exch2_topology_ uploads on a multi-rank run) drive a
many-rank exchange in numpy -- each "rank" owns its (nTiles / nRanks, PY, PX) arrays, reads tagged sources out of the
owner's arrays exactly as gather_kernel / gather_uv_kernel / CG2D's pdst() do -- and the result must be bit-identical to
the literal two-pass buffered algorithm on the whole graph (oracle/exch2_oracle.py), i.e. to what the MPI messages of
exch2_send_rx{1,2}.template / exch2_recv_rx{1,2}.template deliver."""
import numpy as np
import pytest

from mitgcm_b200.exch2 import cubed_sphere_topology, dist_lists, tile_proc
from oracle import exch2_oracle as eo

MASK = 0x0FFFFFFF
CASES = [(32, 32, 16, 4), (8, 4, 4, 2), (6, 6, 6, 3), (32, 16, 8, 2)]


def split(a, nRanks):
    n = a.shape[0] // nRanks
    return [a[r * n:(r + 1) * n].copy() for r in range(nRanks)]


@pytest.mark.parametrize("nRanks", [1, 2, 3, 4, 6])
@pytest.mark.parametrize("nf,sx,sy,OL", CASES)
def test_scalar_exchange_across_ranks_is_the_two_pass_exchange(nf, sx, sy, OL, nRanks):
    T = cubed_sphere_topology(nf, sx, sy)
    if T.nTiles % nRanks:
        pytest.skip("tiles do not divide over the ranks")
    nz = 2
    rng = np.random.default_rng(3)
    a = rng.standard_normal((T.nTiles, nz, sy + 2 * OL, sx + 2 * OL))
    parts = split(a, nRanks)          # every rank's own arrays, (nLocal, nz, PY, PX)
    slab = (sy + 2 * OL) * (sx + 2 * OL)
    flat = [p.transpose(1, 0, 2, 3).reshape(nz, -1) for p in parts]      # level-major views: index = tile*slab + cell
    new = [f.copy() for f in flat]
    for r in range(nRanks):
        sc, _, _ = dist_lists(T, OL, nRanks, r)
        dst, src, own = sc[:, 0].astype(np.int64), (sc[:, 1] & MASK).astype(np.int64), (sc[:, 1] >> 28) & 7
        assert dst.max() < parts[r].shape[0] * slab and (nRanks > 1 or not own.any())
        for o in range(nRanks):
            m = own == o
            new[r][:, dst[m]] = flat[o][:, src[m]]          # reads see the pre-exchange values: sources are interior cells
    eo.exch2_3d(T, a, OL)
    ref = split(a, nRanks)
    for r in range(nRanks):
        got = new[r].reshape(nz, -1, sy + 2 * OL, sx + 2 * OL).transpose(1, 0, 2, 3)
        assert np.array_equal(got, ref[r]), r


@pytest.mark.parametrize("withSigns", [True, False])
@pytest.mark.parametrize("nRanks", [2, 3, 6])
@pytest.mark.parametrize("nf,sx,sy,OL", CASES[:3])
def test_vector_exchange_across_ranks_is_exch2_uv_3d(nf, sx, sy, OL, nRanks, withSigns):
    T = cubed_sphere_topology(nf, sx, sy)
    if T.nTiles % nRanks:
        pytest.skip("tiles do not divide over the ranks")
    rng = np.random.default_rng(4)
    shape = (T.nTiles, 1, sy + 2 * OL, sx + 2 * OL)
    u, v = rng.standard_normal(shape), rng.standard_normal(shape)
    fl = [[p.reshape(-1) for p in split(x, nRanks)] for x in (u, v)]      # fl[array][rank]
    new = [[f.copy() for f in fa] for fa in fl]
    for r in range(nRanks):
        _, _, uv = dist_lists(T, OL, nRanks, r, withSigns)
        for da, dst, sa, src in uv:
            val = fl[sa >> 1][(src >> 28) & 7][src & MASK]
            new[da][r][dst] = -val if sa & 1 else val
    eo.exch2_uv_3d(T, u, v, OL, withSigns)
    for arr, ref in enumerate((u, v)):
        for r, part in enumerate(split(ref, nRanks)):
            assert np.array_equal(new[arr][r], part.reshape(-1)), (arr, r)


@pytest.mark.parametrize("nRanks", [1, 2, 4, 6])
@pytest.mark.parametrize("nf,sx,sy,OL", CASES)
def test_push_table_across_ranks_is_exch2_s3d(nf, sx, sy, OL, nRanks):
    """CG2D's width-1 exchange: every rank stores the edge values of its tiles into the halo cell the table names, on the
    rank it names.  Arrays in the solver's layout (full overlap OL); compared on the (0:sNx+1, 0:sNy+1) frame."""
    T = cubed_sphere_topology(nf, sx, sy)
    if T.nTiles % nRanks:
        pytest.skip("tiles do not divide over the ranks")
    PX, PY = sx + 2 * OL, sy + 2 * OL
    rng = np.random.default_rng(5)
    a = rng.standard_normal((T.nTiles, PY, PX))
    parts = [p.reshape(-1) for p in split(a, nRanks)]
    nL = T.nTiles // nRanks
    idx = lambda i, j, t: (i + OL - 1) + PX * (j + OL - 1) + PX * PY * t      # Fortran (i,j) of local tile t
    for r in range(nRanks):
        _, push, _ = dist_lists(T, OL, nRanks, r)
        assert push.shape == (nL, 2 * sy + 2 * sx)
        for t in range(nL):
            for j in range(1, sy + 1):
                for slot, i in ((j - 1, 1), (sy + j - 1, sx)):
                    e = int(push[t, slot])
                    parts[(e >> 28) & 7][e & MASK] = parts[r][idx(i, j, t)]
            for i in range(1, sx + 1):
                for slot, j in ((2 * sy + i - 1, 1), (2 * sy + sx + i - 1, sy)):
                    e = int(push[t, slot])
                    parts[(e >> 28) & 7][e & MASK] = parts[r][idx(i, j, t)]
    ring = a[:, None, OL - 1:OL + sy + 1, OL - 1:OL + sx + 1].copy()
    eo.exch2_s3d(T, ring)
    got = np.concatenate([p.reshape(nL, PY, PX) for p in parts])[:, OL - 1:OL + sy + 1, OL - 1:OL + sx + 1]
    edge = np.ones((sy + 2, sx + 2), bool)
    edge[0, 0] = edge[0, -1] = edge[-1, 0] = edge[-1, -1] = False      # corners are not part of the width-1 exchange
    assert np.array_equal(got[:, edge], ring[:, 0][:, edge])


def test_tile_proc_and_rank_tiles():
    from mitgcm_b200.grid import Dims, Grid
    from mitgcm_b200.model import rank_tiles
    assert tile_proc(12, 3).tolist() == [1] * 4 + [2] * 4 + [3] * 4
    d = Dims(sNx=4, sNy=3, OLx=1, OLy=1, nSx=6, nSy=1, Nr=2)
    g = Grid(d, dict(rA=np.arange(np.prod(d.shape2), dtype=float).reshape(d.shape2), drF=np.array([1.0, 2.0])))
    st = dict(theta=np.arange(np.prod(d.shape3), dtype=float).reshape(d.shape3), cg2dNorm=3.0)
    gl, (sl,) = rank_tiles(g, [st], 1, 3)
    assert gl.d.nSx == 2 and gl.d.nPx == 3 and gl.d.myPx == 1 and gl.a["rA"].shape == gl.d.shape2
    assert np.array_equal(gl.a["rA"], g.a["rA"][:, 2:4]) and np.array_equal(sl["theta"], st["theta"][:, 2:4])
    assert sl["cg2dNorm"] == 3.0 and np.array_equal(gl.a["drF"], g.a["drF"])


def test_gloo_world2_tile_graph_exchange_and_cg2d_sums():
    """Two real processes (gloo): each holds half of the cube's tiles, fetches the cells its gather list names from
    the owner (all_gather stands in for the NVLink read of the peer arena), and must end with exactly what the literal
    exch2 exchange gives its tiles; the CG2D dot product over the distributed tiles is added in rank order."""
    import torch.multiprocessing as mp
    mp.spawn(_gloo_tile_graph_worker, args=(2,), nprocs=2, join=True)


def _gloo_tile_graph_worker(rank, world):
    import os
    import torch
    import torch.distributed as dist
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29519")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mitgcm_b200.parallel import ordered_global_sum
    nf, sx, sy, OL, nz = 8, 4, 4, 2, 2
    T = cubed_sphere_topology(nf, sx, sy)
    n = T.nTiles // world
    rng = np.random.default_rng(11)                      # the same global field in both processes
    a = rng.standard_normal((T.nTiles, nz, sy + 2 * OL, sx + 2 * OL))
    mine = a[rank * n:(rank + 1) * n].copy()
    t = torch.from_numpy(mine.transpose(1, 0, 2, 3).reshape(nz, -1).copy())
    parts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(parts, t)                            # what the peers hold (pre-exchange values)
    sc, _, _ = dist_lists(T, OL, world, rank)
    new = t.clone()
    dst, src, own = sc[:, 0].astype(np.int64), (sc[:, 1] & MASK).astype(np.int64), (sc[:, 1] >> 28) & 7
    for o in range(world):
        m = own == o
        new[:, dst[m]] = parts[o][:, src[m]]
    eo.exch2_3d(T, a, OL)
    got = new.numpy().reshape(nz, n, sy + 2 * OL, sx + 2 * OL).transpose(1, 0, 2, 3)
    assert np.array_equal(got, a[rank * n:(rank + 1) * n])
    # GLOBAL_SUM_TILE_RL over the distributed tiles: tile sums in tile order inside a rank, ranks in rank order
    tile_sums = [float(np.sum(got[l, :, OL:OL + sy, OL:OL + sx] ** 2)) for l in range(n)]
    part = 0.0
    for v in tile_sums:
        part = part + v
    tot = ordered_global_sum(torch.tensor([part], dtype=torch.float64))
    ref = 0.0
    for r in range(world):
        p = 0.0
        for l in range(n):
            p = p + float(np.sum(a[r * n + l, :, OL:OL + sy, OL:OL + sx] ** 2))
        ref = ref + p
    assert tot.item() == ref
    dist.destroy_process_group()
