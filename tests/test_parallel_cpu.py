"""CPU checks of the multi-rank host logic (process grid, neighbours, exchange strips) --
the N > 1 GPU path itself is checked by scripts/dist_check.py on real GPUs."""
import numpy as np

from mitgcm_b200.parallel import neighbours, process_grid


def test_process_grid_matches_bench_layouts():
    assert [process_grid(n) for n in (1, 2, 4, 8)] == [(1, 1), (1, 2), (2, 2), (2, 4)]


def test_neighbours_are_periodic_and_symmetric():
    for world in (2, 4, 8):
        nPx, nPy = process_grid(world)
        for r in range(world):
            n = neighbours(r, nPx, nPy)
            assert neighbours(n["E"], nPx, nPy)["W"] == r
            assert neighbours(n["N"], nPx, nPy)["S"] == r
            assert n["px"] + nPx * n["py"] == r


def test_strip_exchange_emulation_matches_single_process_exchange():
    """The X-then-Y strip protocol of mitgcm_b200/distributed.py, emulated with numpy on a 2x2
    process grid, reproduces EXCH_XYZ_RL of the same global domain tiled 2x2 in one process."""
    from mitgcm_b200.grid import Dims, exch_xyz
    sNx, sNy, OL, nz = 7, 5, 2, 3
    dG = Dims(sNx, sNy, OL, OL, nSx=2, nSy=2, Nr=nz)
    rng = np.random.default_rng(3)
    A = rng.standard_normal(dG.shape3)
    ref = exch_xyz(dG, A.copy())
    tiles = {(px, py): A[py, px].copy() for px in range(2) for py in range(2)}
    # X phase: OLx columns x sNy interior rows
    new = {k: v.copy() for k, v in tiles.items()}
    for (px, py), t in tiles.items():
        w, e = tiles[((px - 1) % 2, py)], tiles[((px + 1) % 2, py)]
        new[(px, py)][:, OL:OL + sNy, :OL] = w[:, OL:OL + sNy, sNx:sNx + OL]
        new[(px, py)][:, OL:OL + sNy, OL + sNx:] = e[:, OL:OL + sNy, OL:2 * OL]
    tiles = new
    new = {k: v.copy() for k, v in tiles.items()}
    for (px, py), t in tiles.items():       # Y phase: OLy rows x full width (corners propagate)
        s, n = tiles[(px, (py - 1) % 2)], tiles[(px, (py + 1) % 2)]
        new[(px, py)][:, :OL, :] = s[:, sNy:sNy + OL, :]
        new[(px, py)][:, OL + sNy:, :] = n[:, OL:2 * OL, :]
    for (px, py), t in new.items():
        assert np.array_equal(t, ref[py, px])
