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


def _push_emulation(nPx, nPy, sNx, sNy, OL, nz, seed=5):
    """csrc/halo.cu::halo_push_kernel in numpy: every rank stores its edge strips and corner blocks into the 8
    neighbours' halos (index arithmetic copied from the kernel), one phase, no X-before-Y ordering."""
    from mitgcm_b200.grid import Dims, exch_xyz
    dG = Dims(sNx, sNy, OL, OL, nSx=nPx, nSy=nPy, Nr=nz)
    rng = np.random.default_rng(seed)
    A = rng.standard_normal(dG.shape3)
    ref = exch_xyz(dG, A.copy())
    tiles = {(px, py): A[py, px].copy() for px in range(nPx) for py in range(nPy)}
    PX = sNx + 2 * OL
    dxs = [-1, 1, 0, 0, -1, 1, -1, 1]          # 0 W, 1 E, 2 S, 3 N, 4 SW, 5 SE, 6 NW, 7 NE
    dys = [0, 0, -1, 1, -1, -1, 1, 1]
    n0, n2, n4 = OL * sNy, sNx * OL, OL * OL
    perLevel = 2 * n0 + 2 * n2 + 4 * n4
    for (px, py), src in tiles.items():
        flat = src.reshape(nz, -1)
        for c0 in range(perLevel):
            c = c0
            if c < 2 * n0:
                d = int(c >= n0); c -= d * n0; w = OL
            elif c - 2 * n0 < 2 * n2:
                c -= 2 * n0; d = 2 + int(c >= n2); c -= (d - 2) * n2; w = sNx
            else:
                c -= 2 * n0 + 2 * n2; d = 4 + c // n4; c -= (d - 4) * n4; w = OL
            bI, bJ = c % w, c // w
            dx, dy = dxs[d], dys[d]
            i = (sNx - OL + 1 if dx > 0 else 1) + bI
            j = (sNy - OL + 1 if dy > 0 else 1) + bJ
            iD, jD = i - dx * sNx, j - dy * sNy
            s = (i + OL - 1) + PX * (j + OL - 1)
            t = (iD + OL - 1) + PX * (jD + OL - 1)
            dst = tiles[((px + dx) % nPx, (py + dy) % nPy)].reshape(nz, -1)
            dst[:, t] = flat[:, s]
    for (px, py), t in tiles.items():
        assert np.array_equal(t, ref[py, px]), (nPx, nPy, px, py)


def test_single_phase_peer_push_matches_exchange():
    for nPx, nPy in ((1, 2), (2, 2), (2, 4), (1, 1)):
        _push_emulation(nPx, nPy, 7, 5, 2, 3)
    _push_emulation(2, 2, 6, 4, 3, 2)
