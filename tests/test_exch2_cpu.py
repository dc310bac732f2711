"""pkg/exch2 (tile-graph exchange) on the CPU: topology generator,
the exchange compiled into a gather, and the config-4 (global_ocean.cs32x15) known-answer values."""
import json
import os

import numpy as np
import pytest

from helpers import load_cs32
from mitgcm_b200.exch2 import cubed_sphere_topology, halo_gather_map, exchange
from oracle import exch2_oracle as eo

GOLD4 = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "global_ocean.cs32x15.json")))


@pytest.mark.parametrize("nf,sx,sy,OL", [(32, 32, 16, 4), (8, 4, 4, 2), (6, 6, 6, 3), (8, 8, 2, 2), (12, 3, 4, 1)])
def test_compiled_gather_equals_the_two_pass_buffered_exchange(nf, sx, sy, OL):
    """One gather per field == EXCH2_RX1_CUBE(IGNORE_CORNERS) + EXCH2_RX1_CUBE(UPDATE_CORNERS) with
    buffers (exch2_3d_rx.template), bit for bit, including the cube-corner halo cells; same for the
    width-1 EXCH2_S3D_RX."""
    T = cubed_sphere_topology(nf, sx, sy)
    rng = np.random.default_rng(nf + sx)
    a = rng.standard_normal((T.nTiles, 3, sy + 2 * OL, sx + 2 * OL))
    b = a.copy()
    eo.exch2_3d(T, a, OL)
    exchange(T, b, OL)
    assert np.array_equal(a, b)
    a1 = rng.standard_normal((T.nTiles, 1, sy + 2, sx + 2))
    b1 = a1.copy()
    eo.exch2_s3d(T, a1)
    dst, src = halo_gather_map(T, 1, 1, two_pass=False)
    f = b1.reshape(-1)
    f[dst] = f[src]
    assert np.array_equal(a1, b1)
    if sx == sy or (nf // sx) * sx == nf and (nf // sy) * sy == nf and (sx % sy == 0 or sy % sx == 0):
        assert len(dst) == 2 * (sx + sy) * T.nTiles     # tiles that line up across rotated edges: no gaps


def test_topology_tables_are_consistent():
    T = cubed_sphere_topology(32, 32, 16)       # the cs32x15 decomposition: 12 tiles of 32x16
    assert T.nTiles == 12 and list(T.myFace) == [1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6]
    for t in range(T.nTiles):
        for n in range(T.nNeighbours[t]):
            s, m = T.neighbourId[n, t] - 1, T.opposingSend[n, t] - 1
            assert T.neighbourId[m, s] - 1 == t and T.opposingSend[m, s] - 1 == n
            # index maps of the two opposing entries are inverse rotations (w2_set_f2f_index.F:167-200)
            p, q = T.pij[:, n, t], T.pij[:, m, s]
            assert p[0] * q[0] + p[1] * q[2] == 1 and p[0] * q[1] + p[1] * q[3] == 0
            assert p[2] * q[0] + p[3] * q[2] == 0 and p[2] * q[1] + p[3] * q[3] == 1
    # every edge point of every tile feeds exactly one halo cell across each of its edges
    dst, src = halo_gather_map(T, 1, 1, two_pass=False)
    assert len(set(dst.tolist())) == len(dst) == 12 * 2 * (32 + 16)


def test_cs32_halo_cells_are_the_geographic_neighbours():
    """Pins the topology on the reference's own grid files: after the exchange of (xC, yC) the
    great-circle distance between every ring-0 / ring-(sN+1) halo cell centre and the interior cell
    next to it equals the dxC / dyC the files give for that face (to 1e-4 over most of each edge; the
    files use a slightly different edge metric within a few cells of a cube corner, < 1.2 %).  A
    wrong neighbour, rotation or orientation anywhere gives O(1) errors."""
    T, g, _ = load_cs32()
    R, o = 6370e3, 4

    def xyz(lon, lat):
        lo, la = np.deg2rad(lon), np.deg2rad(lat)
        return np.stack([np.cos(la) * np.cos(lo), np.cos(la) * np.sin(lo), np.sin(la)], -1)
    P = xyz(g.xC[0], g.yC[0])
    gc = lambda a, b: R * 2 * np.arcsin(np.minimum(1, 0.5 * np.linalg.norm(a - b, axis=-1)))
    for t in range(12):
        errs = [gc(P[t, o:o + 16, o - 1], P[t, o:o + 16, o]) / g.dxC[0, t, o:o + 16, o] - 1,
                gc(P[t, o:o + 16, o + 32], P[t, o:o + 16, o + 31]) / g.dxC[0, t, o:o + 16, o + 32] - 1,
                gc(P[t, o - 1, o:o + 32], P[t, o, o:o + 32]) / g.dyC[0, t, o, o:o + 32] - 1,
                gc(P[t, o + 16, o:o + 32], P[t, o + 15, o:o + 32]) / g.dyC[0, t, o + 16, o:o + 32] - 1]
        for e in errs:
            assert np.abs(e).max() < 1.2e-2 and np.median(np.abs(e)) < 2e-4
    # interior check of the same formula: the files' dxC is the great-circle distance to 1e-12
    e = gc(P[0, o:o + 16, o + 7], P[0, o:o + 16, o + 8]) / g.dxC[0, 0, o:o + 16, o + 8] - 1
    assert np.abs(e).max() < 1e-12


def test_config4_known_answers():
    """verification/global_ocean.cs32x15/results/output.txt:584-585, 1898: INI_CG2D's normalisation
    factor, the solver tolerance it derives from cg2dTargetResWunit and the global area, from the
    experiment's grid files and bathymetry (R_low / hFac halos through the exch2 exchange)."""
    from mitgcm_b200.model import ini_cg2d_tilegraph
    T, g, P = load_cs32()
    assert f"{P['globalArea']:.15E}" == "3.638867375081599E+14"
    op = ini_cg2d_tilegraph(g, P, T)
    assert f"{op['cg2dNorm']:.16E}" == GOLD4["cg2dNorm"] == "1.9156564154949553E-04"
    assert f"{np.sqrt(op['cg2dTolerance_sq']):.15E}" == "5.809016360175296E-07"
    assert not op["cg2dNormaliseRHS"]


def test_oracle_cg2d_on_the_cube_converges_to_the_configs_tolerance():
    """CG2D (oracle, exchanges through the exch2 hook) on the config-4 operator: symmetric operator
    => monotone convergence to cg2dTolerance in a count comparable with the golden run's 61-62."""
    from mitgcm_b200.model import ini_cg2d_tilegraph
    from oracle.pyoracle import Oracle
    T, g, P = load_cs32()
    d = g.d
    op = ini_cg2d_tilegraph(g, P, T)
    o = Oracle(g, P)
    hook = eo.Exch2Hook(o, T, d.OLx)
    try:
        rng = np.random.default_rng(1)
        jj, ii = d.interior()
        wet = g.maskC[:, :, 0]
        x0 = np.zeros(d.shape2)
        x0[:, :, jj, ii] = rng.standard_normal((1, 12, 16, 32))
        x0 *= wet
        # right-hand side = A x0 for a smooth-ish x0 scaled like the model's (W units)
        b = np.zeros(d.shape2)
        b[:, :, jj, ii] = (rng.standard_normal((1, 12, 16, 32)) * wet[:, :, jj, ii]) * 1e-3
        b[:, :, jj, ii] -= b[:, :, jj, ii].sum() / wet[:, :, jj, ii].sum() * wet[:, :, jj, ii]
        b *= g.rA
        x = np.zeros(d.shape2)
        r = o.cg2d(op, b, x, 200, -1, history=True)
        assert r["numIters"] < 200 and r["lastResidual"] < np.sqrt(op["cg2dTolerance_sq"])
        assert 30 <= r["numIters"] <= 200
    finally:
        hook.close()


@pytest.mark.parametrize("nf,sx,sy,OL", [(8, 4, 4, 2), (32, 32, 16, 4), (6, 6, 6, 3), (16, 8, 4, 2)])
@pytest.mark.parametrize("withSigns", [True, False])
def test_compiled_vector_gather_equals_exch2_uv_3d(nf, sx, sy, OL, withSigns):
    """The library's one-gather form of EXCH_UV_XY(Z) (mitgcm_b200_exch2_uv_map_, host code) against the
    literal restatement of EXCH2_UV_3D_RX: two buffered EXCH2_RX2_CUBE passes with the C-grid offsets of
    EXCH2_GET_UV_BOUNDS, u/v swap and sign across rotated edges, cube-corner fix-ups.  Bit for bit."""
    from mitgcm_b200.exch2 import exchange_uv
    T = cubed_sphere_topology(nf, sx, sy)
    rng = np.random.default_rng(1)
    u = rng.standard_normal((T.nTiles, 2, sy + 2 * OL, sx + 2 * OL))
    v = rng.standard_normal(u.shape)
    u2, v2 = u.copy(), v.copy()
    eo.exch2_uv_3d(T, u, v, OL, withSigns)
    exchange_uv(T, u2, v2, OL, withSigns)
    assert np.array_equal(u, u2) and np.array_equal(v, v2)


def test_cs32_vector_exchange_keeps_a_streamfunction_flow_nondivergent():
    """Pins the vector exchange (swap + sign conventions) on the reference's grid files: volume transports
    derived from a corner streamfunction psi = sin(lat) are set on the interior faces only; after
    EXCH_UV_XYZ(withSigns) the east-halo u-transport and north-halo v-transport of every tile must equal the
    values computed directly from that facet's own corner row/column (index sN+1 of the files), i.e. the
    first ring of cells stays exactly non-divergent across all 12 cube edges."""
    from mitgcm_b200.exch2 import exchange_uv
    T, g, _ = load_cs32()
    o, sx, sy = 4, 32, 16
    psi = np.sin(np.deg2rad(g.yG[0]))                       # (tile, PY, PX), valid on 1..sN+1
    uT = np.zeros_like(psi)
    vT = np.zeros_like(psi)
    uT[:, o:o + sy, o:o + sx + 1] = -(psi[:, o + 1:o + sy + 1, o:o + sx + 1] - psi[:, o:o + sy, o:o + sx + 1])
    vT[:, o:o + sy + 1, o:o + sx] = psi[:, o:o + sy + 1, o + 1:o + sx + 1] - psi[:, o:o + sy + 1, o:o + sx]
    uE, vN = uT[:, o:o + sy, o + sx].copy(), vT[:, o + sy, o:o + sx].copy()    # direct values on index sN+1
    uT[:, :, o + sx:] = 0.0                                 # keep the interior only
    vT[:, o + sy:, :] = 0.0
    exchange_uv(T, uT, vT, o, True)
    scale = np.abs(uE).max()
    assert np.abs(uT[:, o:o + sy, o + sx] - uE).max() < 1e-12 * scale
    assert np.abs(vT[:, o + sy, o:o + sx] - vN).max() < 1e-12 * scale
    div = (uT[:, o:o + sy, o + 1:o + sx + 1] - uT[:, o:o + sy, o:o + sx]) + \
          (vT[:, o + 1:o + sy + 1, o:o + sx] - vT[:, o:o + sy, o:o + sx])
    assert np.abs(div).max() < 1e-12 * scale
