"""CPU restatement of pkg/exch2's scalar exchange on a tile graph (single process, commSetting 'P').

TEST INFRASTRUCTURE ONLY.  Literal, buffer-based restatement of
  EXCH2_GET_SCAL_BOUNDS   pkg/exch2/exch2_get_scal_bounds.F:44-126
  EXCH2_PUT_RX1           pkg/exch2/exch2_put_rx1.template:120-190 (index map isc = pij.(itc,jtc)+o)
  EXCH2_GET_RX1           pkg/exch2/exch2_get_rx1.template:100-140
  EXCH2_RX1_CUBE          pkg/exch2/exch2_rx1_cube.template:95-262 (all PUTs, then all GETs)
  EXCH2_3D_RX             pkg/exch2/exch2_3d_rx.template:60-80   (IGNORE_CORNERS pass, UPDATE_CORNERS pass)
  EXCH2_S3D_RX            pkg/exch2/exch2_s3d_rx.template:45-62  (width 1, IGNORE_CORNERS only)
operating on the topology tables of W2_EXCH2_TOPOLOGY.h (any generator).  Arrays are numpy
(nTiles, nz, sNy+2*OL, sNx+2*OL); pure-Python loops: for small cases only.

Parity status: no reference golden exercises the exchange directly; it is pinned indirectly by the
reference's own cubed-sphere grid files (tests/test_exch2_cpu.py: metric continuity across all 12
cube edges against the dxC/dyC the files hold for the edge points).
"""
from __future__ import annotations

import numpy as np


def get_scal_bounds(T, eWdth, updateCorners, tgTile, tgNb):
    """tgTile, tgNb 0-based. Returns tIlo, tIhi, tJlo, tJhi, tiStride, tjStride."""
    tIlo, tIhi = int(T.iLo[tgNb, tgTile]), int(T.iHi[tgNb, tgTile])
    tJlo, tJhi = int(T.jLo[tgNb, tgTile]), int(T.jHi[tgNb, tgTile])
    tiStride = tjStride = 1
    if tIlo == tIhi and tIlo == 0:                       # west edge overlap
        tIlo = 1 - eWdth
        tiStride = 1
        tjStride = 1 if tJlo <= tJhi else -1
        if updateCorners:
            tJlo, tJhi = tJlo - tjStride * (eWdth - 1), tJhi + tjStride * (eWdth - 1)
        else:
            tJlo, tJhi = tJlo + tjStride, tJhi - tjStride
    if tIlo == tIhi and tIlo > 1:                        # east edge overlap
        tIhi = tIhi + eWdth - 1
        tiStride = 1
        tjStride = 1 if tJlo <= tJhi else -1
        if updateCorners:
            tJlo, tJhi = tJlo - tjStride * (eWdth - 1), tJhi + tjStride * (eWdth - 1)
        else:
            tJlo, tJhi = tJlo + tjStride, tJhi - tjStride
    if tJlo == tJhi and tJlo == 0:                       # south edge overlap
        tJlo = 1 - eWdth
        tjStride = 1
        tiStride = 1 if tIlo <= tIhi else -1
        if updateCorners:
            tIlo, tIhi = tIlo - tiStride * (eWdth - 1), tIhi + tiStride * (eWdth - 1)
        else:
            tIlo, tIhi = tIlo + tiStride, tIhi - tiStride
    if tJlo == tJhi and tJlo > 1:                        # north edge overlap
        tJhi = tJhi + eWdth - 1
        tjStride = 1
        tiStride = 1 if tIlo <= tIhi else -1
        if updateCorners:
            tIlo, tIhi = tIlo - tiStride * (eWdth - 1), tIhi + tiStride * (eWdth - 1)
        else:
            tIlo, tIhi = tIlo + tiStride, tIhi - tiStride
    return tIlo, tIhi, tJlo, tJhi, tiStride, tjStride


def _rng(lo, hi, st):
    return range(lo, hi + (1 if st > 0 else -1), st)


def rx1_cube(T, arr, OL, eWdth, updateCorners):
    """One EXCH2_RX1_CUBE pass on arr (nTiles, nz, PY, PX) with overlap OL: every tile PUTs for
    each of its neighbour entries into a buffer, then every tile GETs.  In place."""
    nT = T.nTiles
    bufs = {}
    for src in range(nT):                                  # thisTile = source
        for N in range(int(T.nNeighbours[src])):
            tg = int(T.neighbourId[N, src]) - 1
            oN = int(T.opposingSend[N, src]) - 1
            tIlo, tIhi, tJlo, tJhi, si, sj = get_scal_bounds(T, eWdth, updateCorners, tg, oN)
            p = T.pij[:, N, src]
            oi, oj = int(T.oi[N, src]), int(T.oj[N, src])
            itb, jtb, isb, jsb = int(T.tBasex[tg]), int(T.tBasey[tg]), int(T.tBasex[src]), int(T.tBasey[src])
            vals = []
            for jtl in _rng(tJlo, tJhi, sj):
                for itl in _rng(tIlo, tIhi, si):
                    itc, jtc = itl + itb, jtl + jtb
                    isl = int(p[0]) * itc + int(p[1]) * jtc + oi - isb
                    jsl = int(p[2]) * itc + int(p[3]) * jtc + oj - jsb
                    assert 1 - OL <= isl <= T.sNx + OL and 1 - OL <= jsl <= T.sNy + OL, "source out of bounds"
                    vals.append(arr[src, :, jsl + OL - 1, isl + OL - 1].copy())
            bufs[(src, N)] = vals
    for tg in range(nT):                                   # thisTile = target
        for N in range(int(T.nNeighbours[tg])):
            tIlo, tIhi, tJlo, tJhi, si, sj = get_scal_bounds(T, eWdth, updateCorners, tg, N)
            src = int(T.neighbourId[N, tg]) - 1
            oNb = int(T.opposingSend[N, tg]) - 1
            vals = bufs[(src, oNb)]
            q = 0
            for jtl in _rng(tJlo, tJhi, sj):
                for itl in _rng(tIlo, tIhi, si):
                    arr[tg, :, jtl + OL - 1, itl + OL - 1] = vals[q]
                    q += 1
            assert q == len(vals)
    return arr


def exch2_3d(T, arr, OL):
    """EXCH2_3D_RL / EXCH_XY_RL / EXCH_XYZ_RL on the tile graph: full-width halo, two passes."""
    rx1_cube(T, arr, OL, OL, False)
    rx1_cube(T, arr, OL, OL, True)
    return arr


def exch2_s3d(T, arr):
    """EXCH2_S3D_RL: arrays (0:sNx+1, 0:sNy+1), width 1, corners ignored."""
    return rx1_cube(T, arr, 1, 1, False)


def probe_map(T, OL, s3d=False):
    """(dst, src) flat-index lists of the exchange, obtained by running the literal restatement on
    a field whose value is its own flat index (exact in float64)."""
    if s3d:
        shape = (T.nTiles, 1, T.sNy + 2, T.sNx + 2)
        a = np.arange(np.prod(shape), dtype=np.float64).reshape(shape)
        exch2_s3d(T, a)
    else:
        shape = (T.nTiles, 1, T.sNy + 2 * OL, T.sNx + 2 * OL)
        a = np.arange(np.prod(shape), dtype=np.float64).reshape(shape)
        exch2_3d(T, a, OL)
    now = a.reshape(-1).astype(np.int64)
    dst = np.nonzero(now != np.arange(now.size))[0].astype(np.int64)
    return dst, now[dst]


class Exch2Hook:
    """Installs the tile graph into the C oracle (og_set_exch2_maps) for the life of the object."""

    def __init__(self, oracle, T, OL):
        import ctypes as C
        self.o = oracle
        self.maps = [np.ascontiguousarray(m) for m in (*probe_map(T, OL), *probe_map(T, OL, s3d=True))]
        p = lambda a: a.ctypes.data_as(C.POINTER(C.c_longlong))
        m = self.maps
        oracle.lib.og_set_exch2_maps(len(m[0]), p(m[0]), p(m[1]), len(m[2]), p(m[2]), p(m[3]))

    def close(self):
        self.o.lib.og_set_exch2_maps(0, None, None, 0, None, None)


# ---- vector pairs on the C grid: EXCH2_RX2_CUBE -----------------------------------------------------
def get_uv_bounds(T, eWdth, updateCorners, tgTile, tgNb):
    """EXCH2_GET_UV_BOUNDS, fCode 'Cg' (pkg/exch2/exch2_get_uv_bounds.F:60-262).  0-based tile / entry.
    Returns (range1, range2, tiStride, tjStride, (oi1, oj1, oi2, oj2)); range = [ilo, ihi, jlo, jhi]."""
    tIlo, tIhi = int(T.iLo[tgNb, tgTile]), int(T.iHi[tgNb, tgTile])
    tJlo, tJhi = int(T.jLo[tgNb, tgTile]), int(T.jHi[tgNb, tgTile])
    soNb = int(T.opposingSend[tgNb, tgTile]) - 1
    soTile = int(T.neighbourId[tgNb, tgTile]) - 1
    oi1, oj1 = int(T.oi[soNb, soTile]), int(T.oj[soNb, soTile])
    pij = [int(v) for v in T.pij[:, soNb, soTile]]
    tIlo1 = tIhi1 = tJlo1 = tJhi1 = 0
    tiStride = tjStride = 1
    if tIlo == tIhi and tIlo == 0:                      # west edge overlap
        tIlo1, tIhi1 = 1 - eWdth, 0
        tiStride = 1
        tjStride = 1 if tJlo <= tJhi else -1
        if updateCorners:
            tJlo1, tJhi1 = tJlo - tjStride * (eWdth - 1), tJhi + tjStride * (eWdth - 1)
        else:
            tJlo1, tJhi1 = tJlo + tjStride, tJhi - tjStride
    if tIlo == tIhi and tIlo > 1:                       # east edge overlap
        tIlo1, tIhi1 = tIlo, tIhi + eWdth - 1
        tiStride = 1
        tjStride = 1 if tJlo <= tJhi else -1
        if updateCorners:
            tJlo1, tJhi1 = tJlo - tjStride * (eWdth - 1), tJhi + tjStride * (eWdth - 1)
        else:
            tJlo1, tJhi1 = tJlo + tjStride, tJhi - tjStride
    if tJlo == tJhi and tJlo == 0:                      # south edge overlap
        tJlo1, tJhi1 = 1 - eWdth, 0
        tjStride = 1
        tiStride = 1 if tIlo <= tIhi else -1
        if updateCorners:
            tIlo1, tIhi1 = tIlo - tiStride * (eWdth - 1), tIhi + tiStride * (eWdth - 1)
        else:
            tIlo1, tIhi1 = tIlo + tiStride, tIhi - tiStride
    if tJlo == tJhi and tJlo > 1:                       # north edge overlap
        tJlo1, tJhi1 = tJlo, tJhi + eWdth - 1
        tjStride = 1
        tiStride = 1 if tIlo <= tIhi else -1
        if updateCorners:
            tIlo1, tIhi1 = tIlo - tiStride * (eWdth - 1), tIhi + tiStride * (eWdth - 1)
        else:
            tIlo1, tIhi1 = tIlo + tiStride, tIhi - tiStride
    tIlo2, tIhi2, tJlo2, tJhi2 = tIlo1, tIhi1, tJlo1, tJhi1
    oi2, oj2 = oi1, oj1
    # UV C-grid specific part
    if pij[0] == -1:
        oi1 += 1
    if pij[2] == -1:
        oj1 += 1
    if pij[1] == -1:
        oi2 += 1
    if pij[3] == -1:
        oj2 += 1
    if updateCorners:
        if pij[0] == -1 or pij[2] == -1:
            tIlo1 += 1
        if pij[1] == -1 or pij[3] == -1:
            tJlo2 += 1
        if tIlo == tIhi and tIlo > 1:                   # east edge touching the face S / N edge
            if T.isEdge["S"][tgTile] == 1:
                tJlo1 = tJlo + 1
                tJlo2 = tJlo + 1
            if T.isEdge["N"][tgTile] == 1:
                tJhi1 = tJhi - 1
                tJhi2 = tJhi
        if tJlo == tJhi and tJlo > 1:                   # north edge touching the face W / E edge
            if T.isEdge["W"][tgTile] == 1:
                tIlo1 = tIlo + 1
                tIlo2 = tIlo + 1
            if T.isEdge["E"][tgTile] == 1:
                tIhi1 = tIhi
                tIhi2 = tIhi - 1
    else:
        if pij[0] == -1 or pij[2] == -1:
            tIlo1 += 1
            tIhi1 += 1
        if pij[1] == -1 or pij[3] == -1:
            tJlo2 += 1
            tJhi2 += 1
    return [tIlo1, tIhi1, tJlo1, tJhi1], [tIlo2, tIhi2, tJlo2, tJhi2], tiStride, tjStride, (oi1, oj1, oi2, oj2)


def rx2_cube(T, u, v, OL, eWdth, updateCorners, withSigns):
    """One EXCH2_RX2_CUBE pass ('Cg') on u, v (nTiles, nz, PY, PX): all PUTs (exch2_put_rx2.template),
    then all GETs (exch2_get_rx2.template)."""
    nT = T.nTiles
    bufs = {}
    for src in range(nT):
        for N in range(int(T.nNeighbours[src])):
            tg = int(T.neighbourId[N, src]) - 1
            oN = int(T.opposingSend[N, src]) - 1
            r1, r2, si, sj, (oIs1, oJs1, oIs2, oJs2) = get_uv_bounds(T, eWdth, updateCorners, tg, oN)
            p = [int(x) for x in T.pij[:, N, src]]
            itb, jtb, isb, jsb = int(T.tBasex[tg]), int(T.tBasey[tg]), int(T.tBasex[src]), int(T.tBasey[src])
            out = []
            for comp, (r, oIs, oJs) in enumerate(((r1, oIs1, oJs1), (r2, oIs2, oJs2))):
                sa1, sa2 = (p[0], p[2]) if comp == 0 else (p[1], p[3])
                if not withSigns:
                    sa1, sa2 = abs(sa1), abs(sa2)
                vals = []
                for jtl in _rng(r[2], r[3], sj):
                    for itl in _rng(r[0], r[1], si):
                        itc, jtc = itl + itb, jtl + jtb
                        isl = p[0] * itc + p[1] * jtc + oIs - isb
                        jsl = p[2] * itc + p[3] * jtc + oJs - jsb
                        assert 1 - OL <= isl <= T.sNx + OL and 1 - OL <= jsl <= T.sNy + OL, "source out of bounds"
                        vals.append(float(sa1) * u[src, :, jsl + OL - 1, isl + OL - 1]
                                    + float(sa2) * v[src, :, jsl + OL - 1, isl + OL - 1])
                out.append(vals)
            bufs[(src, N)] = out
    for tg in range(nT):
        for N in range(int(T.nNeighbours[tg])):
            r1, r2, si, sj, _ = get_uv_bounds(T, eWdth, updateCorners, tg, N)
            src = int(T.neighbourId[N, tg]) - 1
            oNb = int(T.opposingSend[N, tg]) - 1
            b1, b2 = bufs[(src, oNb)]
            for arr, r, vals in ((u, r1, b1), (v, r2, b2)):
                q = 0
                for jtl in _rng(r[2], r[3], sj):
                    for itl in _rng(r[0], r[1], si):
                        arr[tg, :, jtl + OL - 1, itl + OL - 1] = vals[q]
                        q += 1
                assert q == len(vals)


def exch2_uv_3d(T, u, v, OL, withSigns):
    """EXCH_UV_XY(Z)_RL/RS on the tile graph = EXCH2_UV_3D_RX (pkg/exch2/exch2_uv_3d_rx.template:60-230):
    two EXCH2_RX2_CUBE passes, then the four cube-corner fix-ups (needs OL >= 2)."""
    rx2_cube(T, u, v, OL, OL, False, withSigns)
    rx2_cube(T, u, v, OL, OL, True, withSigns)
    sx, sy = T.sNx, T.sNy
    A = lambda a, t, i, j: a[t, :, j + OL - 1, i + OL - 1]
    sg = -1.0 if withSigns else 1.0
    for t in range(T.nTiles):
        W, E, S, N = (T.isEdge[k][t] == 1 for k in "WESN")
        if OL < 2:
            continue
        if W and S:
            u[t, :, 0 + OL - 1, 0 + OL - 1] = A(v, t, 1, 0)
            v[t, :, 0 + OL - 1, 0 + OL - 1] = A(u, t, 0, 1)
        if W and N:
            u[t, :, sy + 1 + OL - 1, 0 + OL - 1] = sg * A(v, t, 1, sy + 2)
            v[t, :, sy + 2 + OL - 1, 0 + OL - 1] = sg * A(u, t, 0, sy)
        if E and S:
            u[t, :, 0 + OL - 1, sx + 2 + OL - 1] = sg * A(v, t, sx, 0)
            v[t, :, 0 + OL - 1, sx + 1 + OL - 1] = sg * A(u, t, sx + 2, 1)
        if E and N:
            u[t, :, sy + 1 + OL - 1, sx + 2 + OL - 1] = A(v, t, sx, sy + 2)
            v[t, :, sy + 2 + OL - 1, sx + 1 + OL - 1] = A(u, t, sx + 2, sy)
    return u, v
