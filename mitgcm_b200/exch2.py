"""Host-side mirror of pkg/exch2's topology set-up (W2_E2SETUP) for the regular 6-facet cube.

What the Fortran model holds in COMMON /W2_EXCH2_TOPO_I/ and /W2_EXCH2_HALO_SPEC/
(pkg/exch2/W2_EXCH2_TOPOLOGY.h:63-122) after W2_SET_CS6_FACETS (w2_set_cs6_facets.F),
W2_SET_MAP_TILES (w2_set_map_tiles.F), W2_SET_F2F_INDEX (w2_set_f2f_index.F) and
W2_SET_TILE2TILES (w2_set_tile2tiles.F) is produced here as numpy int32 tables in the same
(neighbour entry, tile) layout, so they can be handed to `mitgcm_b200_set_exch2_topology_`
exactly as the Fortran shim would hand over the COMMON-block arrays.

Design: everything is an affine index map.  Cell (i, j) of the facet across edge e of facet f,
written in that neighbour's own frame, sits at  M(e,f) . (i, j) + o(e,f)  in f's frame extended
past the edge.  Tile neighbour entries inherit the map of their facet edge.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

N_, S_, E_, W_ = 1, 2, 3, 4   # W2_NORTH .. W2_WEST (W2_EXCH2_TOPOLOGY.h:21-25)


def cs6_links():
    """link[f][e] = (facet, edge) met across edge e of facet f (1-based), the odd/even rule of
    w2_set_cs6_facets.F:60-80."""
    wrap = lambda q: 1 + (q + 5) % 6
    link = {}
    for f in range(1, 7):
        if f % 2 == 1:
            link[f] = {N_: (wrap(f + 2), W_), S_: (wrap(f - 1), N_), E_: (wrap(f + 1), W_), W_: (wrap(f - 2), N_)}
        else:
            link[f] = {N_: (wrap(f + 1), S_), S_: (wrap(f - 2), E_), E_: (wrap(f + 2), S_), W_: (wrap(f - 1), E_)}
    return link


def edge_map(e, ee, dims_f, dims_n):
    """Affine map neighbour frame -> own extended frame for own edge e glued to the neighbour's
    edge ee (w2_set_f2f_index.F:108-165).  Returns (pij[4], oi, oj); dims = (nx, ny)."""
    lo = dims_f[0] if e in (N_, S_) else dims_f[1]     # length of the shared edge
    ident, rotA, rotB = (1, 0, 0, 1), (0, -1, 1, 0), (0, 1, -1, 0)
    if (e, ee) == (N_, S_):
        return ident, 0, dims_f[1]
    if (e, ee) == (S_, N_):
        return ident, 0, -dims_n[1]
    if (e, ee) == (E_, W_):
        return ident, dims_f[0], 0
    if (e, ee) == (W_, E_):
        return ident, -dims_n[0], 0
    if (e, ee) == (N_, W_):
        return rotA, lo + 1, dims_f[1]
    if (e, ee) == (S_, E_):
        return rotA, lo + 1, -dims_n[0]
    if (e, ee) == (E_, S_):
        return rotB, dims_f[0], lo + 1
    if (e, ee) == (W_, N_):
        return rotB, -dims_n[1], lo + 1
    raise ValueError(f"edge connection {e}->{ee} not supported")


def apply(m, i, j):
    p, oi, oj = m
    return p[0] * i + p[1] * j + oi, p[2] * i + p[3] * j + oj


@dataclass
class Exch2Topology:
    """Tables in the reference's layout; index [n, t] = (neighbour entry n+1, tile t+1)."""
    nTiles: int
    maxNeighbours: int
    sNx: int
    sNy: int
    myFace: np.ndarray
    tBasex: np.ndarray
    tBasey: np.ndarray
    nNeighbours: np.ndarray
    neighbourId: np.ndarray
    opposingSend: np.ndarray
    neighbourDir: np.ndarray
    pij: np.ndarray          # [4, n, t]
    oi: np.ndarray
    oj: np.ndarray
    iLo: np.ndarray
    iHi: np.ndarray
    jLo: np.ndarray
    jHi: np.ndarray
    isEdge: dict = field(default_factory=dict)   # 'N','S','E','W' -> int array per tile

    def tables(self):
        """The integer arrays in Fortran memory order (entry index fastest), for the C ABI."""
        f = lambda a: np.ascontiguousarray(a.T if a.ndim == 2 else a.transpose(2, 1, 0), dtype=np.int32)
        return dict(nNeighbours=self.nNeighbours.astype(np.int32), neighbourId=f(self.neighbourId),
                    opposingSend=f(self.opposingSend), neighbourDir=f(self.neighbourDir), pij=f(self.pij),
                    oi=f(self.oi), oj=f(self.oj), iLo=f(self.iLo), iHi=f(self.iHi), jLo=f(self.jLo),
                    jHi=f(self.jHi), tBasex=self.tBasex.astype(np.int32), tBasey=self.tBasey.astype(np.int32),
                    **{"is" + k + "edge": np.ascontiguousarray(self.isEdge[k], dtype=np.int32) for k in "NSEW"})


def cubed_sphere_topology(nFace: int, sNx: int, sNy: int, maxNeighbours: int = 8) -> Exch2Topology:
    """Regular cube of 6 facets nFace x nFace cut into sNx x sNy tiles, tiles numbered facet by
    facet, x fastest (w2_set_map_tiles.F:120-165)."""
    assert nFace % sNx == 0 and nFace % sNy == 0
    link = cs6_links()
    dims = {f: (nFace, nFace) for f in range(1, 7)}
    fmap = {f: {e: edge_map(e, link[f][e][1], dims[f], dims[link[f][e][0]]) for e in (N_, S_, E_, W_)}
            for f in range(1, 7)}
    nbTx, nbTy = nFace // sNx, nFace // sNy
    perFace = nbTx * nbTy
    nT = 6 * perFace
    face = np.zeros(nT, int)
    bx = np.zeros(nT, int)
    by = np.zeros(nT, int)
    owns = {}
    t = 0
    for f in range(1, 7):
        owns[f] = t
        for ty in range(nbTy):
            for tx in range(nbTx):
                face[t], bx[t], by[t] = f, tx * sNx, ty * sNy
                t += 1
    z = lambda: np.zeros((maxNeighbours, nT), int)
    nN = np.zeros(nT, int)
    nid, opp, ndir, oi, oj, iLo, iHi, jLo, jHi = (z() for _ in range(9))
    pij = np.zeros((4, maxNeighbours, nT), int)
    e2e = z()
    isEdge = {k: np.zeros(nT, int) for k in "NSEW"}
    ident = ((1, 0, 0, 1), 0, 0)

    def add(t, tgt, e, ee, m, i1, i2, j1, j2, ddi, ddj):
        n = nN[t]
        assert n < maxNeighbours, "W2_maxNeighbours too small"
        nid[n, t], e2e[n, t] = tgt + 1, 10 * e + ee
        pij[:, n, t] = m[0]
        oi[n, t], oj[n, t] = m[1], m[2]
        iLo[n, t], iHi[n, t] = i1 - ddi - bx[t], i2 + ddi - bx[t]
        jLo[n, t], jHi[n, t] = j1 - ddj - by[t], j2 + ddj - by[t]
        nN[t] += 1

    for t in range(nT):
        f = face[t]
        ilo, ihi, jlo, jhi = bx[t] + 1, bx[t] + sNx, by[t] + 1, by[t] + sNy
        for e in (N_, S_, E_, W_):
            i1, i2, j1, j2 = ilo, ihi, jlo, jhi
            if e == N_:
                j1 = j2 = jhi + 1
                internal = jhi < dims[f][1]
                isEdge["N"][t] = not internal
            elif e == S_:
                j1 = j2 = jlo - 1
                internal = jlo > 1
                isEdge["S"][t] = not internal
            elif e == E_:
                i1 = i2 = ihi + 1
                internal = ihi < dims[f][0]
                isEdge["E"][t] = not internal
            else:
                i1 = i2 = ilo - 1
                internal = ilo > 1
                isEdge["W"][t] = not internal
            ddi, ddj = min(i2 - i1, 1), min(j2 - j1, 1)
            if internal:
                ee = {N_: S_, S_: N_, E_: W_, W_: E_}[e]
                tgt = t + {N_: nbTx, S_: -nbTx, E_: 1, W_: -1}[e]
                add(t, tgt, e, ee, ident, i1, i2, j1, j2, ddi, ddj)
                continue
            nf, ee = link[f][e]
            to_n = fmap[nf][ee]          # own frame -> neighbour frame
            back = fmap[f][e]            # neighbour frame -> own frame
            a1, b1 = apply(to_n, i1, j1)
            a2, b2 = apply(to_n, i2, j2)
            tx1, tx2 = sorted(((a1 - 1) // sNx, (a2 - 1) // sNx))
            ty1, ty2 = sorted(((b1 - 1) // sNy, (b2 - 1) // sNy))
            for ty in range(ty1, ty2 + 1):
                for tx in range(tx1, tx2 + 1):
                    it = owns[nf] + tx + ty * nbTx
                    clipx = lambda v: min(max(v, bx[it] + 1), bx[it] + sNx)
                    clipy = lambda v: min(max(v, by[it] + 1), by[it] + sNy)
                    s1 = apply(back, clipx(a1), clipy(b1))
                    s2 = apply(back, clipx(a2), clipy(b2))
                    add(t, it, e, ee, back, s1[0], s2[0], s1[1], s2[1], ddi, ddj)
    # opposing entries (w2_set_tile2tiles.F:250-290)
    for t in range(nT):
        for n in range(nN[t]):
            e = e2e[n, t] // 10
            ndir[n, t] = e
            it = nid[n, t] - 1
            found = [m for m in range(nN[it]) if nid[m, it] - 1 == t and e2e[m, it] % 10 == e]
            assert len(found) == 1, "tile connection is not one to one"
            opp[n, t] = found[0] + 1
    return Exch2Topology(nT, maxNeighbours, sNx, sNy, face, bx, by, nN, nid, opp, ndir, pij, oi, oj,
                         iLo, iHi, jLo, jHi, isEdge)


def tile_proc(nTiles: int, nRanks: int) -> np.ndarray:
    """W2_tileProc as W2_MAP_PROCS fills it (w2_map_procs.F:60-91): consecutive blocks of nTiles / nRanks tile ids
    per process, 1-based process numbers."""
    assert nTiles % nRanks == 0
    return (np.arange(nTiles) // (nTiles // nRanks) + 1).astype(np.int32)


def set_topology(topo: Exch2Topology, myTileList=None, tileProc=None):
    """Hands the tables to the CUDA library (mitgcm_b200_set_exch2_topology_): from then on the
    width-1 exchange inside CG2D and mitgcm_b200_exch_ follow the exch2 tile graph instead of the
    periodic nSx x nSy tiling.  myTileList(nSx*nSy) = W2_myTileList, default 1..nTiles.
    tileProc(nTiles) = W2_tileProc (1-based owner of every tile): the graph is spread over several ranks
    (after distributed.setup) and myTileList defaults to this rank's tiles."""
    from . import _lib
    L = _lib.lib()
    tb = topo.tables()
    ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))
    ierr = C.c_int(0)
    if tileProc is not None:
        import torch.distributed as dist
        tp = np.ascontiguousarray(tileProc, np.int32)
        L.mitgcm_b200_set_exch2_tile_proc_(C.byref(C.c_int(topo.nTiles)), ip(tp), C.byref(ierr))
        if ierr.value:
            raise RuntimeError(f"set_exch2_tile_proc failed: {L.mitgcm_b200_last_error_string().decode()}")
        if myTileList is None:
            myTileList = np.nonzero(tp == dist.get_rank() + 1)[0] + 1
    tl = np.arange(1, topo.nTiles + 1, dtype=np.int32) if myTileList is None else np.ascontiguousarray(myTileList, np.int32)
    L.mitgcm_b200_set_exch2_topology_(
        C.byref(C.c_int(topo.nTiles)), C.byref(C.c_int(topo.maxNeighbours)), ip(tb["nNeighbours"]),
        ip(tb["neighbourId"]), ip(tb["opposingSend"]), ip(tb["neighbourDir"]), ip(tb["pij"]), ip(tb["oi"]),
        ip(tb["oj"]), ip(tb["iLo"]), ip(tb["iHi"]), ip(tb["jLo"]), ip(tb["jHi"]), ip(tb["tBasex"]),
        ip(tb["tBasey"]), ip(tb["isNedge"]), ip(tb["isSedge"]), ip(tb["isEedge"]), ip(tb["isWedge"]), ip(tl),
        C.byref(ierr))
    if ierr.value:
        raise RuntimeError(f"set_exch2_topology failed: {L.mitgcm_b200_last_error_string().decode()}")
    # facet corners / facet numbers of the local tiles (resident MOM_VECINV: mom_calc_relvort3.F:79-97)
    N, S_, E, W = (np.asarray(tb[n]) for n in ("isNedge", "isSedge", "isEedge", "isWedge"))
    corners = ((W & S_) * 1 + (E & S_) * 2 + (E & N) * 4 + (W & N) * 8).astype(np.int32)[tl - 1]
    faces = np.ascontiguousarray(np.asarray(topo.myFace, dtype=np.int32)[tl - 1])
    edges = np.ascontiguousarray((N * 1 + S_ * 2 + E * 4 + W * 8).astype(np.int32)[tl - 1])
    L.mitgcm_b200_set_cs_tiles_(ip(np.ascontiguousarray(corners)), ip(faces), ip(edges), C.byref(ierr))
    if ierr.value:
        raise RuntimeError(f"set_cs_tiles failed: {L.mitgcm_b200_last_error_string().decode()}")


# ---- the exchange as a gather ----------------------------------------------------------------------
def _target_range(T: Exch2Topology, n: int, t: int, eW: int, corners: bool):
    """Index range of tile t's halo filled through its neighbour entry n (0-based): the topology
    range, cut back by one at both ends when corners are ignored, widened by eW-1 when they are
    updated, and eW deep (pkg/exch2/exch2_get_scal_bounds.F:44-126).  Returns i0, i1, j0, j1."""
    i0, i1, j0, j1 = int(T.iLo[n, t]), int(T.iHi[n, t]), int(T.jLo[n, t]), int(T.jHi[n, t])
    grow = (eW - 1) if corners else -1
    if i0 == i1 and i0 == 0:            # west edge overlap
        i0 = 1 - eW
        j0, j1 = j0 - grow, j1 + grow
    if i0 == i1 and i0 > 1:             # east edge overlap
        i1 = i1 + eW - 1
        j0, j1 = j0 - grow, j1 + grow
    if j0 == j1 and j0 == 0:            # south edge overlap
        j0 = 1 - eW
        i0, i1 = i0 - grow, i1 + grow
    if j0 == j1 and j0 > 1:             # north edge overlap
        j1 = j1 + eW - 1
        i0, i1 = i0 - grow, i1 + grow
    return i0, i1, j0, j1


def halo_gather_map(T: Exch2Topology, OL: int, eW: int | None = None, two_pass: bool = True):
    """The scalar exch2 exchange (EXCH2_3D_RX: an IGNORE_CORNERS pass then an UPDATE_CORNERS pass of
    EXCH2_RX1_CUBE; EXCH2_S3D_RX: the first pass only) compiled into ONE gather.  Each pass copies,
    for every tile and neighbour entry in order, the target range from the source tile through the
    entry's index map, all reads before all writes; composing the two passes cell by cell gives, for
    every halo cell, the cell whose pre-exchange value it ends up holding.  Returns int64 arrays
    (dst, src) of flat indices into an array of shape (nTiles, PY, PX); cells that keep their own
    value are omitted."""
    eW = OL if eW is None else eW
    PX, PY = T.sNx + 2 * OL, T.sNy + 2 * OL
    flat = lambda t, i, j: (t * PY + (j + OL - 1)) * PX + (i + OL - 1)
    prov = np.arange(T.nTiles * PY * PX, dtype=np.int64)
    for corners in ((False, True) if two_pass else (False,)):
        new = prov.copy()
        for t in range(T.nTiles):
            for n in range(int(T.nNeighbours[t])):
                s = int(T.neighbourId[n, t]) - 1
                m = int(T.opposingSend[n, t]) - 1
                p = T.pij[:, m, s]
                i0, i1, j0, j1 = _target_range(T, n, t, eW, corners)
                ii, jj = np.meshgrid(np.arange(i0, i1 + 1), np.arange(j0, j1 + 1))
                if ii.size == 0:
                    continue
                ic, jc = ii + int(T.tBasex[t]), jj + int(T.tBasey[t])
                si = p[0] * ic + p[1] * jc + int(T.oi[m, s]) - int(T.tBasex[s])
                sj = p[2] * ic + p[3] * jc + int(T.oj[m, s]) - int(T.tBasey[s])
                assert si.min() >= 1 - OL and si.max() <= T.sNx + OL and sj.min() >= 1 - OL and sj.max() <= T.sNy + OL
                new[flat(t, ii, jj).ravel()] = prov[flat(s, si, sj).ravel()]
        prov = new
    dst = np.nonzero(prov != np.arange(prov.size))[0]
    return dst, prov[dst]


def exchange(T: Exch2Topology, a: np.ndarray, OL: int, gmap=None) -> np.ndarray:
    """EXCH_XY_RL / EXCH_XYZ_RL on the tile graph for a host array (nTiles, [nz,] PY, PX); in place."""
    dst, src = gmap if gmap is not None else halo_gather_map(T, OL)
    if a.ndim == 3:
        f = a.reshape(-1)
        f[dst] = f[src]
    else:
        nT, nz, PY, PX = a.shape
        v = a.transpose(1, 0, 2, 3).reshape(nz, -1)   # copy: (nz, nT*PY*PX)
        v[:, dst] = v[:, src]
        a[...] = v.reshape(nz, nT, PY, PX).transpose(1, 0, 2, 3)
    return a


def uv_gather_map(T: Exch2Topology, OL: int, withSigns: bool):
    """EXCH_UV_XY / EXCH_UV_XYZ on the tile graph (EXCH2_UV_3D_RX) as one gather, compiled by the library
    (mitgcm_b200_exch2_uv_map_, host code: works without a GPU).  Returns int arrays
    (dstArr, dst, srcArr, neg, src): component 0 = u, 1 = v; flat indices into (nTiles, PY, PX)."""
    from . import _lib
    L = _lib.lib()
    tb = T.tables()
    ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))
    cap = 2 * T.nTiles * (T.sNx + 2 * OL) * (T.sNy + 2 * OL)
    out = np.zeros(4 * cap, dtype=np.int32)
    n, ierr = C.c_int(0), C.c_int(0)
    dims3 = (C.c_int * 3)(T.sNx, T.sNy, OL)
    L.mitgcm_b200_exch2_uv_map_(
        dims3, C.byref(C.c_int(int(withSigns))), C.byref(C.c_int(T.nTiles)), C.byref(C.c_int(T.maxNeighbours)),
        ip(tb["nNeighbours"]), ip(tb["neighbourId"]), ip(tb["opposingSend"]), ip(tb["neighbourDir"]), ip(tb["pij"]),
        ip(tb["oi"]), ip(tb["oj"]), ip(tb["iLo"]), ip(tb["iHi"]), ip(tb["jLo"]), ip(tb["jHi"]), ip(tb["tBasex"]),
        ip(tb["tBasey"]), ip(tb["isNedge"]), ip(tb["isSedge"]), ip(tb["isEedge"]), ip(tb["isWedge"]),
        C.byref(C.c_int(cap)), C.byref(n), ip(out), C.byref(ierr))
    if ierr.value:
        raise RuntimeError(f"exch2_uv_map failed: {L.mitgcm_b200_last_error_string().decode()}")
    e = out[:4 * n.value].reshape(-1, 4)
    return e[:, 0].copy(), e[:, 1].astype(np.int64), e[:, 2] >> 1, (e[:, 2] & 1).astype(bool), e[:, 3].astype(np.int64)


def dist_lists(T: Exch2Topology, OL: int, nRanks: int, myRank: int, withSigns: bool = True, tileProc=None):
    """What rank myRank of nRanks is given when the tile graph is spread over ranks (mitgcm_b200_exch2_dist_lists_,
    host code: works without a GPU): (scalar (n,2): dst, owner << 28 | src;  push (nLocalTiles, 2 sNy + 2 sNx):
    owner << 28 | halo index;  uv (n,4): dst array, dst, src array << 1 | negate, owner << 28 | src), indices relative to
    the owning rank's (nTiles / nRanks, PY, PX) arrays -- exactly the lists the GPU kernels read."""
    from . import _lib
    L = _lib.lib()
    tb = T.tables()
    ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))
    tp = np.ascontiguousarray(tile_proc(T.nTiles, nRanks) if tileProc is None else tileProc, np.int32)
    nL = T.nTiles // nRanks
    cells = nL * (T.sNx + 2 * OL) * (T.sNy + 2 * OL)
    sizes = (C.c_int * 3)(2 * cells, nL * (2 * T.sNx + 2 * T.sNy), 8 * cells)
    sc, pu, uv = (np.zeros(n, dtype=np.int32) for n in sizes)
    ierr = C.c_int(0)
    dims3 = (C.c_int * 3)(T.sNx, T.sNy, OL)
    L.mitgcm_b200_exch2_dist_lists_(
        dims3, C.byref(C.c_int(nRanks)), C.byref(C.c_int(myRank)), ip(tp), C.byref(C.c_int(int(withSigns))),
        C.byref(C.c_int(T.nTiles)), C.byref(C.c_int(T.maxNeighbours)),
        ip(tb["nNeighbours"]), ip(tb["neighbourId"]), ip(tb["opposingSend"]), ip(tb["neighbourDir"]), ip(tb["pij"]),
        ip(tb["oi"]), ip(tb["oj"]), ip(tb["iLo"]), ip(tb["iHi"]), ip(tb["jLo"]), ip(tb["jHi"]), ip(tb["tBasex"]),
        ip(tb["tBasey"]), ip(tb["isNedge"]), ip(tb["isSedge"]), ip(tb["isEedge"]), ip(tb["isWedge"]),
        sizes, ip(sc), ip(pu), ip(uv), C.byref(ierr))
    if ierr.value:
        raise RuntimeError(f"exch2_dist_lists failed: {L.mitgcm_b200_last_error_string().decode()}")
    return sc[:sizes[0]].reshape(-1, 2), pu[:sizes[1]].reshape(nL, -1), uv[:sizes[2]].reshape(-1, 4)


def exchange_uv(T: Exch2Topology, u: np.ndarray, v: np.ndarray, OL: int, withSigns: bool, gmap=None):
    """In-place EXCH_UV_XY(Z) of host arrays (nTiles, [nz,] PY, PX)."""
    da, d, sa, neg, s = gmap if gmap is not None else uv_gather_map(T, OL, withSigns)
    two_d = u.ndim == 3
    U = u[:, None] if two_d else u
    V = v[:, None] if two_d else v
    nT, nz, PY, PX = U.shape
    fu = U.transpose(1, 0, 2, 3).reshape(nz, -1).copy()
    fv = V.transpose(1, 0, 2, 3).reshape(nz, -1).copy()
    src = np.where(sa[None, :] == 0, fu[:, s], fv[:, s])
    src = np.where(neg[None, :], -src, src)
    mu, mv = da == 0, da == 1
    fu[:, d[mu]] = src[:, mu]
    fv[:, d[mv]] = src[:, mv]
    U[...] = fu.reshape(nz, nT, PY, PX).transpose(1, 0, 2, 3)
    V[...] = fv.reshape(nz, nT, PY, PX).transpose(1, 0, 2, 3)
    return u, v
