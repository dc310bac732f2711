"""Tile layout, grid metrics and masks in the reference's own memory layout.

This is host-side set-up (numpy), the counterpart of what the Fortran model has
already done before the hot path is entered: INI_CARTESIAN_GRID
(model/src/ini_cartesian_grid.F:56-140), INI_LOCAL_GRID (ini_local_grid.F:100-165),
INI_MASKS_ETC (ini_masks_etc.F:100-300, 400-480), INI_CORI (ini_cori.F:60-90),
INI_LINEAR_PHISURF (ini_linear_phisurf.F:84-85).  Arrays are C-contiguous numpy
arrays of shape (nSy, nSx, [Nr,] sNy+2*OLy, sNx+2*OLx), which is byte-for-byte
the Fortran layout (1-OLx:sNx+OLx, 1-OLy:sNy+OLy, [Nr,] nSx, nSy) of
model/inc/SIZE.h / GRID.h, so the same buffers can be handed to the C-ABI.
"""
from __future__ import annotations

from dataclasses import dataclass, field
import numpy as np

GRID2D = ("dxC dyC dxG dyG dxF dyF dxV dyU rA rAw rAs rAz "
          "recip_dxC recip_dyC recip_dxG recip_dyG recip_dxF recip_dyF recip_dxV recip_dyU "
          "recip_rA recip_rAw recip_rAs recip_rAz fCori fCoriG tanPhiAtU tanPhiAtV recip_Bo Bo_surf").split()
GRID3D = "hFacC hFacW hFacS recip_hFacC recip_hFacW recip_hFacS maskC maskW maskS".split()
GRID1D = "drF drC recip_drF recip_drC".split()
GRIDJ = "cosFacU cosFacV".split()


@dataclass(frozen=True)
class Dims:
    """Compile-time sizes of model/inc/SIZE.h for one process (one GPU)."""
    sNx: int
    sNy: int
    OLx: int
    OLy: int
    nSx: int = 1
    nSy: int = 1
    Nr: int = 1
    nPx: int = 1
    nPy: int = 1
    myPx: int = 0   # this process's position in the nPx x nPy process grid
    myPy: int = 0

    @property
    def PX(self):
        return self.sNx + 2 * self.OLx

    @property
    def PY(self):
        return self.sNy + 2 * self.OLy

    @property
    def shape2(self):
        return (self.nSy, self.nSx, self.PY, self.PX)

    @property
    def shape3(self):
        return (self.nSy, self.nSx, self.Nr, self.PY, self.PX)

    def shape3n(self, n):
        return (self.nSy, self.nSx, n, self.PY, self.PX)

    @property
    def Nx(self):
        return self.sNx * self.nSx * self.nPx

    @property
    def Ny(self):
        return self.sNy * self.nSy * self.nPy

    def interior(self):
        return (slice(self.OLy, self.OLy + self.sNy), slice(self.OLx, self.OLx + self.sNx))


def exch_xyz(d: Dims, a: np.ndarray) -> np.ndarray:
    """EXCH_XY(Z)_RL on one periodic process (eesupp/src/exch1_rx.template:170-201):
    X phase first, then Y phase over the full X range so corners propagate.
    `a` has shape (nSy, nSx, ..., PY, PX); updated in place."""
    assert d.nPx == 1 and d.nPy == 1, "host exchange helper is single-process"
    ox, oy, sx, sy = d.OLx, d.OLy, d.sNx, d.sNy
    for bj in range(d.nSy):
        for bi in range(d.nSx):
            bw, be = (bi - 1) % d.nSx, (bi + 1) % d.nSx
            a[bj, bi, ..., oy:oy + sy, 0:ox] = a[bj, bw, ..., oy:oy + sy, sx:sx + ox]
            a[bj, bi, ..., oy:oy + sy, ox + sx:] = a[bj, be, ..., oy:oy + sy, ox:2 * ox]
    tmp = a.copy()
    for bj in range(d.nSy):
        bs, bn = (bj - 1) % d.nSy, (bj + 1) % d.nSy
        a[bj, :, ..., 0:oy, :] = tmp[bs, :, ..., sy:sy + oy, :]
        a[bj, :, ..., oy + sy:, :] = tmp[bn, :, ..., oy:2 * oy, :]
    return a


@dataclass
class Grid:
    d: Dims
    a: dict = field(default_factory=dict)

    def __getattr__(self, name):
        try:
            return self.__dict__["a"][name]
        except KeyError as e:
            raise AttributeError(name) from e

    def set_recips(self):
        for n in "dxC dyC dxG dyG dxF dyF dxV dyU rA rAw rAs rAz".split():
            v = self.a[n]
            r = np.zeros_like(v)
            np.divide(1.0, v, out=r, where=v != 0.0)   # ini_grid.F zero-guarded reciprocals
            self.a["recip_" + n] = r


def cartesian_grid(d: Dims, delX, delY, delR, xgOrigin=0.0, ygOrigin=0.0,
                   f0=1e-4, beta=1e-11, gBaro=9.81, fsin=None) -> Grid:
    """usingCartesianGrid: INI_CARTESIAN_GRID + INI_CORI(selectCoriMap=1) +
    INI_LINEAR_PHISURF.  delX/delY are the global spacings (length Nx/Ny).
    fsin = (amplitude, period in m): f = f0 + amplitude * sin(2 pi y / period) instead of the beta plane -- what a
    selectCoriMap = 3 run reads from fCoriC / fCoriG files (ini_cori.F:119-178); periodic in y, so a doubly
    periodic domain has no jump of f across its seam."""
    delX = np.asarray(delX, dtype=np.float64)
    delY = np.asarray(delY, dtype=np.float64)
    delR = np.asarray(delR, dtype=np.float64)
    assert len(delX) == d.Nx and len(delY) == d.Ny and len(delR) == d.Nr
    g = Grid(d)
    z2 = lambda: np.zeros(d.shape2)
    for n in GRID2D:
        g.a[n] = z2()
    xC, yC, xG, yG = z2(), z2(), z2(), z2()
    ox, oy, sx, sy = d.OLx, d.OLy, d.sNx, d.sNy
    for bj in range(d.nSy):
        for bi in range(d.nSx):
            iG0 = (d.myPx * d.nSx + bi) * sx
            jG0 = (d.myPy * d.nSy + bj) * sy
            # sequential sums, same order as ini_local_grid.F:127-145
            xG0 = np.float64(xgOrigin)
            for i in range(iG0):
                xG0 = xG0 + delX[i]
            for i in range(1, ox + 1):
                xG0 = xG0 - delX[(iG0 - i + ox * d.Nx) % d.Nx]
            yG0 = np.float64(ygOrigin)
            for j in range(jG0):
                yG0 = yG0 + delY[j]
            for j in range(1, oy + 1):
                yG0 = yG0 - delY[(jG0 - j + oy * d.Ny) % d.Ny]
            dXl = delX[[(iG0 + i - 1 + ox * d.Nx) % d.Nx for i in range(1 - ox, sx + ox + 1)]]
            dYl = delY[[(jG0 + j - 1 + oy * d.Ny) % d.Ny for j in range(1 - oy, sy + oy + 1)]]
            xGl = np.empty(d.PX + 1)
            xGl[0] = xG0
            for i in range(d.PX):
                xGl[i + 1] = xGl[i] + dXl[i]
            yGl = np.empty(d.PY + 1)
            yGl[0] = yG0
            for j in range(d.PY):
                yGl[j + 1] = yGl[j] + dYl[j]
            XG = np.broadcast_to(xGl[None, :], (d.PY + 1, d.PX + 1))
            YG = np.broadcast_to(yGl[:, None], (d.PY + 1, d.PX + 1))
            xG[bj, bi] = XG[:-1, :-1]
            yG[bj, bi] = YG[:-1, :-1]
            xC[bj, bi] = 0.25 * (((XG[:-1, :-1] + XG[:-1, 1:]) + XG[1:, :-1]) + XG[1:, 1:])
            yC[bj, bi] = 0.25 * (((YG[:-1, :-1] + YG[:-1, 1:]) + YG[1:, :-1]) + YG[1:, 1:])
            dxF = np.broadcast_to(dXl[None, :], (d.PY, d.PX)).copy()
            dyF = np.broadcast_to(dYl[:, None], (d.PY, d.PX)).copy()
            g.a["dxF"][bj, bi] = dxF
            g.a["dyF"][bj, bi] = dyF
            g.a["dxG"][bj, bi] = dxF
            g.a["dyG"][bj, bi] = dyF
            g.a["dxC"][bj, bi][:, 1:] = 0.5 * (dxF[:, 1:] + dxF[:, :-1])
            g.a["dyC"][bj, bi][1:, :] = 0.5 * (dyF[1:, :] + dyF[:-1, :])
            g.a["dxV"][bj, bi][1:, 1:] = 0.5 * (dxF[1:, 1:] + dxF[1:, :-1])
            g.a["dyU"][bj, bi][1:, 1:] = 0.5 * (dyF[1:, 1:] + dyF[:-1, 1:])
            g.a["rA"][bj, bi] = dxF * dyF
            g.a["rAw"][bj, bi] = g.a["dxC"][bj, bi] * dyF
            g.a["rAs"][bj, bi] = dxF * g.a["dyC"][bj, bi]
            g.a["rAz"][bj, bi] = g.a["dxV"][bj, bi] * g.a["dyU"][bj, bi]
    g.a["xC"], g.a["yC"], g.a["xG"], g.a["yG"] = xC, yC, xG, yG
    if fsin is None:
        g.a["fCori"] = f0 + beta * yC * 1.0
        g.a["fCoriG"] = f0 + beta * yG * 1.0
    else:
        amp, period = fsin
        g.a["fCori"] = f0 + amp * np.sin(2.0 * np.pi * yC / period)
        g.a["fCoriG"] = f0 + amp * np.sin(2.0 * np.pi * yG / period)
    g.a["Bo_surf"] = np.full(d.shape2, gBaro)
    g.a["recip_Bo"] = np.full(d.shape2, 1.0 / gBaro)
    g.a["cosFacU"] = np.ones((d.nSy, d.nSx, d.PY))
    g.a["cosFacV"] = np.ones((d.nSy, d.nSx, d.PY))
    g.set_recips()
    set_vertical(g, delR)
    return g


def spherical_polar_grid(d: Dims, delX, delY, delR, xgOrigin=0.0, ygOrigin=0.0,
                         rSphere=6370e3, rotationPeriod=86164.0, gBaro=9.81, cosPower=0.0) -> Grid:
    """usingSphericalPolarGrid: INI_SPHERICAL_POLAR_GRID (model/src/ini_spherical_polar_grid.F:60-230)
    + INI_LOCAL_GRID (ini_local_grid.F:100-165) + INI_CORI with selectCoriMap = 2 (ini_cori.F) +
    INI_LINEAR_PHISURF.  delX/delY in degrees.  Transcendentals go through libm (math.sin/cos/tan),
    which is what the gfortran-built reference calls."""
    import math
    assert cosPower == 0.0, "cosPower != 0 (cosFacU/V) not restated"
    delX = np.asarray(delX, dtype=np.float64)
    delY = np.asarray(delY, dtype=np.float64)
    assert len(delX) == d.Nx and len(delY) == d.Ny and len(delR) == d.Nr
    PI = 3.14159265358979323844
    deg2rad = 2.0 * PI / 360.0
    omega = 2.0 * PI / rotationPeriod
    g = Grid(d)
    z2 = lambda: np.zeros(d.shape2)
    for n in GRID2D:
        g.a[n] = z2()
    xC, yC, xG, yG = z2(), z2(), z2(), z2()
    ox, oy, sx, sy = d.OLx, d.OLy, d.sNx, d.sNy
    PX, PY = d.PX, d.PY
    for bj in range(d.nSy):
        for bi in range(d.nSx):
            iG0 = (d.myPx * d.nSx + bi) * sx
            jG0 = (d.myPy * d.nSy + bj) * sy
            xG0 = float(xgOrigin)
            for i in range(iG0):
                xG0 = xG0 + delX[i]
            for i in range(1, ox + 1):
                xG0 = xG0 - delX[(iG0 - i + ox * d.Nx) % d.Nx]
            yG0 = float(ygOrigin)
            for j in range(jG0):
                yG0 = yG0 + delY[j]
            for j in range(1, oy + 1):
                yG0 = yG0 - delY[(jG0 - j + oy * d.Ny) % d.Ny]
            # delXloc(0-OLx:sNx+OLx): array index a <-> Fortran index a - ox
            dXl = [float(delX[(iG0 + (a - ox) - 1 + ox * d.Nx) % d.Nx]) for a in range(PX + 1)]
            dYl = [float(delY[(jG0 + (a - oy) - 1 + oy * d.Ny) % d.Ny]) for a in range(PY + 1)]
            xGl = [xG0]
            for a in range(1, PX + 1):
                xGl.append(xGl[-1] + dXl[a])
            yGl = [yG0]
            for a in range(1, PY + 1):
                yGl.append(yGl[-1] + dYl[a])
            A = g.a
            for jj in range(PY):
                for ii in range(PX):
                    xG[bj, bi, jj, ii] = xGl[ii]
                    yG[bj, bi, jj, ii] = yGl[jj]
                    xC[bj, bi, jj, ii] = 0.25 * (((xGl[ii] + xGl[ii + 1]) + xGl[ii]) + xGl[ii + 1])
                    yc = 0.25 * (((yGl[jj] + yGl[jj]) + yGl[jj + 1]) + yGl[jj + 1])
                    yC[bj, bi, jj, ii] = yc
                    dlon, dlat = dXl[ii + 1], dYl[jj + 1]
                    A["dxF"][bj, bi, jj, ii] = rSphere * math.cos(yc * deg2rad) * dlon * deg2rad
                    A["dyF"][bj, bi, jj, ii] = rSphere * dlat * deg2rad
                    lat = 0.5 * (yGl[jj] + yGl[jj])
                    dxg = rSphere * math.cos(deg2rad * lat) * dlon * deg2rad
                    A["dxG"][bj, bi, jj, ii] = 0.0 if dxg < 1.0 else dxg
                    A["dyG"][bj, bi, jj, ii] = rSphere * dlat * deg2rad
                    A["rA"][bj, bi, jj, ii] = rSphere * rSphere * dlon * deg2rad * abs(
                        math.sin((lat + dlat) * deg2rad) - math.sin(lat * deg2rad))
                    dlatS = 0.5 * (dYl[jj + 1] + dYl[jj])
                    ras = rSphere * rSphere * dlon * deg2rad * abs(
                        math.sin(yc * deg2rad) - math.sin((yc - dlatS) * deg2rad))
                    A["rAs"][bj, bi, jj, ii] = 0.0 if (abs(yc) > 90.0 or abs(yc - dlatS) > 90.0) else ras
                    latZ = 0.5 * (yGl[jj] + yGl[jj + 1])
                    dlonZ = 0.5 * (dXl[ii + 1] + dXl[ii])
                    raz = rSphere * rSphere * dlonZ * deg2rad * abs(
                        math.sin(latZ * deg2rad) - math.sin((latZ - dlatS) * deg2rad))
                    A["rAz"][bj, bi, jj, ii] = 0.0 if (abs(latZ) > 90.0 or abs(latZ - dlatS) > 90.0) else raz
                    A["tanPhiAtU"][bj, bi, jj, ii] = math.tan(latZ * deg2rad)
                    A["tanPhiAtV"][bj, bi, jj, ii] = math.tan(lat * deg2rad)
                    A["fCori"][bj, bi, jj, ii] = 2.0 * omega * math.sin(yc * deg2rad)
                    A["fCoriG"][bj, bi, jj, ii] = 2.0 * omega * math.sin(yGl[jj] * deg2rad)
            dxF, dyF, dxG, dyG, rA = (A[n][bj, bi] for n in "dxF dyF dxG dyG rA".split())
            A["dxC"][bj, bi][:, 1:] = 0.5 * (dxF[:, 1:] + dxF[:, :-1])
            A["dyC"][bj, bi][1:, :] = 0.5 * (dyF[1:, :] + dyF[:-1, :])
            A["dxV"][bj, bi][1:, 1:] = 0.5 * (dxG[1:, 1:] + dxG[1:, :-1])
            A["dyU"][bj, bi][1:, 1:] = 0.5 * (dyG[1:, 1:] + dyG[:-1, 1:])
            A["rAw"][bj, bi][:, 1:] = 0.5 * (rA[:, 1:] + rA[:, :-1])
    g.a["xC"], g.a["yC"], g.a["xG"], g.a["yG"] = xC, yC, xG, yG
    g.a["Bo_surf"] = np.full(d.shape2, gBaro)
    g.a["recip_Bo"] = np.full(d.shape2, 1.0 / gBaro)
    g.a["cosFacU"] = np.ones((d.nSy, d.nSx, d.PY))
    g.a["cosFacV"] = np.ones((d.nSy, d.nSx, d.PY))
    g.set_recips()
    set_vertical(g, delR)
    return g


def set_vertical(g: Grid, delR) -> None:
    """INI_VERTICAL_GRID for z coordinates: drF = delR, drC(1) = delR(1)/2,
    drC(k) = (delR(k-1)+delR(k))/2, drC(Nr+1) = delR(Nr)/2."""
    delR = np.asarray(delR, dtype=np.float64)
    Nr = len(delR)
    drF = delR.copy()
    drC = np.empty(Nr + 1)
    drC[0] = 0.5 * delR[0]
    for k in range(1, Nr):
        drC[k] = 0.5 * (delR[k - 1] + delR[k])
    drC[Nr] = 0.5 * delR[Nr - 1]
    g.a["drF"], g.a["drC"] = drF, drC
    g.a["recip_drF"], g.a["recip_drC"] = 1.0 / drF, 1.0 / drC
    rF = np.empty(Nr + 1)
    rF[0] = 0.0
    for k in range(Nr):
        rF[k + 1] = rF[k] - delR[k]
    g.a["rF"] = rF
    rC = np.empty(Nr)
    rC[0] = rF[0] - drC[0]
    for k in range(1, Nr):
        rC[k] = rC[k - 1] - drC[k]
    g.a["rC"] = rC


def masks_from_depth(g: Grid, depth_global: np.ndarray, hFacMin=1.0, hFacMinDr=0.0) -> None:
    """INI_DEPTHS + INI_MASKS_ETC for z coordinates with Ro_surf = 0 and a linear
    free surface: hFacC from R_low with the hFacMin / hFacMinDr rule
    (ini_masks_etc.F:100-125), hFacW/S = min of the two neighbours (:238-262),
    EXCH_UV_XYZ_RS (:402), reciprocals and masks (:461-478).
    depth_global is (Ny, Nx), negative below the surface, 0 on land."""
    d = g.d
    ox, oy, sx, sy = d.OLx, d.OLy, d.sNx, d.sNy
    assert depth_global.shape == (d.Ny, d.Nx)
    R_low = np.zeros(d.shape2)
    for bj in range(d.nSy):
        for bi in range(d.nSx):
            iG0 = (d.myPx * d.nSx + bi) * sx
            jG0 = (d.myPy * d.nSy + bj) * sy
            jj = [(jG0 + j - oy) % d.Ny for j in range(d.PY)]
            ii = [(iG0 + i - ox) % d.Nx for i in range(d.PX)]
            R_low[bj, bi] = depth_global[np.ix_(jj, ii)]     # periodic read + exchange
    drF, rdrF, rF = g.a["drF"], g.a["recip_drF"], g.a["rF"]
    hFacC = np.zeros(d.shape3)
    for k in range(d.Nr):
        mn = max(hFacMin, min(hFacMinDr * rdrF[k], 1.0))
        h = (rF[k] - R_low) * rdrF[k]
        h = np.minimum(np.maximum(h, 0.0), 1.0)
        hFacC[:, :, k] = np.where((h < mn * 0.5) | (R_low >= 0.0), 0.0, np.maximum(h, mn))
    hFacW = np.zeros(d.shape3)
    hFacS = np.zeros(d.shape3)
    hFacW[..., :, 1:] = np.minimum(hFacC[..., :, 1:], hFacC[..., :, :-1])
    hFacS[..., 1:, :] = np.minimum(hFacC[..., 1:, :], hFacC[..., :-1, :])
    if d.nPx == 1 and d.nPy == 1:
        exch_xyz(d, hFacW)
        exch_xyz(d, hFacS)
    set_hfac(g, hFacC, hFacW, hFacS)


def set_hfac(g: Grid, hFacC, hFacW, hFacS) -> None:
    for n, h in (("C", hFacC), ("W", hFacW), ("S", hFacS)):
        r = np.zeros_like(h)
        np.divide(1.0, h, out=r, where=h != 0.0)
        g.a["hFac" + n] = np.ascontiguousarray(h)
        g.a["recip_hFac" + n] = r
        g.a["mask" + n] = (h != 0.0).astype(np.float64)


def global_area(g: Grid) -> float:
    """globalArea of INI_MASKS_ETC (ini_masks_etc.F:430-450): per-tile sums of rA*maskInC over the
    interior, i fastest then j, tiles added in (bi, bj) order (GLOBAL_SUM_TILE_RL)."""
    d = g.d
    jj, ii = d.interior()
    tot = 0.0
    for bj in range(d.nSy):
        for bi in range(d.nSx):
            x = (g.a["rA"][bj, bi, jj, ii] * g.a["maskC"][bj, bi, 0, jj, ii]).ravel()
            tot = tot + (float(np.cumsum(x)[-1]) if x.size else 0.0)
    return tot


# ---- curvilinear (cubed-sphere) grids --------------------------------------------------------------
MITGRID_FIELDS = ("xC yC dxF dyF rA xG yG dxV dyU rAz dxC dyC rAw rAs dxG dyG angleCosC angleSinC").split()


def read_mitgrid_faces(prefix: str, nFace: int, nFacets: int = 6):
    """horizGridFile face files `<prefix>.face00N.bin`: 18 records of (nFace+1) x (nFace+1) big-endian
    float64 in the order INI_CURVILINEAR_GRID reads them (model/src/ini_curvilinear_grid.F:318-352).
    Returns a list (one per facet) of dicts name -> (nFace+1, nFace+1) array [j, i]."""
    faces = []
    for f in range(1, nFacets + 1):
        a = np.fromfile(f"{prefix}.face{f:03d}.bin", dtype=">f8")
        assert a.size == len(MITGRID_FIELDS) * (nFace + 1) ** 2, "unexpected mitgrid file size"
        a = a.reshape(len(MITGRID_FIELDS), nFace + 1, nFace + 1).astype(np.float64)
        faces.append({n: a[q] for q, n in enumerate(MITGRID_FIELDS)})
    return faces


def cubed_sphere_grid(d: Dims, topo, faces, delR, gBaro=9.81, rotationPeriod=86164.0, rSphereFac=1.0,
                      Bo_surf=None) -> Grid:
    """usingCurvilinearGrid with pkg/exch2: tiles laid out as one row (nSx = nTiles, nSy = 1, the
    reference's own cs32 SIZE.h), metrics from the face files.  Each tile takes its interior plus the
    (sNx+1, sNy+1) row and column the files carry (MDS_FACEF_READ); cell-centred scalars (xC, yC, rA)
    then get their halos from the scalar exch2 exchange and the C-grid pairs (dxC, dyC), (rAw, rAs),
    (dyG, dxG) from the unsigned vector-pair exchange, as in ini_curvilinear_grid.F:360-370.  The A-grid
    (dxF, dyF), B-grid (dxV, dyU) and corner-point (xG, yG, rAz) exchanges of the reference are NOT done:
    those arrays are valid on 1..sN+1 only: enough for every interior tendency of MOM_FLUXFORM without
    viscosity and of MOM_VECINV (vorticity points 1..sN+1; pinned by solid-body.cs-32x32x1).
    rSphereFac = rSphere / radius_fromHorizGrid rescales lengths and areas (ini_curvilinear_grid.F:375-398);
    Bo_surf overrides gBaro (uniformLin_PhiSurf in p coordinates: 1/rhoConst, ini_linear_phisurf.F:63-73)."""
    from .exch2 import exchange, exchange_uv, halo_gather_map, uv_gather_map
    assert d.nSy == 1 and d.nSx == topo.nTiles and d.OLx == d.OLy
    g = Grid(d)
    for n in GRID2D:
        g.a[n] = np.zeros(d.shape2)
    for n in ("xC", "yC", "xG", "yG", "angleCosC", "angleSinC"):
        g.a[n] = np.zeros(d.shape2)
    ox, oy, sx, sy = d.OLx, d.OLy, d.sNx, d.sNy
    for t in range(topo.nTiles):
        F = faces[int(topo.myFace[t]) - 1]
        bx, by = int(topo.tBasex[t]), int(topo.tBasey[t])
        for n in MITGRID_FIELDS:
            if n in F:      # a reduced set of records is enough for the CG2D operator
                g.a[n][0, t, oy:oy + sy + 1, ox:ox + sx + 1] = F[n][by:by + sy + 1, bx:bx + sx + 1]
    gm = halo_gather_map(topo, ox)
    for n in ("xC", "yC", "rA"):
        exchange(topo, g.a[n][0], ox, gm)
    gmuv = uv_gather_map(topo, ox, False)
    for a, b in (("dxC", "dyC"), ("rAw", "rAs"), ("dyG", "dxG")):
        if g.a[a].any():
            exchange_uv(topo, g.a[a][0], g.a[b][0], ox, False, gmuv)
    if rSphereFac != 1.0:
        for n in "dxC dyC dxG dyG dxF dyF dxV dyU".split():
            g.a[n] = g.a[n] * rSphereFac
        fac2 = rSphereFac * rSphereFac
        for n in "rA rAz rAw rAs".split():
            g.a[n] = g.a[n] * fac2
    PI = 3.14159265358979323844
    omega = 2.0 * PI / rotationPeriod
    import math
    g.a["fCori"] = 2.0 * omega * np.vectorize(math.sin)(g.a["yC"] * (2.0 * PI / 360.0))
    g.a["fCoriG"] = 2.0 * omega * np.vectorize(math.sin)(g.a["yG"] * (2.0 * PI / 360.0))
    bo = gBaro if Bo_surf is None else Bo_surf
    g.a["Bo_surf"] = np.full(d.shape2, bo)
    g.a["recip_Bo"] = np.full(d.shape2, 1.0 / bo)
    g.a["cosFacU"] = np.ones((d.nSy, d.nSx, d.PY))
    g.a["cosFacV"] = np.ones((d.nSy, d.nSx, d.PY))
    g.a["omega"] = omega
    g.set_recips()
    set_vertical(g, delR)
    return g


def cs_corner_flags(topo):
    """Per tile: bit mask of the facet corners it owns (1 SW, 2 SE, 4 NE, 8 NW), the test
    mom_calc_relvort3.F:79-97 / fill_cs_corner_tr_rl.F:62-72 makes from exch2_is{N,S,E,W}edge."""
    tb = topo.tables()
    N, S_, E, W = (np.asarray(tb[n]) for n in ("isNedge", "isSedge", "isEedge", "isWedge"))
    return ((W & S_) * 1 + (E & S_) * 2 + (E & N) * 4 + (W & N) * 8).astype(np.int32)


def cube_masks_from_depth(g: Grid, topo, depth_xstack: np.ndarray, hFacMin=1.0, hFacMinDr=0.0) -> None:
    """INI_DEPTHS + INI_MASKS_ETC on the tile graph: bathymetry in the exch2 global-IO layout
    (W2_mapIO = -1: facets stacked along x, shape (nFace, 6*nFace)), R_low halos by the scalar exch2
    exchange (ini_depths.F: _EXCH_XY_RS(R_low)), hFacC by the hFacMin / hFacMinDr rule, hFacW/S as the
    minimum of the two neighbouring cells wherever both are known (ini_masks_etc.F:238-262)."""
    from .exch2 import exchange
    d = g.d
    ox, oy, sx, sy = d.OLx, d.OLy, d.sNx, d.sNy
    nFace = depth_xstack.shape[0]
    R_low = np.zeros(d.shape2)
    for t in range(topo.nTiles):
        f = int(topo.myFace[t]) - 1
        bx, by = int(topo.tBasex[t]), int(topo.tBasey[t])
        R_low[0, t, oy:oy + sy, ox:ox + sx] = depth_xstack[by:by + sy, f * nFace + bx:f * nFace + bx + sx]
    exchange(topo, R_low[0], ox)
    drF, rdrF, rF = g.a["drF"], g.a["recip_drF"], g.a["rF"]
    hFacC = np.zeros(d.shape3)
    for k in range(d.Nr):
        mn = max(hFacMin, min(hFacMinDr * rdrF[k], 1.0))
        h = (rF[k] - R_low) * rdrF[k]
        h = np.minimum(np.maximum(h, 0.0), 1.0)
        hFacC[:, :, k] = np.where((h < mn * 0.5) | (R_low >= 0.0), 0.0, np.maximum(h, mn))
    hFacW = np.zeros(d.shape3)
    hFacS = np.zeros(d.shape3)
    hFacW[..., :, 1:] = np.minimum(hFacC[..., :, 1:], hFacC[..., :, :-1])
    hFacS[..., 1:, :] = np.minimum(hFacC[..., 1:, :], hFacC[..., :-1, :])
    from .exch2 import exchange_uv
    exchange_uv(topo, hFacW[0], hFacS[0], ox, False)       # EXCH_UV_XYZ_RS(hFacW, hFacS, .FALSE.), ini_masks_etc.F:402
    g.a["R_low"] = R_low
    set_hfac(g, hFacC, hFacW, hFacS)
