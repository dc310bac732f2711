"""Resident model driver: the synthetic doubly-periodic channel of BASELINE.json (config 5) and
its small-grid variants, stepping entirely on the device through the C ABI
(mitgcm_b200_forward_step_).  Set-up mirrors what the Fortran model has done before the first
FORWARD_STEP: grid, masks, INI_CG2D operator, initial fields and forcing in the mirrors."""
from __future__ import annotations

import numpy as np

from . import runtime as rt
from .grid import Dims, Grid, cartesian_grid, masks_from_depth, exch_xyz, global_area

DEFAULTS = dict(deltaTMom=1200.0, deltaTFreeSurf=1200.0, deltaTtracer=1200.0, abEps=0.01,
                viscAhD=400.0, viscAhZ=400.0, viscAr=1e-2, diffKhT=1e3, diffK4T=0.0, diffKrT=1e-5,
                no_slip_sides=1, no_slip_bottom=1, sideDragFactor=2.0, selectBotDragQuadr=-1,
                tempAdvScheme=2, tempStepping=1, cg2dTargetResidual=1e-7, cg2dMaxIters=1000,
                momForcing=1, momDissip_In_AB=1, useSRCGSolver=0,
                # eosType = 'LINEAR' (SURVEY.md section 8(d)); coupled to the momentum equations when buoyancyLinear = 1
                buoyancyLinear=0, gravity=9.81, tAlpha=2e-4, sBeta=0.0, rhoNil=1000.0, rhoConst=1000.0, ivdc_kappa=0.0)

LIB_PARAMS = ("deltaTMom deltaTFreeSurf abEps viscAhD viscAhZ viscA4D viscA4Z sideDragFactor bottomDragLinear "
              "bottomDragQuadratic no_slip_sides no_slip_bottom bottomVisc_pCell selectBotDragQuadr "
              "useBiharmonicVisc implicitViscosity selectCoriScheme rigidLid momAdvection momViscosity "
              "diffKhT diffK4T diffKrT viscAr tempStepping cg2dMaxIters momForcing momDissip_In_AB "
              "implicitDiffusion useSRCGSolver usingSphericalPolarGrid selectMetricTerms recip_rSphere "
              "exactConserv buoyancyLinear doThetaClimRelax gravity tAlpha sBeta rhoNil rhoConst ivdc_kappa "
              "implicSurfPress implicDiv2DFlow rkSign vectorInvariantMomentum useCoriolis useAbsVorticity "
              "selectVortScheme selectKEscheme useJamartMomAdv upwindShear multiDimAdvection "
              "gad_multidim_compressible saltStepping diffKhS diffK4S diffKrS highOrderVorticity upwindVorticity").split()


def channel_state(g: Grid, seed=20261018, tau0=0.1, rhoConst=1000.0, period=None, tTop=20.0, tBot=2.0, tNoise=0.1):
    """Initial state and forcing of the synthetic channel: smooth multi-mode
    flow of 0.1 m/s + noise, theta = tRef(k) + 0.1 N(0,1), eta = 0.1 sin cos, zonal wind stress
    -tau0 cos(2 pi y / Ly).  Generated per global index, so any tiling sees the same field.
    period = (cells in x, cells in y) of the smooth part (default: the global domain); the weak-scaled
    bench keeps it at one per-GPU block and repeats the noise too, so the domain at N ranks is the exact
    periodic tiling of the one-block problem: same flow, same CG2D iteration counts at every rank count."""
    d = g.d
    rng = np.random.default_rng(seed)
    Nx, Ny, Nr = d.Nx, d.Ny, d.Nr
    perx, pery = period or (Nx, Ny)
    X = (np.arange(Nx) + 0.5) / perx
    Y = (np.arange(Ny) + 0.5) / pery
    YY, XX = np.meshgrid(Y, X, indexing="ij")
    psi_u = 0.1 * np.sin(2 * np.pi * XX) * np.cos(4 * np.pi * YY) + 0.03 * np.cos(6 * np.pi * XX) * np.sin(2 * np.pi * YY)
    psi_v = 0.1 * np.cos(2 * np.pi * XX) * np.sin(4 * np.pi * YY) - 0.03 * np.sin(6 * np.pi * XX) * np.cos(2 * np.pi * YY)
    def noise():
        if period and Nx % perx == 0 and Ny % pery == 0:      # the same noise on every block: an exact periodic tiling
            return np.tile(rng.standard_normal((Nr, pery, perx)), (1, Ny // pery, Nx // perx))
        return rng.standard_normal((Nr, Ny, Nx))
    glob = {
        "uVel": psi_u[None] * np.linspace(1.0, 0.2, Nr)[:, None, None] + 1e-3 * noise(),
        "vVel": psi_v[None] * np.linspace(1.0, 0.2, Nr)[:, None, None] + 1e-3 * noise(),
        "theta": np.linspace(tTop, tBot, Nr)[:, None, None] + tNoise * noise(),
    }
    eta = 0.1 * np.sin(2 * np.pi * XX) * np.cos(2 * np.pi * YY)
    tau = -tau0 * np.cos(2 * np.pi * YY)

    def tile(a3):
        out = np.zeros((d.nSy, d.nSx) + a3.shape[:-2] + (d.PY, d.PX))
        for bj in range(d.nSy):
            for bi in range(d.nSx):
                out[bj, bi, ..., d.OLy:d.OLy + d.sNy, d.OLx:d.OLx + d.sNx] = \
                    a3[..., bj * d.sNy:(bj + 1) * d.sNy, bi * d.sNx:(bi + 1) * d.sNx]
        return exch_xyz(d, out)
    s = {k: tile(v) for k, v in glob.items()}
    s["uVel"] *= g.maskW
    s["vVel"] *= g.maskS
    s["theta"] *= g.maskC
    s["wVel"] = np.zeros(d.shape3)
    s["etaN"] = tile(eta) * g.maskC[:, :, 0]
    s["surfForcU"] = tile(tau) * (1.0 / rhoConst)
    s["surfForcV"] = np.zeros(d.shape2)
    s["tRef"] = np.linspace(tTop, tBot, Nr)       # reference profile of the linear equation of state
    s["sRef"] = np.zeros(Nr)
    return s


BENCH_FSIN_AMP = 6.5e-5      # bench workload: f = 1e-4 + 6.5e-5 sin(2 pi y / block length): max f dt = 0.198


def make_channel(sNx, sNy, Nr, nSx=1, nSy=1, OL=2, dx=20e3, dz=100.0, land_frac=0.0, seed=20261018, block=None,
                 **params):
    """block = (cells in x, cells in y): the bench workload -- Coriolis parameter and smooth initial flow periodic
    over one block (the per-GPU domain of the weak-scaled run); default: beta plane over the whole domain."""
    d = Dims(sNx=sNx, sNy=sNy, OLx=OL, OLy=OL, nSx=nSx, nSy=nSy, Nr=Nr)
    g = cartesian_grid(d, [dx] * d.Nx, [dx] * d.Ny, [dz] * Nr, f0=1e-4, beta=1e-11, gBaro=9.81,
                       fsin=(BENCH_FSIN_AMP, block[1] * dx) if block else None)
    rng = np.random.default_rng(seed + 1)
    depth = -dz * Nr * np.ones((d.Ny, d.Nx))
    if land_frac > 0:
        for _ in range(max(1, int(land_frac * 40))):      # rectangular islands
            j0, i0 = rng.integers(0, d.Ny), rng.integers(0, d.Nx)
            h, w = rng.integers(1, max(2, d.Ny // 6)), rng.integers(1, max(2, d.Nx // 6))
            depth[j0:j0 + h, i0:i0 + w] = 0.0
        depth *= 0.4 + 0.6 * rng.random(depth.shape)       # partial cells
    masks_from_depth(g, depth, hFacMin=0.2 if land_frac > 0 else 1.0)
    P = dict(DEFAULTS)
    P.update(params)
    P["globalArea"] = global_area(g)
    strat = {k: P.pop(k) for k in ("tTop", "tBot", "tNoise") if k in P}
    return g, P, channel_state(g, seed, period=block, **strat)


def rank_tiles(g: Grid, arrays: list, rank: int, world: int):
    """This rank's share of a tile graph held as (1, nTiles, ...) arrays: consecutive blocks of nTiles / world tile ids
    per rank (W2_MAP_PROCS, w2_map_procs.F:60-91).  Returns (local Grid, the dicts of `arrays` cut the same way)."""
    d = g.d
    assert d.nSy == 1 and d.nSx % world == 0 and d.nPx == d.nPy == 1
    n = d.nSx // world
    lo, hi = rank * n, (rank + 1) * n

    def cut(v):
        if isinstance(v, np.ndarray) and v.ndim >= 4 and v.shape[:2] == (1, d.nSx):
            return v[:, lo:hi].copy()
        return v
    dl = Dims(sNx=d.sNx, sNy=d.sNy, OLx=d.OLx, OLy=d.OLy, nSx=n, nSy=1, Nr=d.Nr, nPx=world, nPy=1, myPx=rank, myPy=0)
    return Grid(dl, {k: cut(v) for k, v in g.a.items()}), [{k: cut(v) for k, v in a.items()} for a in arrays]


class Model:
    """Device-resident model: Model(grid, params, state).step() == one FORWARD_STEP.
    ranks = (rank, world) with an exch2 topology: the tile graph is spread over `world` GPUs (one process each, NCCL
    process group initialised); g, state and op describe the WHOLE graph and every rank keeps its share."""

    def __init__(self, g: Grid, P: dict, state: dict, op: dict, device=-1, topo=None, ranks=None):
        self.dist = ranks is not None and ranks[1] > 1
        if self.dist:
            assert topo is not None, "several ranks with several tiles each need the exch2 tile graph"
            g, (state, op) = rank_tiles(g, [state, op], *ranks)
        self.g, self.d, self.P = g, g.d, P
        rt.init(g.d, device)
        rt.set_grid(g)
        if self.dist:
            from . import distributed
            from .exch2 import set_topology, tile_proc
            distributed.setup(g.d)
            set_topology(topo, tileProc=tile_proc(topo.nTiles, ranks[1]))
        elif topo is not None:        # pkg/exch2 tile graph (cubed sphere): all exchanges follow it
            from .exch2 import set_topology
            set_topology(topo)
        rt.set_params(**{k: P[k] for k in LIB_PARAMS if k in P})
        rt.set_params(deltaTtracer=P.get("deltaTtracer", P["deltaTMom"]), tempAdvScheme=P.get("tempAdvScheme", 2),
                      tempVertAdvScheme=P.get("tempAdvScheme", 2), saltAdvScheme=P.get("saltAdvScheme", 2),
                      saltVertAdvScheme=P.get("saltAdvScheme", 2), nIter0=0)
        rt.set_cg2d_operator(op)
        for n, fid in (("uVel", "uVel"), ("vVel", "vVel"), ("wVel", "wVel"), ("theta", "theta"), ("etaN", "etaN"),
                       ("surfForcU", "surfForcU"), ("surfForcV", "surfForcV")):
            rt.set_field(fid, np.ascontiguousarray(state[n]))
        # optional physics of the wider configurations (include/mitgcm_b200.h, forward step)
        for n in ("salt", "SST", "lambdaThetaClimRelax", "etaH"):
            if n in state:
                rt.set_field(n, np.ascontiguousarray(state[n]))
        for n in ("tRef", "sRef", "rF", "rC"):
            src = state.get(n, g.a.get(n))
            if src is not None:
                v = np.zeros(g.d.Nr + 1)
                v[:len(src)] = src
                rt.set_field(n, v)
        rt.fill_field("kappaRU", P.get("viscAr", 0.0))
        rt.fill_field("kappaRV", P.get("viscAr", 0.0))
        rt.fill_field("kappaRT", P.get("diffKrT", 0.0))
        for n in ("gU", "gV", "guNm1", "gvNm1", "gtNm1", "theta2", "cg2d_b", "cg2d_x"):
            rt.fill_field(n, 0.0)
        if P.get("saltStepping"):
            rt.fill_field("kappaRS", P.get("diffKrS", 0.0))
            rt.fill_field("gsNm1", 0.0)
            rt.fill_field("salt2", 0.0)
        self.it = 0

    def step(self):
        r = rt.forward_step(self.it)
        self.it += 1
        return r

    def get(self, name):
        d = self.d
        shape = d.shape2 if rt.field_id(name) < 100 else d.shape3
        return rt.get_field(name, np.zeros(shape))

    def close(self):
        if self.dist:
            from . import distributed
            distributed.teardown()
        rt.finalize()


def ini_cg2d(g: Grid, P: dict, hfac_flat: float | None = None, exch=None, exch_uv=None) -> dict:
    """INI_CG2D (model/src/ini_cg2d.F:76-234) vectorised over the horizontal, level loop kept in
    order so the sums are bit-identical to the Fortran.  hfac_flat: use hFacW = hFacS = const
    instead of the 3-D arrays (flat-bottom set-ups whose masks live only on the device).
    This is model set-up (it runs once, or once per step under NLFS), not the hot path.
    exch(d, a): halo update used for pC (EXCH_XY_RS); exch_uv(d, a, b): unsigned vector-pair update used for
    (aW2d, aS2d) and (pW, pS) (EXCH_UV_XY_RS(.., .FALSE.), ini_cg2d.F:133, 233).  Defaults: the single-process
    periodic exchange, where a vector-pair exchange is two scalar ones."""
    d = g.d
    exch_xyz = exch or globals()["exch_xyz"]
    if exch_uv is None:
        def exch_uv(dd, a, b):
            exch_xyz(dd, a)
            exch_xyz(dd, b)
    jj, ii = d.interior()
    I = (slice(None), slice(None), jj, ii)
    aW, aS = np.zeros(d.shape2), np.zeros(d.shape2)
    fac = P.get("implicSurfPress", 1.0) * P.get("implicDiv2DFlow", 1.0)
    for k in range(d.Nr):
        hW = hfac_flat if hfac_flat is not None else g.hFacW[:, :, k][I]
        hS = hfac_flat if hfac_flat is not None else g.hFacS[:, :, k][I]
        aW[I] = aW[I] + fac * (g.dyG[I] * g.drF[k] * hW) * g.recip_dxC[I]
        aS[I] = aS[I] + fac * (g.dxG[I] * g.drF[k] * hS) * g.recip_dyC[I]
    myNorm = max(np.abs(aW[I]).max(), np.abs(aS[I]).max())
    myNorm = 1.0 / myNorm if myNorm != 0 else 1.0
    aW[I] = aW[I] * myNorm
    aS[I] = aS[I] * myNorm
    exch_uv(d, aW, aS)
    normalise = P.get("cg2dTargetResWunit", -1.0) <= 0.0
    tol = P.get("cg2dTargetResidual", 1e-7) if normalise else \
        myNorm * P["cg2dTargetResWunit"] * P["globalArea"] / P["deltaTMom"]
    oy, ox = d.OLy, d.OLx
    C0 = (slice(None), slice(None), slice(oy - 1, oy + d.sNy), slice(ox - 1, ox + d.sNx))       # 0..sN
    CE = (slice(None), slice(None), slice(oy - 1, oy + d.sNy), slice(ox, ox + d.sNx + 1))       # i+1
    CN = (slice(None), slice(None), slice(oy, oy + d.sNy + 1), slice(ox - 1, ox + d.sNx))       # j+1
    aC = np.zeros(d.shape2)
    aC[C0] = -((((aW[C0] + aW[CE]) + aS[C0]) + aS[CN]) +
               P.get("freeSurfFac", 1.0) * myNorm * g.recip_Bo[C0] * g.rA[C0] / P["deltaTMom"] / P["deltaTFreeSurf"])
    W = (slice(None), slice(None), jj, slice(ox - 1, ox + d.sNx - 1))
    S = (slice(None), slice(None), slice(oy - 1, oy + d.sNy - 1), ii)
    pC, pW, pS = np.zeros(d.shape2), np.zeros(d.shape2), np.zeros(d.shape2)
    off = P.get("cg2dpcOffDFac", 0.51)
    with np.errstate(divide="ignore", invalid="ignore"):
        pC[I] = np.where(aC[I] == 0.0, 1.0, 1.0 / aC[I])
        tw = off * (aC[W] + aC[I])
        pW[I] = np.where(aC[I] + aC[W] == 0.0, 0.0, -aW[I] / (tw * tw))
        ts = off * (aC[S] + aC[I])
        pS[I] = np.where(aC[I] + aC[S] == 0.0, 0.0, -aS[I] / (ts * ts))
    exch_xyz(d, pC)
    exch_uv(d, pW, pS)
    return dict(aW2d=aW, aS2d=aS, aC2d=aC, pW=pW, pS=pS, pC=pC, cg2dNorm=myNorm, cg2dTolerance_sq=tol * tol,
                cg2dNormaliseRHS=normalise)


def ini_cg2d_tilegraph(g: Grid, P: dict, topo) -> dict:
    """INI_CG2D on a pkg/exch2 tile graph (cubed sphere): the same arithmetic as ini_cg2d with the
    reference's halo updates done on the tile graph -- EXCH_UV_XY_RS(aW2d, aS2d, .FALSE.) and
    EXCH_UV_XY_RS(pW, pS, .FALSE.) as unsigned vector-pair exchanges (u/v swap across rotated facet
    edges), EXCH_XY_RS(pC) as a scalar one (ini_cg2d.F:133, 231-233)."""
    from .exch2 import exchange, exchange_uv, halo_gather_map, uv_gather_map
    ol = g.d.OLx
    gm, gmuv = halo_gather_map(topo, ol), uv_gather_map(topo, ol, False)
    assert g.d.nSy == 1

    def ex(d, a):
        exchange(topo, a[0], ol, gm)
        return a

    def ex_uv(d, a, b):
        exchange_uv(topo, a[0], b[0], ol, False, gmuv)
    return ini_cg2d(g, P, exch=ex, exch_uv=ex_uv)
