"""GAD_ADVECTION (multi-dimensional advection) on the GPU through the C ABI
with the reference argument list, against the oracle (pinned to verification/advect_xy) and against that
experiment's golden output with the CUDA kernels in the loop."""
import json
import os

import numpy as np
import pytest

from helpers import make_grid
from oracle.pyoracle import Oracle

pytestmark = pytest.mark.gpu


@pytest.fixture()
def rt():
    from mitgcm_b200 import runtime
    yield runtime
    runtime.finalize()


@pytest.mark.parametrize("compressible", [0, 1], ids=["default", "compressible"])
@pytest.mark.parametrize("scheme", [33, 77, 30, 20, 1, 7])
def test_gad_advection_matches_oracle(rt, scheme, compressible):
    """Random flow (with vertical velocity) on a land-masked, partial-cell grid, 2 x 2 tiles, 6 levels:
    the whole halo'd slab of gTracer, every level, <= 1e-13 relative (expected bit-identical)."""
    g = make_grid(sNx=24, sNy=14, OL=4, nSx=2, nSy=2, Nr=6, seed=17)
    d = g.d
    o = Oracle(g, {})
    rng = np.random.default_rng(3)
    u = 0.3 * rng.standard_normal(d.shape3) * g.maskW
    v = 0.3 * rng.standard_normal(d.shape3) * g.maskS
    w = 2e-4 * rng.standard_normal(d.shape3) * g.maskC
    T = (10.0 + rng.standard_normal(d.shape3)) * g.maskC
    dT = np.full(d.Nr, 600.0)
    rt.init(d)
    rt.set_grid(g)
    rt.set_params(gad_multidim_compressible=compressible)
    for bj in range(1, d.nSy + 1):
        for bi in range(1, d.nSx + 1):
            t = (bj - 1, bi - 1)
            uf, vf, wf = (np.ascontiguousarray(a[t]) for a in (u, v, w))
            go, gg = np.zeros((d.Nr, d.PY, d.PX)), np.zeros((d.Nr, d.PY, d.PX))
            assert o.gad_advection(bi, bj, scheme, scheme, 0, compressible, dT, u, v, w, T, go) == 0
            rt.gad_advection(0, scheme, scheme, 1, dT, uf, vf, wf, T, gg, bi, bj)
            wet = g.hFacC[t] > 0          # dry cells of the compressible form divide by a unit volume: compare all
            scale = np.abs(go).max()
            assert scale > 0
            assert np.abs(gg - go).max() <= 1e-13 * scale, (bi, bj)
            assert np.isfinite(gg[wet]).all()
    if not compressible:          # X+Y passes only (vertical part left to GAD_IMPLICIT_R)
        go, gg = np.zeros((d.Nr, d.PY, d.PX)), np.zeros((d.Nr, d.PY, d.PX))
        o.gad_advection(1, 1, scheme, scheme, 1, 0, dT, u, v, w, T, go)
        rt.gad_advection(1, scheme, scheme, 1, dT, np.ascontiguousarray(u[0, 0]), np.ascontiguousarray(v[0, 0]),
                         np.ascontiguousarray(w[0, 0]), T, gg, 1, 1)
        assert np.abs(gg - go).max() <= 1e-13 * np.abs(go).max()


def test_gad_advection_rejects_linear_schemes(rt):
    g = make_grid(sNx=8, sNy=8, OL=3, Nr=2, seed=1)
    d = g.d
    rt.init(d)
    rt.set_grid(g)
    z = np.zeros((d.Nr, d.PY, d.PX))
    with pytest.raises(rt.B200Error):      # centred 2nd order is not a multi-dim scheme (gad_advection.F:447)
        rt.gad_advection(0, 2, 2, 1, np.ones(d.Nr), z, z, z, np.zeros(d.shape3), z.copy(), 1, 1)


def test_advect_xy_golden_with_the_cuda_kernel(rt):
    """verification/advect_xy (salt, scheme 33) with gad_advection_b200_ in the loop: every printed digit
    of %MON dynstat_salt_* at steps 0, 16, ..., 80."""
    from oracle import advect_xy as ax
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "advect_xy.json")))
    d, g, _ = ax.setup()
    rt.init(d)
    rt.set_grid(g)
    rt.set_params(gad_multidim_compressible=1)

    def cuda_advect(bi, bj, scheme, vscheme, implicitAdvection, compressible, dT, u, v, w, tracer, gS):
        t = (bj - 1, bi - 1)
        rt.gad_advection(implicitAdvection, scheme, vscheme, 2, dT, np.ascontiguousarray(u[t]), np.ascontiguousarray(v[t]),
                         np.ascontiguousarray(w[t]), tracer, gS, bi, bj)
        return 0
    out = ax.run(80, advect=cuda_advect)
    for r, mx, mn, me, sd in zip(out, gold["dynstat_salt_max"], gold["dynstat_salt_min"], gold["dynstat_salt_mean"],
                                 gold["dynstat_salt_sd"]):
        assert (f"{r['max']:.13E}", f"{r['min']:.13E}", f"{r['mean']:.13E}", f"{r['sd']:.13E}") == (mx, mn, me, sd)


class _CubeEngine:
    """gad_advection with the Oracle's signature, routed to gad_advection_b200_ (the facet number and the facet
    edges come from the topology handed to the library, as the reference reads W2_EXCH2_TOPOLOGY.h)."""

    def __init__(self, rt):
        self.rt = rt

    def setup(self, g, params, op, topo):
        from mitgcm_b200.exch2 import set_topology
        self.rt.init(g.d)
        self.rt.set_grid(g)
        set_topology(topo)
        self.rt.set_params(rkSign=params["rkSign"])

    def gad_advection(self, bi, bj, scheme, vscheme, implicitAdvection, compressible, dT, u, v, w, tracer, gT, nCFace=0, edges=0):
        t = (bj - 1, bi - 1)
        self.rt.set_params(gad_multidim_compressible=int(compressible))
        self.rt.gad_advection(implicitAdvection, scheme, vscheme, 1, dT, np.ascontiguousarray(u[t]), np.ascontiguousarray(v[t]),
                              np.ascontiguousarray(w[t]), tracer, gT, bi, bj)
        return 0


@pytest.mark.parametrize("compressible", [0, 1], ids=["default", "compressible"])
@pytest.mark.parametrize("scheme", [33, 7, 1])
def test_gad_advection_on_the_cubed_sphere_matches_oracle(rt, scheme, compressible):
    """The three facet-dependent passes on 24 tiles of 16x16 (cs32 grid files, random land): interior tiles, edge
    tiles and corner tiles of every facet; whole halo'd slab of the tendency <= 1e-13."""
    from oracle import advect_cs as acs
    from mitgcm_b200.grid import cube_masks_from_depth
    T, d, g, _, _, _ = acs.setup(16, 16)
    rng = np.random.default_rng(11)
    depth = np.where(rng.random((32, 192)) < 0.1, 0.0, -1.0e5)
    g.a["rF"] = np.array([0.0, -1.0e5])
    cube_masks_from_depth(g, T, depth, hFacMin=1.0, hFacMinDr=0.0)
    o = Oracle(g, dict(rkSign=-1.0))
    eng = _CubeEngine(rt)
    eng.setup(g, o.params, None, T)
    edges = acs.tile_edges(T)
    u = 0.5 * rng.standard_normal(d.shape3) * g.maskW
    v = 0.5 * rng.standard_normal(d.shape3) * g.maskS
    w = np.zeros(d.shape3)
    tr = (1.0 + rng.random(d.shape3)) * g.maskC
    dT = np.full(d.Nr, 2700.0)
    seen = set()
    for bi in range(1, d.nSx + 1):
        a, b = np.zeros((d.Nr, d.PY, d.PX)), np.zeros((d.Nr, d.PY, d.PX))
        assert o.gad_advection(bi, 1, scheme, scheme, 0, compressible, dT, u, v, w, tr, a, int(T.myFace[bi - 1]), int(edges[bi - 1])) == 0
        eng.gad_advection(bi, 1, scheme, scheme, 0, compressible, dT, u, v, w, tr, b)
        fin = np.isfinite(a)
        assert np.array_equal(fin, np.isfinite(b)), bi
        scale = max(np.abs(a[fin]).max(), 1e-300)
        assert np.abs(a[fin] - b[fin]).max() <= 1e-13 * scale, (bi, int(T.myFace[bi - 1]), int(edges[bi - 1]))
        seen.add((int(T.myFace[bi - 1]), int(edges[bi - 1])))
    assert len(seen) == 24


def test_advect_cs_golden_with_the_cuda_kernel(rt):
    """verification/advect_cs (theta, scheme 33 on the cs32 cube) with gad_advection_b200_ in the loop: every
    printed digit of %MON dynstat_theta_* at steps 0, 8, ..., 192."""
    from oracle import advect_cs as acs
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "advect_cs.json")))
    out = acs.run(192, engine=_CubeEngine(rt))
    assert len(out) == 25
    for r, mx, mn, me, sd in zip(out, gold["dynstat_theta_max"], gold["dynstat_theta_min"], gold["dynstat_theta_mean"],
                                 gold["dynstat_theta_sd"]):
        assert (f"{r['max']:.13E}", f"{r['min']:.13E}", f"{r['mean']:.13E}", f"{r['sd']:.13E}") == (mx, mn, me, sd)
