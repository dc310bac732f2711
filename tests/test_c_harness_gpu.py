"""The C ABI called from compiled code the way the Fortran shims call it (tests/harness/b200_harness.c: static
COMMON-style arrays, every scalar by reference, LOGICAL as a 4-byte integer, `_`-suffixed symbols resolved by the
linker, no ctypes): CG2D on a tiled grid with bathymetry and one GAD_CALC_RHS level, compared with the CPU oracle.
LOGICAL .TRUE. is tried in both common representations (1: gfortran, -1: Intel)."""
import os
import subprocess

import numpy as np
import pytest

from helpers import make_grid, cg2d_problem
from test_gad_mom_gpu import rand_state

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHAPE = dict(sNx=40, sNy=30, OL=3, nSx=2, nSy=1, Nr=4)


def build_harness(tmp):
    from mitgcm_b200 import build as b
    lib = b.build()
    exe = os.path.join(tmp, "b200_harness")
    defs = [f"-D{k}={v}" for k, v in dict(SNX=SHAPE["sNx"], SNY=SHAPE["sNy"], OLX=SHAPE["OL"], OLY=SHAPE["OL"],
                                          NSX=SHAPE["nSx"], NSY=SHAPE["nSy"], NR=SHAPE["Nr"]).items()]
    subprocess.check_call(["gcc", "-O1", "-std=c99"] + defs + [os.path.join(ROOT, "tests", "harness", "b200_harness.c"),
                          "-I", os.path.join(ROOT, "include"), "-L", os.path.dirname(lib), "-lmitgcm_b200",
                          f"-Wl,-rpath,{os.path.dirname(lib)}", "-o", exe])
    return exe


def test_harness_compiles_and_links_against_every_symbol_it_uses(tmp_path):
    """CPU part: the harness builds and links (the symbols exist with C linkage and the declared prototypes)."""
    exe = build_harness(str(tmp_path))
    assert os.path.exists(exe)


@pytest.mark.gpu
@pytest.mark.parametrize("true_value", [1, -1], ids=["gfortran_true", "intel_true"])
def test_harness_cg2d_and_gad_calc_rhs_match_oracle(tmp_path, true_value):
    from mitgcm_b200 import _lib
    from mitgcm_b200.runtime import GRID_FIELD_NAMES, field_id
    from oracle.pyoracle import Oracle
    E = _lib.ENUMS
    exe = build_harness(str(tmp_path))
    g = make_grid(**SHAPE, seed=31)
    d = g.d
    o, op, b, x = cg2d_problem(g, tol=1e-9)
    recs = []

    def put(*a):
        for v in a:
            recs.append(np.ascontiguousarray(v, dtype=np.float64).ravel())
    # GRID.h records
    grid = [(field_id(n), np.ascontiguousarray(g.a[n], dtype=np.float64)) for n in GRID_FIELD_NAMES if n in g.a]
    for n in ("drF", "drC", "recip_drF", "recip_drC"):
        v = np.zeros(d.Nr + 1)
        v[:len(g.a[n])] = g.a[n]
        grid.append((field_id(n), v))
    put(len(grid))
    for fid, a in grid:
        put(fid, a.size, a)
    put(2, E["MP_CG2DNORM"], op["cg2dNorm"], E["MP_CG2DTOLERANCE_SQ"], op["cg2dTolerance_sq"])
    put(1, E["MI_CG2DNORMALISERHS"], int(op["cg2dNormaliseRHS"]))
    # CG2D
    for n in "aW2d aS2d aC2d pW pS pC".split():
        put(op[n])
    put(b, x, 500, -1)
    # one GAD_CALC_RHS level: c2 advection + diffusion + AB on the tracer
    u, v, w, T = rand_state(g, 5)
    rng = np.random.default_rng(9)
    TAB = T + 0.01 * rng.standard_normal(d.shape3)
    ns = (d.PY, d.PX)
    bi, bj, k = 2, 1, 2
    kUp, kDown = 1 + (k + 1) % 2, 1 + k % 2
    sl = {n: np.zeros(ns) for n in "xA yA maskUp uFld vFld wFld uTrans vTrans rTransKp1".split()}
    rTrans = np.zeros(ns)
    og = Oracle(g, {})
    og.calc_adv_flow(bi, bj, k + 1, u, v, w, *(np.zeros(ns) for _ in range(8)), rTrans, np.zeros(ns))      # rTrans of level k+1
    og.calc_adv_flow(bi, bj, k, u, v, w, sl["xA"], sl["yA"], sl["maskUp"], sl["uFld"], sl["vFld"], sl["wFld"], sl["uTrans"],
                     sl["vTrans"], rTrans, sl["rTransKp1"])
    rT_in = rTrans
    KappaR = 1e-4 * (1 + rng.random(ns))
    kr4, dT = np.zeros(d.Nr), np.full(d.Nr, 1200.0)
    t = (bj - 1, bi - 1)
    fV0 = 1e-3 * rng.standard_normal((2,) + ns)
    gT0 = np.zeros((d.Nr,) + ns)
    flags = [1, 0, 1, 0, 0, 0, 0]
    put(bi, bj, 1, d.sNx, 1, d.sNy, k, max(1, k - 1), kUp, kDown)
    put(sl["xA"], sl["yA"], sl["maskUp"], sl["uFld"], sl["vFld"], sl["wFld"], sl["uTrans"], sl["vTrans"], rT_in, sl["rTransKp1"])
    put(1e3, 0.0, KappaR, kr4, T[t], TAB[t], dT, 1, 2, 2, *flags, fV0, gT0)
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    np.concatenate(recs).tofile(fin)
    r = subprocess.run([exe, fin, fout, str(true_value)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    out = np.fromfile(fout)
    n2, s2 = int(np.prod(d.shape2)), d.PY * d.PX
    pos = 0

    def take(n, shape=None):
        nonlocal pos
        a = out[pos:pos + n]
        pos += n
        return a.reshape(shape) if shape else a
    bg, xg = take(n2, d.shape2), take(n2, d.shape2)
    first, minsq, last, nit, nmin = take(5)
    # oracle
    bo, xo = b.copy(), x.copy()
    ro = o.cg2d(op, bo, xo, 500, -1)
    jj, ii = d.interior()
    assert abs(int(nit) - ro["numIters"]) <= 1
    assert np.array_equal(bg[:, :, jj, ii], bo[:, :, jj, ii])
    assert first == pytest.approx(ro["firstResidual"], rel=1e-13)
    assert last < 1e-9
    assert np.abs(xg[:, :, jj, ii] - xo[:, :, jj, ii]).max() < 1e-6 * np.abs(xo).max()
    fZo, fMo, fVo, gTo = np.zeros(ns), np.zeros(ns), fV0.copy(), gT0.copy()
    og.gad_calc_rhs(bi, bj, 1, d.sNx, 1, d.sNy, k, max(1, k - 1), kUp, kDown, sl["xA"], sl["yA"], sl["maskUp"], sl["uFld"],
                    sl["vFld"], sl["wFld"], sl["uTrans"], sl["vTrans"], rT_in, sl["rTransKp1"], 1e3, 0.0, KappaR, kr4,
                    np.ascontiguousarray(T[t]), np.ascontiguousarray(TAB[t]), dT, 2, 2, *flags[:4], fZo, fMo, fVo, gTo)
    fZg, fMg, fVg, gTg = take(s2, ns), take(s2, ns), take(2 * s2, (2,) + ns), take(d.Nr * s2, (d.Nr,) + ns)
    launches = take(1)[0]
    assert launches >= 2          # the harness ran CUDA kernels, not a fallback
    close = lambda a, c: np.abs(a - c).max() <= 1e-13 * max(np.abs(c).max(), 1e-300)
    assert np.abs(gTo).max() > 0
    assert close(fZg, fZo) and close(fMg, fMo)
    assert close(fVg[kUp - 1], fVo[kUp - 1])
    assert close(gTg[k - 1, :-1, :-1], gTo[k - 1, :-1, :-1])
