"""Pins the CPU oracle against the reference's own golden output
(verification/tutorial_barotropic_gyre/results/output.txt, extracted to
tests/golden/tutorial_barotropic_gyre.json by tests/golden/extract_golden.py).

Pass rule = the reference's own (verification/testreport:956-987): number of
matching significant digits, here required on every printed digit (the goldens
print 15 digits for cg2d_*, 14 for %MON)."""
import json
import os

import numpy as np
import pytest

from oracle import barotropic_gyre as bg

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "tutorial_barotropic_gyre.json")))


@pytest.fixture(scope="module")
def run10():
    return bg.run(10)


def fmt(v, nd):
    return f"{v:.{nd}E}"


def test_inputs_match_reference_files():
    ref = "/root/reference/verification/tutorial_barotropic_gyre/input"
    if not os.path.isdir(ref):
        pytest.skip("reference tree not present (GPU box)")
    h, t = bg.gen_inputs()
    assert h.tobytes() == open(os.path.join(ref, "bathy.bin"), "rb").read()
    assert t.tobytes() == open(os.path.join(ref, "windx_cosy.bin"), "rb").read()


def test_cg2d_norm(run10):
    norm, _ = run10
    assert fmt(norm, 16) == GOLD["cg2dNorm"]          # ini_cg2d.F:161 '(1PE23.16)'


def test_cg2d_lines_all_steps(run10):
    _, out = run10
    assert [r["numIters"] for r in out] == GOLD["cg2d_iters"]
    for r, ir, lr, (sr, rm) in zip(out, GOLD["cg2d_init_res"], GOLD["cg2d_last_res"], GOLD["sumRHS_rhsMax"]):
        assert fmt(r["firstResidual"], 14) == ir       # solve_for_pressure.F:337 '(1PE23.14)'
        assert fmt(r["lastResidual"], 14) == lr
        assert fmt(r["rhsMax"], 14) == rm              # cg2d.F:199 '(1P2E22.14)'
        assert abs(r["sumRHS"] - float(sr)) <= 1e-14 * max(1.0, abs(float(sr))) or fmt(r["sumRHS"], 14) == sr


@pytest.mark.parametrize("fld", ["eta", "uvel", "vvel"])
@pytest.mark.parametrize("st", ["max", "min", "sd"])
def test_monitor_dynstats(run10, fld, st):
    _, out = run10
    gold = GOLD[f"dynstat_{fld}_{st}"][1:]             # entry 0 is the initial state
    for r, gv in zip(out, gold):
        a, b = r[fld][st], float(gv)
        # testreport's tr_cmpnum: digits = -log10(|a-b| / (0.5(|a|+|b|))); the monitor prints 14
        assert a == pytest.approx(b, rel=5e-13, abs=1e-30), (fld, st)


def test_multi_tile_decomposition_reproduces_golden():
    """SIZE.h_mpi-style 2x2 tiling of the same domain (31x31 tiles): with the ordered
    tile sum the iteration counts stay identical and init_res agrees to >= 12 digits
    (summation order inside GLOBAL_SUM_TILE changes with the tiling)."""
    _, out = bg.run(4, nSx=2, nSy=2)
    assert [r["numIters"] for r in out] == GOLD["cg2d_iters"][:4]
    for r, ir in zip(out, GOLD["cg2d_init_res"]):
        assert r["firstResidual"] == pytest.approx(float(ir), rel=1e-11)


def test_cg2d_sr_converges_to_same_solution():
    """CG2D_SR (cg2d_sr.F) is a different recurrence: same tolerance, close counts."""
    _, a = bg.run(3)
    _, b = bg.run(3, sr=True)
    for ra, rb in zip(a, b):
        assert abs(ra["numIters"] - rb["numIters"]) <= 3
        assert rb["lastResidual"] < 1e-7
        assert rb["eta"]["max"] == pytest.approx(ra["eta"]["max"], rel=1e-6)


# ---------------------------------------------------------------------------------------
# Config 2: verification/tutorial_baroclinic_gyre (62x62x15 spherical polar, 2x2 tiles).
# Pins GAD_CALC_RHS (c2 advection, Laplacian diffusion, implicit vertical diffusion with
# IVDC) and the Nr > 1 / spherical branches of MOM_FLUXFORM (vertical advection and
# viscosity, metric terms, no-slip sides) against results/output.txt.
# ---------------------------------------------------------------------------------------
from oracle import baroclinic_gyre as bc

GOLD2 = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "tutorial_baroclinic_gyre.json")))


@pytest.fixture(scope="module")
def bc10():
    return bc.run(10)


def test_config2_inputs_match_reference_files():
    ref = "/root/reference/verification/tutorial_baroclinic_gyre/input"
    if not os.path.isdir(ref):
        pytest.skip("reference tree not present (GPU box)")
    h, t, s = bc.gen_inputs()
    assert h.tobytes() == open(os.path.join(ref, "bathy.bin"), "rb").read()
    assert t.tobytes() == open(os.path.join(ref, "windx_cosy.bin"), "rb").read()
    assert s.tobytes() == open(os.path.join(ref, "SST_relax.bin"), "rb").read()


def test_config2_cg2d_norm_spherical_grid(bc10):
    norm, _ = bc10
    assert fmt(norm, 16) == GOLD2["cg2dNorm"]         # 1.4846576448792053E-04


def test_config2_cg2d_lines_all_steps(bc10):
    _, out = bc10
    assert [r["numIters"] for r in out] == GOLD2["cg2d_iters"]
    for r, ir, lr, (sr, rm) in zip(out, GOLD2["cg2d_init_res"], GOLD2["cg2d_last_res"], GOLD2["sumRHS_rhsMax"]):
        assert fmt(r["firstResidual"], 14) == ir
        assert fmt(r["lastResidual"], 14) == lr
        assert fmt(r["rhsMax"], 14) == rm


@pytest.mark.parametrize("fld", ["eta", "uvel", "vvel", "wvel", "theta"])
@pytest.mark.parametrize("st", ["max", "min", "mean", "sd"])
def test_config2_monitor_dynstats(bc10, fld, st):
    """Every printed digit of %MON dynstat_* for the 10 steps (the means of eta, vvel, wvel are
    round-off residues of order 1e-20 and are compared like the others: they only match when the
    summation order is the reference's)."""
    _, out = bc10
    gold = GOLD2[f"dynstat_{fld}_{st}"][1:]
    for r, gv in zip(out, gold):
        assert fmt(r[fld][st], 13) == gv.replace("-0.0000000000000E+00", "0.0000000000000E+00") or \
            r[fld][st] == pytest.approx(float(gv), rel=5e-13, abs=1e-30), (fld, st)


# ---------------------------------------------------------------------------------------
# Config 3: verification/global_ocean.90x40x15 -- operator level: the CG2D
# operator built by INI_CG2D from the experiment's real bathymetry (partial cells,
# hFacMin = 0.05, hFacMinDr = 50 m) on the 4-degree spherical-polar grid, 9x4 tiles of 10x10.
# ---------------------------------------------------------------------------------------
def config3_grid():
    from mitgcm_b200.grid import Dims, spherical_polar_grid, masks_from_depth
    d = Dims(sNx=10, sNy=10, OLx=3, OLy=3, nSx=9, nSy=4, Nr=15)
    delR = [50., 70., 100., 140., 190., 240., 290., 340., 390., 440., 490., 540., 590., 640., 690.]
    g = spherical_polar_grid(d, [4.0] * 90, [4.0] * 40, delR, xgOrigin=0.0, ygOrigin=-80.0)
    f = os.path.join(os.path.dirname(__file__), "golden", "inputs", "global_oce_latlon_bathymetry.bin")
    bathy = np.fromfile(f, ">f4").reshape(40, 90).astype(np.float64)
    masks_from_depth(g, bathy, hFacMin=0.05, hFacMinDr=50.0)
    return g


def test_config3_cg2d_norm_and_area_from_real_bathymetry():
    from mitgcm_b200.grid import global_area
    from oracle.pyoracle import Oracle
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "global_ocean.90x40x15.json")))
    g = config3_grid()
    area = global_area(g)
    assert area == pytest.approx(3.450614146649756E+14, rel=1e-15)   # results/output.txt:1897
    o = Oracle(g, dict(deltaTMom=1800.0, deltaTFreeSurf=86400.0, cg2dTargetResidual=1e-13, globalArea=area))
    op = o.ini_cg2d()
    assert fmt(op["cg2dNorm"], 16) == gold["cg2dNorm"]               # 6.5682677425711703E-05, :766


def test_config3_operator_solve_converges_like_the_reference():
    """CG2D on the real-bathymetry operator with the experiment's tolerance (1e-13): the golden run
    needs 122-128 iterations per step (results/output.txt:2208-2484); a smooth random right-hand
    side on the same operator must converge in a comparable count."""
    from mitgcm_b200.grid import global_area
    from oracle.pyoracle import Oracle
    g = config3_grid()
    d = g.d
    o = Oracle(g, dict(deltaTMom=1800.0, deltaTFreeSurf=86400.0, cg2dTargetResidual=1e-13, globalArea=global_area(g)))
    op = o.ini_cg2d()
    rng = np.random.default_rng(3)
    jj, ii = d.interior()
    b = np.zeros(d.shape2)
    b[:, :, jj, ii] = rng.standard_normal((d.nSy, d.nSx, d.sNy, d.sNx))
    b *= g.maskC[:, :, 0] * g.rA / 1800.0
    x = np.zeros(d.shape2)
    r = o.cg2d(op, b, x, 500, -1)
    assert 80 <= r["numIters"] <= 200 and r["lastResidual"] < 1e-13


# ---------------------------------------------------------------------------------------
# verification/adjustment.cs-32x32x1: barotropic adjustment on the cs32 cubed sphere (48 tiles,
# pkg/exch2).  Pins the exch2 exchanges (scalar, signed and unsigned vector), CG2D on the cube,
# the Crank-Nicolson free surface (implicSurfPress = implicDiv2DFlow = 0.5, exactConserv) and
# MOM_FLUXFORM's Coriolis term on a curvilinear grid against results/output.txt, 24 steps.
# ---------------------------------------------------------------------------------------
from oracle import adjustment_cs as ac

GOLDA = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "adjustment.cs-32x32x1.json")))


@pytest.fixture(scope="module")
def ac24():
    return ac.run(24)


def test_adjustment_cs_cg2d_lines_all_steps(ac24):
    """Iteration counts identical for all 24 steps; cg2d_init_res, Sum(rhs), rhsMax to >= 13 digits
    (all 15 printed digits except a last-digit flip at step 21; the reference's own pass rule is 10)."""
    norm, out = ac24
    assert fmt(norm, 16) == GOLDA["cg2dNorm"]          # 2.9024774185495597E-03
    assert [r["numIters"] for r in out] == GOLDA["cg2d_iters"]
    exact = 0
    for r, ir, (sr, rm) in zip(out, GOLDA["cg2d_init_res"], GOLDA["sumRHS_rhsMax"]):
        assert r["firstResidual"] == pytest.approx(float(ir), rel=1e-13)
        assert r["rhsMax"] == pytest.approx(float(rm), rel=1e-13)
        assert r["sumRHS"] == pytest.approx(float(sr), rel=1e-12)
        exact += fmt(r["firstResidual"], 14) == ir
    assert exact >= 22
    # cg2d_res (the last residual, ~1e-14 = round-off of a 1e-13 solve): same magnitude
    for r, lr in zip(out, GOLDA["cg2d_last_res"]):
        assert r["lastResidual"] == pytest.approx(float(lr), rel=1e-3)


@pytest.mark.parametrize("fld", ["eta", "uvel", "vvel", "wvel"])
@pytest.mark.parametrize("st", ["max", "min", "mean", "sd"])
def test_adjustment_cs_monitor_dynstats(ac24, fld, st):
    """max / min / sd and the mean of eta: every printed digit (13) for the 24 steps, allowing one unit in
    the last place.  The means of u, v, w are sums that cancel to round-off on the symmetric cube (1e-19 at
    step 1, amplified by the flow): they are compared absolutely against 1e-12 of the field's magnitude."""
    _, out = ac24
    gold = GOLDA[f"dynstat_{fld}_{st}"][1:]
    for r, gv in zip(out, gold):
        if st == "mean" and fld != "eta":
            scale = max(abs(r[fld]["max"]), abs(r[fld]["min"]))
            assert abs(r[fld][st] - float(gv)) < 1e-12 * scale, (fld, st)
        else:
            assert r[fld][st] == pytest.approx(float(gv), rel=2e-13, abs=1e-30), (fld, st)


# ---------------------------------------------------------------------------------------
# verification/advect_xy (salt): multi-dimensional advection GAD_ADVECTION, scheme 33 (DST3 flux limiter),
# GAD_MULTIDIM_COMPRESSIBLE build, uniform diagonal flow on a doubly periodic grid, 80 steps.
# ---------------------------------------------------------------------------------------
def test_advect_xy_salt_statistics_every_printed_digit():
    from oracle import advect_xy as ax
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "advect_xy.json")))
    out = ax.run(80)
    assert len(out) == len(gold["dynstat_salt_sd"]) == 6          # steps 0, 16, ..., 80
    for r, mx, mn, me, sd in zip(out, gold["dynstat_salt_max"], gold["dynstat_salt_min"], gold["dynstat_salt_mean"],
                                 gold["dynstat_salt_sd"]):
        assert (fmt(r["max"], 13), fmt(r["min"], 13), fmt(r["mean"], 13), fmt(r["sd"], 13)) == (mx, mn, me, sd)


def test_advect_xy_is_independent_of_the_tiling():
    """2 x 2 tiles of 10 x 10 instead of 1 x 2 tiles of 20 x 10: same field statistics (the passes only
    read what the halo exchange provides)."""
    from oracle import advect_xy as ax
    a, b = ax.run(32), ax.run(32, nSx=2, nSy=2)
    for ra, rb in zip(a, b):
        assert ra["sd"] == pytest.approx(rb["sd"], rel=1e-13) and ra["max"] == rb["max"]


# ---------------------------------------------------------------------------------------
# verification/advect_xy, input.ab3_c4 (results/output.ab3_c4.txt): tempAdvScheme = saltAdvScheme = 4 through
# GAD_CALC_RHS (gad_c4_adv_{x,y}.F) with the third-order Adams-Bashforth scheme on the tendencies
# (adams_bashforth3.F), 100 steps.  Pins scheme 4 and AB3.
# ---------------------------------------------------------------------------------------
def test_advect_xy_ab3_c4_statistics_every_printed_digit():
    from oracle import advect_xy as ax
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "advect_xy.ab3_c4.json")))
    out = ax.run_ab3(100)
    assert len(out) == len(gold["dynstat_theta_sd"]) == 11          # steps 0, 10, ..., 100
    for i, (t, s) in enumerate(out):
        for r, fld in ((t, "theta"), (s, "salt")):
            for st in ("max", "min", "mean", "sd"):
                if fld == "theta" and st == "min" and i == 0:
                    assert r[st] == pytest.approx(float(gold[f"dynstat_{fld}_{st}"][i]), rel=1e-12)      # 6.5e-28: EXP round-off
                    continue
                assert fmt(r[st], 13) == gold[f"dynstat_{fld}_{st}"][i], (i, fld, st)


def test_ab3_start_up_factors():
    """adams_bashforth3.F:66-83: forward step, then the two-level form, then the full three-level form."""
    from oracle.advect_xy import ab3_factors
    a, b = 0.5, 0.281105
    assert ab3_factors(0, 0, 0, a, b) == (0.0, 0.0, 0.0)
    assert ab3_factors(1, 0, 0, a, b) == (a, -a, 0.0)
    assert ab3_factors(0, 0, 1, a, b) == (a, -a, 0.0)
    assert ab3_factors(2, 0, 0, a, b) == (a + b, -a - 2.0 * b, b)


# ---------------------------------------------------------------------------------------
# verification/solid-body.cs-32x32x1: solid-body rotation of a one-layer atmosphere on the cs32 cube,
# vectorInvariantMomentum = T.  Pins MOM_VECINV (relative vorticity with the three-cell facet corners,
# KE gradient, vorticity advection, Coriolis) and GAD_CALC_RHS on the cube (passive salt), 25 steps.
# ---------------------------------------------------------------------------------------
from oracle import solid_body_cs as sbc

GOLDS = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "solid-body.cs-32x32x1.json")))


@pytest.fixture(scope="module")
def sb25():
    return sbc.run(25)


def test_solid_body_cs_cg2d_lines_all_steps(sb25):
    """cg2dNorm, iteration counts, cg2d_init_res and rhsMax: every printed digit for all 25 steps
    (one unit in the last place allowed); Sum(rhs) is round-off of a 7e4 field and compared absolutely."""
    norm, out = sb25
    assert fmt(norm, 16) == GOLDS["cg2dNorm"]          # 9.8929536739060584E-06
    assert [r["numIters"] for r in out[1:]] == GOLDS["cg2d_iters"]
    exact = 0
    for r, ir, (sr, rm) in zip(out[1:], GOLDS["cg2d_init_res"], GOLDS["sumRHS_rhsMax"]):
        assert r["firstResidual"] == pytest.approx(float(ir), rel=2e-14)
        assert r["rhsMax"] == pytest.approx(float(rm), rel=2e-14)
        assert abs(r["sumRHS"] - float(sr)) < 1e-15 * 7.4e4 * 6144
        exact += fmt(r["firstResidual"], 14) == ir and fmt(r["rhsMax"], 14) == rm
    assert exact >= 23


@pytest.mark.parametrize("fld", ["eta", "uvel", "vvel", "wvel", "salt"])
@pytest.mark.parametrize("st", ["max", "min", "mean", "sd"])
def test_solid_body_cs_monitor_dynstats(sb25, fld, st):
    """Initial state and 25 steps: max / min / sd and the means of u, v, salt to every printed digit (one
    unit in the 13th place allowed); the means of eta and w cancel to round-off (1e-11 of 1e4, 1e-17) and
    are compared absolutely against 1e-13 of the field's magnitude."""
    _, out = sb25
    gold = GOLDS[f"dynstat_{fld}_{st}"]
    assert len(gold) == len(out) == 26
    for r, gv in zip(out, gold):
        if st == "mean" and fld in ("eta", "wvel"):
            scale = max(abs(r[fld]["max"]), abs(r[fld]["min"]), 1e-300)
            assert abs(r[fld][st] - float(gv)) < 1e-13 * scale, (fld, st)
        else:
            assert r[fld][st] == pytest.approx(float(gv), rel=2e-13, abs=1e-30), (fld, st)


# ---------------------------------------------------------------------------------------
# verification/advect_cs (theta): GAD_ADVECTION on the cs32 cubed sphere -- three facet-dependent passes with
# FILL_CS_CORNER_TR_RL / _UV_RS, scheme 33, GAD_MULTIDIM_COMPRESSIBLE build; 192 steps, monitor every 8.
# ---------------------------------------------------------------------------------------
def test_advect_cs_theta_statistics_every_printed_digit():
    from oracle import advect_cs as acs
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "advect_cs.json")))
    out = acs.run(192)
    assert len(out) == len(gold["dynstat_theta_sd"]) == 25          # steps 0, 8, ..., 192
    for r, mx, mn, me, sd in zip(out, gold["dynstat_theta_max"], gold["dynstat_theta_min"], gold["dynstat_theta_mean"],
                                 gold["dynstat_theta_sd"]):
        assert (fmt(r["max"], 13), fmt(r["min"], 13), fmt(r["mean"], 13), fmt(r["sd"], 13)) == (mx, mn, me, sd)


# ---------------------------------------------------------------------------------------
# CG3D: operator-level known answer only (the solver lines need the whole non-hydrostatic step).
# verification/tutorial_deep_convection: 100 x 100 x 50 cells of 20 m, 2 x 2 tiles of 50 x 50.
# ---------------------------------------------------------------------------------------
def test_ini_cg3d_normalisation_factor_of_tutorial_deep_convection():
    from mitgcm_b200.grid import Dims, cartesian_grid, masks_from_depth, global_area
    from oracle.pyoracle import Oracle
    d = Dims(sNx=50, sNy=50, OLx=2, OLy=2, nSx=2, nSy=2, Nr=50)
    g = cartesian_grid(d, [20.0] * 100, [20.0] * 100, [20.0] * 50, f0=1e-4, beta=0.0, gBaro=10.0)
    masks_from_depth(g, -1000.0 * np.ones((100, 100)), hFacMin=1.0)
    o = Oracle(g, dict(deltaTMom=20.0, deltaTFreeSurf=20.0, globalArea=global_area(g)))
    op = o.ini_cg3d(1.0, 1e-9, -1.0)
    assert fmt(op["cg3dNorm"], 16) == "5.0000000000000003E-02"      # results/output.txt: INI_CG3D: CG3D normalisation factor
    # the operator is symmetric-negative-definite on the wet cells: CG reduces the residual monotonically enough
    rng = np.random.default_rng(0)
    jj, ii = d.interior()
    b = np.zeros(d.shape3)
    b[..., jj, ii] = rng.standard_normal(b[..., jj, ii].shape)
    x = np.zeros(d.shape3)
    r = o.cg3d(op, b, x, 40)
    assert r["lastResidual"] < 0.02 * r["firstResidual"]


# ---------------------------------------------------------------------------------------
# verification/tutorial_deep_convection, all 3 steps of the golden: the NON-HYDROSTATIC step (CALC_GW, TIMESTEP_WVEL,
# the NH right-hand sides, PRE_CG3D, CG3D with 100 iterations per step, the correction with phi_nh) from the
# experiment's own start files.  PINS CG3D (cg3d_oracle.c) and, through cg2dUseMinResSol = 1, the minimum-residual
# solution of CG2D.  oracle/deep_convection.py; the restatement reproduced every printed digit on its first run.
# ---------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def deep_conv():
    from oracle import deep_convection as dc
    return dc.run(3)


DC_GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "tutorial_deep_convection.json")))


def test_deep_convection_solver_lines_every_printed_digit(deep_conv):
    op, op3, out, rec0 = deep_conv
    G = DC_GOLD
    assert fmt(op["cg2dNorm"], 16) == G["cg2dNorm"] and fmt(op3["cg3dNorm"], 16) == G["cg3dNorm"]
    assert [r["numIters"] for r in out] == G["cg2d_iters"] == [100, 123, 122]
    assert [r["nIterMin"] for r in out] == G["cg2d_iters_min"]                   # cg2dUseMinResSol = 1
    assert [r["cg3d"]["numIters"] for r in out] == G["cg3d_iters"] == [100, 100, 100]
    for n, r in enumerate(out):
        assert (fmt(r["sumRHS"], 14), fmt(r["rhsMax"], 14)) == tuple(G["sumRHS_rhsMax"][n])
        assert fmt(r["firstResidual"], 14) == G["cg2d_init_res"][n]
        assert fmt(r["lastResidual"], 14) == G["cg2d_last_res"][n]
        assert fmt(np.sqrt(r["minResidualSq"]), 14) == G["cg2d_min_res"][n]
        c = r["cg3d"]
        # Sum(rhs) is the round-off of a 5e5-term sum of 4e-3 numbers (2e-13): reproduced digit for digit all the same
        assert (fmt(c["sumRHS"], 14), fmt(c["rhsMax"], 14)) == tuple(G["cg3d_sumRHS_rhsMax"][n])
        assert fmt(c["firstResidual"], 14) == G["cg3d_init_res"][n]
        assert fmt(c["lastResidual"], 14) == G["cg3d_last_res"][n]               # after 100 CG3D iterations


@pytest.mark.parametrize("fld", ["eta", "uvel", "vvel", "wvel", "theta"])
def test_deep_convection_monitor_statistics_every_printed_digit(deep_conv, fld):
    _, _, out, rec0 = deep_conv
    for n, r in enumerate([rec0] + out):
        for st in ("max", "min", "mean", "sd"):
            gold = DC_GOLD[f"dynstat_{fld}_{st}"][n]
            if abs(float(gold)) < 1e-12:          # means that cancel to round-off (w, eta)
                assert abs(r[fld][st] - float(gold)) < 1e-15, (n, st)
            else:
                assert fmt(r[fld][st], 13) == gold, (n, fld, st)


# ---------------------------------------------------------------------------------------
# verification/adjustment.128x64x1: gravity-wave adjustment of a one-layer atmosphere (p coordinates) on the global
# lat-lon grid, pole to pole (zero-width faces closed by ADD_WALLS2MASKS), 2 x 2 tiles, 24 steps.  The golden was
# written by an older model version (older SOLVE_FOR_PRESSURE print-out): agreement is >= 11 digits, not every digit.
# ---------------------------------------------------------------------------------------
def test_adjustment_128x64_lat_lon_all_steps():
    from oracle import adjustment_latlon as al
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "adjustment.128x64x1.json")))
    norm, out = al.run(24)
    assert fmt(norm, 15) == "2.454122852291226E-07"            # golden: 0.2454122852291226263129944E-06
    assert [r["numIters"] for r in out] == gold["cg2d_iters"]
    for n, r in enumerate(out):
        assert r["firstResidual"] == pytest.approx(float(gold["cg2d_init_res"][n]), rel=1e-11)
        for f in ("eta", "uvel", "vvel"):
            for st in ("max", "min", "mean", "sd"):
                assert r[f][st] == pytest.approx(float(gold[f"dynstat_{f}_{st}"][n + 1]), rel=2e-11, abs=1e-12), (n, f, st)


# verification/tutorial_advection_in_gyre: the barotropic gyre restarted from a 10-year spin-up (|u| up to 0.26 m/s):
# MOM_FLUXFORM's advective terms on a developed flow, no-slip sides AND bottom (viscAz = 0.01), AB2 continued from the
# pickup's GuNm1 / GvNm1, CG2D started from Bo_surf * etaN (oracle/advection_in_gyre.py).  The flow is almost steady,
# so cg2d_init_res (7e-10) and the wvel statistics (1e-14) are differences of nearly equal numbers: reproducing their
# printed digits means the tendencies are the reference's bit for bit.
GOLD_AG = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "tutorial_advection_in_gyre.json")))


@pytest.fixture(scope="module")
def run_ag():
    from oracle import advection_in_gyre as ag
    return ag.run(4)


def test_advection_in_gyre_inputs_match_reference_files():
    ref = "/root/reference/verification/tutorial_advection_in_gyre/input"
    if not os.path.isdir(ref):
        pytest.skip("reference tree not present (GPU box)")
    from oracle import advection_in_gyre as ag
    z = np.load(ag.FIXTURE)
    pk = np.fromfile(os.path.join(ref, "pickup.0000259200.data"), ">f8").reshape(11, 60, 60)
    for q, n in ((0, "Uvel"), (1, "Vvel"), (4, "GuNm1"), (5, "GvNm1"), (8, "EtaN")):
        assert np.array_equal(z[n], pk[q]), n
    assert np.array_equal(z["topog"], np.fromfile(os.path.join(ref, "topog.box5000"), ">f8").reshape(60, 60))
    assert np.array_equal(z["windx"], np.fromfile(os.path.join(ref, "windx.m01cos2y"), ">f8").reshape(60, 60))


def test_advection_in_gyre_solver_lines_every_digit(run_ag):
    norm, _, out = run_ag
    assert fmt(norm, 16) == GOLD_AG["cg2dNorm"]
    assert [r["numIters"] for r in out] == GOLD_AG["cg2d_iters"] == [11, 10, 10, 11]
    for r, ir, lr, (sr, rm) in zip(out, GOLD_AG["cg2d_init_res"], GOLD_AG["cg2d_last_res"], GOLD_AG["sumRHS_rhsMax"]):
        assert fmt(r["firstResidual"], 14) == ir
        assert fmt(r["lastResidual"], 14) == lr
        assert fmt(r["rhsMax"], 14) == rm
        assert fmt(r["sumRHS"], 14) == sr


@pytest.mark.parametrize("fld", ["eta", "uvel", "vvel", "wvel"])
@pytest.mark.parametrize("st", ["max", "min", "mean", "sd"])
def test_advection_in_gyre_monitor_every_digit(run_ag, fld, st):
    _, first, out = run_ag
    gold = GOLD_AG[f"dynstat_{fld}_{st}"]
    assert len(gold) == 5
    assert fmt(first[fld][st], 13) == gold[0], "statistics of the pickup state"
    for i, r in enumerate(out):
        assert fmt(r[fld][st], 13) == gold[i + 1], (fld, st, i)


# verification/flt_example (the ocean underneath the float package): wind-driven f-plane channel over a bump with
# PARTIAL CELLS (hFacMin = 0.2), 80 x 42 x 8, stratified, explicit vertical diffusion: the time stepping on hFac < 1
# (oracle/flt_example.py); golden: results/output.with_flt.txt
GOLD_FE = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "flt_example.with_flt.json")))
FE_STEPS = 18


@pytest.fixture(scope="module")
def run_fe():
    from oracle import flt_example as fe
    return fe.run(FE_STEPS)


def test_flt_example_inputs_match_reference_files_and_have_partial_cells():
    from oracle import flt_example as fe
    z, d, g, P = fe.setup()
    part = (g.hFacC > 0) & (g.hFacC < 1)
    assert part.sum() > 1000 and g.hFacC[g.hFacC > 0].min() >= 0.2          # the point of this experiment
    ref = "/root/reference/verification/flt_example/input"
    if not os.path.isdir(ref):
        pytest.skip("reference tree not present (GPU box)")
    assert np.array_equal(z["topog"], np.fromfile(os.path.join(ref, "topog.bump"), ">f8").reshape(42, 80))
    assert np.array_equal(z["windx"], np.fromfile(os.path.join(ref, "windx.sin_y"), ">f8").reshape(42, 80))


def test_flt_example_solver_lines_every_digit(run_fe):
    norm, _, out = run_fe
    assert fmt(norm, 16) == GOLD_FE["cg2dNorm"]
    assert [r["numIters"] for r in out] == GOLD_FE["cg2d_iters"][:FE_STEPS]
    for r, ir, lr, (sr, rm) in zip(out, GOLD_FE["cg2d_init_res"], GOLD_FE["cg2d_last_res"], GOLD_FE["sumRHS_rhsMax"]):
        assert fmt(r["firstResidual"], 14) == ir
        assert fmt(r["lastResidual"], 14) == lr
        assert fmt(r["rhsMax"], 14) == rm
        assert abs(r["sumRHS"] - float(sr)) <= 1e-14 or fmt(r["sumRHS"], 14) == sr      # a sum of round-off (1e-12)


@pytest.mark.parametrize("fld", ["eta", "uvel", "vvel", "wvel", "theta"])
@pytest.mark.parametrize("st", ["max", "min", "mean", "sd"])
def test_flt_example_monitor_every_digit(run_fe, fld, st):
    _, first, out = run_fe
    gold = GOLD_FE[f"dynstat_{fld}_{st}"]
    assert float(fmt(first[fld][st], 13)) == float(gold[0]), "start state"
    for i, r in enumerate(out):
        assert fmt(r[fld][st], 13) == gold[i + 1], (fld, st, i)      # incl. the eta / wvel means, which are round-off (1e-17, 1e-22)


# verification/inverted_barometer: a closed stratified f-plane box under an atmospheric pressure load (phi0surf = pLoad /
# rhoConst added to the hydrostatic potential), no-slip bottom at Nr = 4, free-slip walls (oracle/inverted_barometer.py).
# The golden output was written by checkpoint62c (2010), before the bottom-drag routines were re-arranged: as for
# adjustment.128x64x1 the agreement is the reference's own pass rule (matching digits), not every printed digit.
GOLD_IB = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "inverted_barometer.json")))


@pytest.fixture(scope="module")
def run_ib():
    from oracle import inverted_barometer as ib
    return ib.run(40)


def _digits(a, b):
    return 99.0 if a == b else -np.log10(abs(a - b) / (0.5 * (abs(a) + abs(b))))


def test_inverted_barometer_solver_lines(run_ib):
    norm, _, out = run_ib
    assert fmt(norm, 16) == GOLD_IB["cg2dNorm"]
    assert [r["numIters"] for r in out] == GOLD_IB["cg2d_iters"]              # 40 solves to 1e-13: 35 36 35 35 34 ...
    assert fmt(out[0]["firstResidual"], 14) == GOLD_IB["cg2d_init_res"][0]   # the first step: every printed digit
    assert fmt(out[0]["lastResidual"], 14) == GOLD_IB["cg2d_last_res"][0]
    for r, ir, (_, rm) in zip(out, GOLD_IB["cg2d_init_res"], GOLD_IB["sumRHS_rhsMax"]):
        assert _digits(r["firstResidual"], float(ir)) >= 13.5
        assert _digits(r["rhsMax"], float(rm)) >= 13.5


@pytest.mark.parametrize("fld", ["eta", "uvel", "vvel", "wvel", "theta"])
@pytest.mark.parametrize("st", ["max", "min", "sd"])
def test_inverted_barometer_monitor(run_ib, fld, st):
    _, first, out = run_ib
    gold = GOLD_IB[f"dynstat_{fld}_{st}"]
    assert len(gold) == 41
    assert fmt(out[0][fld][st], 13) == gold[1]                                # the first step: every printed digit
    for i, r in enumerate(out):
        assert _digits(r[fld][st], float(gold[i + 1])) >= 12.5, (fld, st, i)


def test_inverted_barometer_adjusts_towards_the_load(run_ib):
    """what the experiment is about: the sea surface moves towards -pLoad / (rhoConst g)"""
    from oracle import inverted_barometer as ib
    z, d, g, P = ib.setup()
    target = np.abs(z["pLoad"]).max() / (999.8 * 9.81)
    _, _, out = run_ib
    assert out[0]["eta"]["max"] < out[10]["eta"]["max"] < 1.3 * target and out[-1]["eta"]["max"] > 0.5 * target


def test_cg2d_sr_in_place_of_cg2d_meets_the_flt_example_golden():
    """CG2D_SR has no golden of its own (the one experiment that sets useSRCGSolver needs sea ice and GM/Redi).  Held to a
    reference output anyway: with the single-reduction solver in place of CG2D the flt_example run must pass the reference's
    own rule against results/output.with_flt.txt -- here: the same iteration count in all 18 solves (about 130 each) and
    >= 13 of the 14 printed digits of cg2d_init_res and of every max / min / sd."""
    from oracle import flt_example as fe
    _, _, out = fe.run(18, sr=True)
    assert [r["numIters"] for r in out] == GOLD_FE["cg2d_iters"]
    for i, r in enumerate(out):
        if float(GOLD_FE["cg2d_init_res"][i]) != 0.0:
            assert _digits(r["firstResidual"], float(GOLD_FE["cg2d_init_res"][i])) >= 13.0, i
        for fld in ("eta", "uvel", "vvel", "wvel", "theta"):
            for st in ("max", "min", "sd"):
                assert _digits(r[fld][st], float(GOLD_FE[f"dynstat_{fld}_{st}"][i + 1])) >= 13.0, (i, fld, st)


# verification/matrix_example: another barotropic gyre restarted from a pickup (32 x 32 cells of 50 km, 2 x 4 tiles of
# 16 x 8, OL = 3, deltaT = 20000 s, rhoConst = 1035): the sequence of oracle/advection_in_gyre.py with other numbers
GOLD_MX = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "matrix_example.json")))


def test_matrix_example_every_digit():
    from oracle import matrix_example as mx
    norm, first, out = mx.run(10)
    assert fmt(norm, 16) == GOLD_MX["cg2dNorm"]
    assert [r["numIters"] for r in out] == GOLD_MX["cg2d_iters"]
    for i, r in enumerate(out):
        assert fmt(r["firstResidual"], 14) == GOLD_MX["cg2d_init_res"][i]
        assert fmt(r["lastResidual"], 14) == GOLD_MX["cg2d_last_res"][i]
        assert (fmt(r["sumRHS"], 14), fmt(r["rhsMax"], 14)) == tuple(GOLD_MX["sumRHS_rhsMax"][i])
    for fld in ("eta", "uvel", "vvel", "wvel"):
        for st in ("max", "min", "mean", "sd"):
            gold = GOLD_MX[f"dynstat_{fld}_{st}"]
            assert fmt(first[fld][st], 13) == gold[0], (fld, st)
            for i, r in enumerate(out):
                assert fmt(r[fld][st], 13) == gold[i + 1], (fld, st, i)
