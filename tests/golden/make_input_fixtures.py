"""Copies the small binary INPUT vectors of the reference's verification experiments that the
CPU tests need (run in the build container, where /root/reference exists; the GPU box has no
reference tree).  Outputs go to tests/golden/inputs/.  Usage: python tests/golden/make_input_fixtures.py"""
import os
import shutil

REF = "/root/reference/verification"
HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "inputs")
FILES = {
    # config 3 (global_ocean.90x40x15 links its *.bin from tutorial_global_oce_latlon, input/prepare_run)
    "global_oce_latlon_bathymetry.bin": "tutorial_global_oce_latlon/input/bathymetry.bin",
}

def cs32_fixture():
    """Config 4 (global_ocean.cs32x15): the records of grid_cs32.face00N.bin (horizGridFile, taken from
    tutorial_held_suarez_cs/input by prepare_run) that the CG2D operator needs, and bathy_Hmin50.bin,
    as one compressed npz (float64, exact)."""
    import numpy as np
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
    from mitgcm_b200.grid import read_mitgrid_faces
    faces = read_mitgrid_faces(os.path.join(REF, "tutorial_held_suarez_cs/input/grid_cs32"), 32)
    keep = "xC yC rA xG yG dxC dyC dxG dyG rAw rAs".split()
    out = {f"{n}_{f}": faces[f][n] for f in range(6) for n in keep}
    out["bathy_Hmin50"] = np.fromfile(os.path.join(REF, "global_ocean.cs32x15/input/bathy_Hmin50.bin"), ">f8").reshape(32, 192).astype(np.float64)
    # adjustment.cs-32x32x1 (same grid): flat 1366 m bathymetry and the initial free-surface bump
    adj = os.path.join(REF, "adjustment.cs-32x32x1/input")
    out["adj_bathy_f2"] = np.fromfile(os.path.join(adj, "bathy_f2.bin"), ">f8").reshape(32, 192).astype(np.float64)
    out["adj_ssh_eq"] = np.fromfile(os.path.join(adj, "ssh_eq.bin"), ">f8").reshape(32, 192).astype(np.float64)
    # advect_cs (same grid): initial tracer (W2_mapIO = -1: facets stacked along x)
    out["advcs_T_init"] = np.fromfile(os.path.join(REF, "advect_cs/input/T.init"), ">f8").reshape(32, 192).astype(np.float64)
    np.savez_compressed(os.path.join(HERE, "cs32_grid_bathy.npz"), **out)
    print("wrote cs32_grid_bathy.npz")


def solid_body_fixture():
    """solid-body.cs-32x32x1: its own tile00N.mitgrid files (16 records, no angles; a different cs32
    grid from grid_cs32.face00N.bin) and the initial tracer S_init.bin, as one compressed npz."""
    import numpy as np
    sb = os.path.join(REF, "solid-body.cs-32x32x1/input")
    names = "xC yC dxF dyF rA xG yG dxV dyU rAz dxC dyC rAw rAs dxG dyG".split()
    out = {}
    for f in range(6):
        a = np.fromfile(os.path.join(sb, f"tile{f + 1:03d}.mitgrid"), ">f8").reshape(16, 33, 33).astype(np.float64)
        for q, n in enumerate(names):
            out[f"{n}_{f}"] = a[q]
    out["S_init"] = np.fromfile(os.path.join(sb, "S_init.bin"), ">f8").reshape(192, 32).astype(np.float64)   # W2_mapIO = 1: facets stacked along y
    np.savez_compressed(os.path.join(HERE, "solid_body_cs32.npz"), **out)
    print("wrote solid_body_cs32.npz")


def deep_convection_fixture():
    """tutorial_deep_convection (the non-hydrostatic step around CG3D): start state T / U / V / Eta at 120 min and the
    surface heat flux, real*4 big-endian as the model reads them (readBinaryPrec = 32), kept as float32."""
    import numpy as np
    dc = os.path.join(REF, "tutorial_deep_convection/input")
    out = {}
    for n, f in (("T", "T.120mn.bin"), ("U", "U.120mn.bin"), ("V", "V.120mn.bin")):
        out[n] = np.fromfile(os.path.join(dc, f), ">f4").reshape(50, 100, 100).astype(np.float32)
    for n, f in (("Eta", "Eta.120mn.bin"), ("Qnet", "Qnet_p32.bin")):
        out[n] = np.fromfile(os.path.join(dc, f), ">f4").reshape(100, 100).astype(np.float32)
    np.savez_compressed(os.path.join(HERE, "deep_convection.npz"), **out)
    print("wrote deep_convection.npz")


def advection_in_gyre_fixture():
    """tutorial_advection_in_gyre: the spun-up barotropic gyre it restarts from (pickup.0000259200.data, 11 records of
    60 x 60 float64: Uvel Vvel Theta Salt GuNm1 GvNm1 GtNm1 GsNm1 EtaN dEtaHdt EtaH), the basin and the wind stress."""
    import numpy as np
    ag = os.path.join(REF, "tutorial_advection_in_gyre/input")
    pk = np.fromfile(os.path.join(ag, "pickup.0000259200.data"), ">f8").reshape(11, 60, 60).astype(np.float64)
    names = "Uvel Vvel Theta Salt GuNm1 GvNm1 GtNm1 GsNm1 EtaN dEtaHdt EtaH".split()
    out = {n: pk[q] for q, n in enumerate(names) if n in ("Uvel", "Vvel", "Theta", "GuNm1", "GvNm1", "GtNm1", "EtaN")}
    out["topog"] = np.fromfile(os.path.join(ag, "topog.box5000"), ">f8").reshape(60, 60).astype(np.float64)
    out["windx"] = np.fromfile(os.path.join(ag, "windx.m01cos2y"), ">f8").reshape(60, 60).astype(np.float64)
    np.savez_compressed(os.path.join(HERE, "advection_in_gyre.npz"), **out)
    print("wrote advection_in_gyre.npz")


def inverted_barometer_fixture():
    """inverted_barometer: the closed basin and the atmospheric pressure load, 60 x 60 float64."""
    import numpy as np
    ib = os.path.join(REF, "inverted_barometer/input")
    np.savez_compressed(os.path.join(HERE, "inverted_barometer.npz"),
                        topog=np.fromfile(os.path.join(ib, "topog.box"), ">f8").reshape(60, 60).astype(np.float64),
                        pLoad=np.fromfile(os.path.join(ib, "pLoad.bin"), ">f8").reshape(60, 60).astype(np.float64))
    print("wrote inverted_barometer.npz")


def matrix_example_fixture():
    """matrix_example: pickup.0000200000.data (no field list: the older record order Uvel GuNm1 Vvel GvNm1 Theta GtNm1
    Salt GsNm1 EtaN dEtaHdt EtaH, 32 x 32 float64), basin and wind stress (float32 files: readBinaryPrec = 32)."""
    import numpy as np
    me = os.path.join(REF, "matrix_example/input")
    pk = np.fromfile(os.path.join(me, "pickup.0000200000.data"), ">f8").reshape(11, 32, 32).astype(np.float64)
    names = "Uvel GuNm1 Vvel GvNm1 Theta GtNm1 Salt GsNm1 EtaN dEtaHdt EtaH".split()
    out = {n: pk[q] for q, n in enumerate(names) if n in ("Uvel", "Vvel", "Theta", "GuNm1", "GvNm1", "GtNm1", "EtaN")}
    out["topog"] = np.fromfile(os.path.join(me, "topo_box.bin"), ">f4").reshape(32, 32).astype(np.float64)
    out["windx"] = np.fromfile(os.path.join(me, "taux_cosY.bin"), ">f4").reshape(32, 32).astype(np.float64)
    np.savez_compressed(os.path.join(HERE, "matrix_example.npz"), **out)
    print("wrote matrix_example.npz")


def flt_example_fixture():
    """flt_example: the bump topography (partial cells with hFacMin = 0.2) and the zonal wind stress, 42 x 80 float64."""
    import numpy as np
    fe = os.path.join(REF, "flt_example/input")
    out = dict(topog=np.fromfile(os.path.join(fe, "topog.bump"), ">f8").reshape(42, 80).astype(np.float64),
               windx=np.fromfile(os.path.join(fe, "windx.sin_y"), ">f8").reshape(42, 80).astype(np.float64))
    np.savez_compressed(os.path.join(HERE, "flt_example.npz"), **out)
    print("wrote flt_example.npz")


if __name__ == "__main__":
    os.makedirs(HERE, exist_ok=True)
    cs32_fixture()
    solid_body_fixture()
    deep_convection_fixture()
    advection_in_gyre_fixture()
    flt_example_fixture()
    inverted_barometer_fixture()
    matrix_example_fixture()
    for dst, src in FILES.items():
        shutil.copyfile(os.path.join(REF, src), os.path.join(HERE, dst))
        print("copied", src, "->", dst)
