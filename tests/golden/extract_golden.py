"""Extracts the golden values that pin the oracle from the reference's own
results/output.txt files (run in the build container, where /root/reference
exists; the JSON it writes is committed so the GPU box does not need the
reference).  Usage: python tests/golden/extract_golden.py"""
import json
import os
import re

REF = "/root/reference/verification"
HERE = os.path.dirname(os.path.abspath(__file__))
NUM = r"([-+]?\d\.\d+E[-+]\d+)"


def parse(exp, out="results/output.txt"):
    txt = open(os.path.join(REF, exp, out)).read()
    g = {}
    m = re.search(r"CG2D normalisation factor =\s*" + NUM, txt)
    g["cg2dNorm"] = m.group(1) if m else None
    g["sumRHS_rhsMax"] = re.findall(r"cg2d: Sum\(rhs\),rhsMax =\s*" + NUM + r"\s+" + NUM, txt)
    g["cg2d_init_res"] = re.findall(r"cg2d_init_res =\s*" + NUM, txt)
    g["cg2d_iters"] = [int(x) for x in re.findall(r"cg2d_iters\(min,last\) =\s*-?\d+\s+(\d+)", txt)]
    g["cg2d_last_res"] = re.findall(r"cg2d_last_res =\s*" + NUM, txt)
    if not g["cg2d_iters"]:      # output written by an older SOLVE_FOR_PRESSURE: "cg2d_iters =  5", "cg2d_res = ..."
        g["cg2d_iters"] = [int(x) for x in re.findall(r"cg2d_iters =\s*(\d+)", txt)]
        g["cg2d_last_res"] = re.findall(r"cg2d_res =\s*" + NUM, txt)
    # the non-hydrostatic solver lines (solve_for_pressure.F:437-447, cg3d.F:243-244) and the min-residual iteration
    g["cg3d_sumRHS_rhsMax"] = re.findall(r"cg3d: Sum\(rhs\),rhsMax =\s*" + NUM + r"\s+" + NUM, txt)
    g["cg3d_init_res"] = re.findall(r"cg3d_init_res =\s*" + NUM, txt)
    g["cg3d_last_res"] = re.findall(r"cg3d_last_res =\s*" + NUM, txt)
    g["cg3d_iters"] = [int(x) for x in re.findall(r"cg3d_iters \(last\) =\s*(\d+)", txt)]
    g["cg2d_iters_min"] = [int(x) for x in re.findall(r"cg2d_iters\(min,last\) =\s*(-?\d+)\s+\d+", txt)]
    g["cg2d_min_res"] = re.findall(r"cg2d_min_res  =\s*" + NUM, txt)
    m = re.search(r"CG3D normalisation factor =\s*" + NUM, txt)
    g["cg3dNorm"] = m.group(1) if m else None
    for fld in ("eta", "uvel", "vvel", "wvel", "theta", "salt"):
        for st in ("max", "min", "mean", "sd"):
            g[f"dynstat_{fld}_{st}"] = re.findall(rf"%MON dynstat_{fld}_{st}\s+=\s*" + NUM, txt)
    return g


if __name__ == "__main__":
    for exp in ("tutorial_barotropic_gyre", "tutorial_baroclinic_gyre", "global_ocean.90x40x15",
                "global_ocean.cs32x15", "adjustment.cs-32x32x1", "advect_xy", "solid-body.cs-32x32x1", "advect_cs", "adjustment.128x64x1",
                "tutorial_deep_convection", "tutorial_advection_in_gyre", "inverted_barometer", "matrix_example"):
        with open(os.path.join(HERE, exp + ".json"), "w") as f:
            json.dump(parse(exp), f, indent=1)
        print("wrote", exp)
    # secondary outputs of an experiment (results/output.<name>.txt)
    for exp, name in (("advect_xy", "ab3_c4"), ("flt_example", "with_flt")):
        with open(os.path.join(HERE, f"{exp}.{name}.json"), "w") as f:
            json.dump(parse(exp, f"results/output.{name}.txt"), f, indent=1)
        print("wrote", exp, name)
