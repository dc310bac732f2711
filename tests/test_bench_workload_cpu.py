"""The bench workload (bench.py: block-periodic channel, f = f0 + df sin(2 pi y / Ly_block), linear EOS) must be a
stable model run: the CPU oracle steps a small block of it and the flow has to stay bounded.  Round 1's beta plane
over the global extent reached f dt = 1.1 ... 2.1 at 2, 4, 8 ranks and blew up under explicit AB2 Coriolis.

python tests/test_bench_workload_cpu.py [n [nr [steps [nSx nSy]]]] runs longer probes."""
import sys
import os

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def bench_params():
    import bench
    return bench.params(0)


def run(n=32, nr=12, steps=60, nSx=1, nSy=1, verbose=False):
    from mitgcm_b200.model import make_channel
    from oracle.channel import ChannelOracle
    P = bench_params()
    P.update({k: float(os.environ[k]) for k in ("tTop", "tBot", "tNoise", "deltaTMom") if k in os.environ})
    if "deltaTMom" in os.environ:
        P["deltaTFreeSurf"] = P["deltaTtracer"] = P["deltaTMom"]
    g, P2, s = make_channel(n, n, nr, nSx=nSx, nSy=nSy, block=(n, n), dz=5000.0 / nr, **P)      # n x n cells per tile = per block
    fdt = np.abs(g.a["fCori"]).max() * P["deltaTMom"]
    co = ChannelOracle(g, P2, s, threads=min(8, nSx * nSy))
    hist = []
    for it in range(steps):
        r = co.step()
        umax = max(np.abs(co.s["uVel"]).max(), np.abs(co.s["vVel"]).max())
        hist.append((r["numIters"], umax, np.abs(co.s["etaN"]).max()))
        if verbose and (it % 10 == 0 or it == steps - 1):
            print(it, r["numIters"], f"{umax:.4f} eta {hist[-1][2]:.4f} theta {co.s['theta'].min():.3f}..{co.s['theta'].max():.3f}", flush=True)
        assert np.isfinite(umax), f"non-finite at step {it}"
    return fdt, hist


def test_bench_workload_is_stable():
    fdt, hist = run(32, 12, 80)
    assert fdt <= 0.2
    u = np.array([h[1] for h in hist])
    assert u.max() < 0.5 and u[-1] < 1.5 * u[:10].max(), u
    its = np.array([h[0] for h in hist])
    assert its.max() < 1000


def test_bench_workload_same_on_more_blocks():
    """2 x 2 blocks of the same workload = what 4 ranks step in the weak-scaled bench: the state is the exact periodic
    tiling of one block, so the flow statistics and the CG2D iteration counts are those of the one-block run."""
    _, h1 = run(48, 4, 8)
    _, h4 = run(48, 4, 8, 2, 2)
    for a, b in zip(h1, h4):
        # CG2D stops on the GLOBAL sum of r^2 (cg2d.F:204, 337): 4 identical blocks carry 4 x the sum, which costs
        # the couple of iterations that halve the residual -- nothing else may move
        assert 0 <= b[0] - a[0] <= 3, (a, b)
        assert abs(a[1] - b[1]) < 1e-9 * a[1], (a, b)


if __name__ == "__main__":
    a = [int(x) for x in sys.argv[1:]]
    n, nr, steps = (a + [64, 50, 200][len(a):])[:3]
    nS = a[3:5] if len(a) >= 5 else [1, 1]
    fdt, hist = run(n, nr, steps, nS[0], nS[1], verbose=True)
    print("f dt max", fdt)
