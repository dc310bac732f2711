"""pkg/exch2 tile graph spread over several GPUs (SURVEY.md section 8 rows a10 / e), run under torchrun by
tests/test_exch2_dist_gpu.py:  torchrun --nproc-per-node N tests/dist_cs_worker.py

Every rank holds nTiles / N tiles of the cubed sphere (W2_tileProc in consecutive blocks, w2_map_procs.F) and
checks ITS tiles against results computed for the whole graph on the CPU:
  A  EXCH2_3D_RX (scalar, 2-D and 3-D) and EXCH2_UV_3D_RX (signed and unsigned) across ranks: bit-identical to
     the literal two-pass buffered algorithm of the oracle (oracle/exch2_oracle.py) -- the MPI messages of
     exch2_send_rx{1,2}.template / exch2_recv_rx{1,2}.template become reads of the owner's peer arena;
  B  CG2D / CG2D_SR on the config-4 operator (cs32, real bathymetry) across ranks, fixed iteration counts:
     normalised RHS bit-exact, x to 1e-11 * iterations, residual history 1e-9 (as the one-GPU test);
  C  verification/adjustment.cs-32x32x1 stepped entirely on the devices (48 tiles over N GPUs, semi-implicit free
     surface, exactConserv, signed vector exchange) against the experiment's golden output, 24 steps.
Prints "DIST_CS PASS" on rank 0 when every rank passed."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
import numpy as np
import torch
import torch.distributed as dist

from mitgcm_b200 import distributed, runtime as rt
from mitgcm_b200.exch2 import cubed_sphere_topology, set_topology, tile_proc
from mitgcm_b200.grid import Dims
from mitgcm_b200.model import Model, ini_cg2d_tilegraph, rank_tiles
from oracle import adjustment_cs as ac
from oracle import exch2_oracle as eo
from oracle.baroclinic_gyre import mon_stats
from oracle.pyoracle import Oracle

local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
fails = []
PARTS = os.environ.get("DIST_CS_PARTS", "ABC")


def check(ok, what):
    if not ok:
        fails.append(what)
        print(f"rank {rank}: FAIL {what}", flush=True)


def gather_tiles(a):
    """(1, nLocal, ...) arrays of every rank -> the (1, nTiles, ...) array of the whole graph, on every rank"""
    t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    parts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(parts, t)
    return torch.cat(parts, dim=1).cpu().numpy()


# ---- A: exchanges ---------------------------------------------------------------------------------
for nf, sx, sy, OL, Nr in ((32, 32, 16, 4, 3), (8, 4, 4, 2, 2), (6, 6, 6, 3, 1)):
    T = cubed_sphere_topology(nf, sx, sy)
    if T.nTiles % world or "A" not in PARTS:
        continue
    n = T.nTiles // world
    lo, hi = rank * n, (rank + 1) * n
    dG = Dims(sNx=sx, sNy=sy, OLx=OL, OLy=OL, nSx=T.nTiles, nSy=1, Nr=Nr)
    d = Dims(sNx=sx, sNy=sy, OLx=OL, OLy=OL, nSx=n, nSy=1, Nr=Nr, nPx=world, nPy=1, myPx=rank, myPy=0)
    rt.init(d, local)
    distributed.setup(d)
    set_topology(T, tileProc=tile_proc(T.nTiles, world))
    rng = np.random.default_rng(7)          # the same global fields on every rank
    a3, a2 = rng.standard_normal(dG.shape3), rng.standard_normal(dG.shape2)
    for signs in (True, False):
        u3, v3 = rng.standard_normal(dG.shape3), rng.standard_normal(dG.shape3)
        rt.set_field("uVel", np.ascontiguousarray(u3[:, lo:hi]))
        rt.set_field("vVel", np.ascontiguousarray(v3[:, lo:hi]))
        rt.exch_uv("uVel", "vVel", signs)
        gu, gv = rt.get_field("uVel", np.zeros(d.shape3)), rt.get_field("vVel", np.zeros(d.shape3))
        eo.exch2_uv_3d(T, u3[0], v3[0], OL, signs)
        check(np.array_equal(gu, u3[:, lo:hi]) and np.array_equal(gv, v3[:, lo:hi]), f"A uv signs={signs} {nf}/{sx}x{sy}")
    rt.set_field("theta", np.ascontiguousarray(a3[:, lo:hi]))
    rt.set_field("etaN", np.ascontiguousarray(a2[:, lo:hi]))
    for _ in range(3):                      # repeated exchanges: the barrier sequence numbers advance
        rt.exch("theta")
        rt.exch("etaN")
    g3, g2 = rt.get_field("theta", np.zeros(d.shape3)), rt.get_field("etaN", np.zeros(d.shape2))
    eo.exch2_3d(T, a3[0], OL)
    r2 = a2[0][:, None].copy()
    eo.exch2_3d(T, r2, OL)
    check(np.array_equal(g3, a3[:, lo:hi]) and np.array_equal(g2[0], r2[lo:hi, 0]), f"A scalar {nf}/{sx}x{sy}")
    dist.barrier()
    distributed.teardown()
    rt.finalize()
    dist.barrier()

# ---- B: CG2D on the config-4 operator --------------------------------------------------------------
from helpers import load_cs32

T, g, P = load_cs32()
if T.nTiles % world == 0 and "B" in PARTS:
    dG = g.d
    op = ini_cg2d_tilegraph(g, P, T)
    o = Oracle(g, P)
    hook = eo.Exch2Hook(o, T, dG.OLx)
    rng = np.random.default_rng(5)
    jj, ii = dG.interior()
    wet = g.maskC[:, :, 0]
    b = np.zeros(dG.shape2)
    b[:, :, jj, ii] = rng.standard_normal((1, 12, 16, 32)) * wet[:, :, jj, ii] * 1e-3
    b[:, :, jj, ii] -= b[:, :, jj, ii].sum() / wet[:, :, jj, ii].sum() * wet[:, :, jj, ii]
    b *= g.rA
    x = 0.01 * rng.standard_normal(dG.shape2) * wet
    gl, (opl,) = rank_tiles(g, [op], rank, world)
    d = gl.d
    n = dG.nSx // world
    lo, hi = rank * n, (rank + 1) * n

    def solves(seq, label="B", x0=x, device=False):
        """fresh context, then the solves of seq = [(sr, nit), ...] back to back"""
        rt.init(d, local)
        rt.set_grid(gl)
        distributed.setup(d)
        set_topology(T, tileProc=tile_proc(T.nTiles, world))
        rt.set_cg2d_operator(opl)
        for sr, nit in seq:
            bo, xo = b.copy(), x0.copy()
            ro = o.cg2d(op, bo, xo, nit, -1, sr=sr, history=True)
            bg, xg = b[:, lo:hi].copy(), x0[:, lo:hi].copy()      # copies: the solver works in place (a slice of a (1, n, ..) array is a view)
            if device:
                tb, tx = torch.from_numpy(bg).cuda(), torch.from_numpy(xg).cuda()
                rg = rt.cg2d(tb, tx, nit, -1, sr=sr, residuals=True)
                bg, xg = tb.cpu().numpy(), tx.cpu().numpy()
            else:
                rg = rt.cg2d(bg, xg, nit, -1, sr=sr, residuals=True)
            sc = np.abs(xo[:, :, jj, ii]).max()
            check(rg["numIters"] == ro["numIters"] == nit, f"{label} iters sr={sr} nit={nit}")
            check(np.array_equal(bg[:, :, jj, ii], bo[:, lo:hi][:, :, jj, ii]), f"{label} rhs sr={sr} nit={nit}")
            ex = np.abs(xg[:, :, jj, ii] - xo[:, lo:hi][:, :, jj, ii]).max() / sc
            check(ex <= 1e-11 * nit, f"{label} x sr={sr} nit={nit}: rel err {ex:.3e}")
            nh = nit - 1 if sr else nit       # CG2D_SR records iterations 1 .. numIters-1 (cg2d_sr.F: the last residual is lastResidual)
            check(np.allclose(rg["hist"][:nh], np.asarray(ro["hist"])[:nh], rtol=1e-9, atol=0), f"{label} residual history sr={sr} nit={nit}")
            check(abs(rg["lastResidual"] / ro["lastResidual"] - 1.0) <= 1e-9, f"{label} lastResidual sr={sr} nit={nit}")
            if os.environ.get("DIST_CS_VERBOSE"):
                e = np.abs(xg[:, :, jj, ii] - xo[:, lo:hi][:, :, jj, ii])[0] / sc
                bad = e > 1e-9
                edge = np.zeros_like(bad)
                edge[:, 0, :] = edge[:, -1, :] = edge[:, :, 0] = edge[:, :, -1] = True
                at = np.unravel_index(e.argmax(), e.shape)
                print(f"rank {rank} {label} sr={sr} nit={nit}: rel err {ex:.3e} at (tile,j,i)={tuple(int(v) for v in at)}; cells off: "
                      f"{int(bad.sum())} of {bad.size}, on tile edges {int((bad & edge).sum())}, per tile {bad.reshape(bad.shape[0], -1).sum(1).tolist()}; "
                      f"firstRes {rg['firstResidual']:.6e} vs {ro['firstResidual']:.6e} hist {rg['hist'][:3]} vs {np.asarray(ro['hist'])[:3]}", flush=True)
        dist.barrier()
        distributed.teardown()
        rt.finalize()
        dist.barrier()

    solves([(sr, nit) for sr in (False, True) for nit in (1, 2, 7, 25)])
    solves([(False, 25), (True, 7)], "B device pointers", device=True)
    hook.close()

# ---- C: adjustment.cs-32x32x1 on the devices ---------------------------------------------------------
GOLD = json.load(open(os.path.join(HERE, "golden", "adjustment.cs-32x32x1.json")))
T, dG, g, P, ssh = ac.setup()
if T.nTiles % world == 0 and "C" in PARTS:
    P = dict(P)
    P.update(abEps=0.1, deltaTtracer=900.0, viscAr=0.0, tempStepping=0, cg2dMaxIters=600, momForcing=1,
             momDissip_In_AB=1, exactConserv=1, diffKhT=0.0, diffK4T=0.0, diffKrT=0.0)
    op = ini_cg2d_tilegraph(g, P, T)
    etaN = ac.tile_from_xstack(T, dG, ssh)
    eo.exch2_3d(T, etaN[0][:, None], dG.OLx)
    z3 = np.zeros(dG.shape3)
    state = dict(uVel=z3, vVel=z3, wVel=z3, theta=z3, etaN=etaN, etaH=etaN.copy(), surfForcU=np.zeros(dG.shape2),
                 surfForcV=np.zeros(dG.shape2))
    m = Model(g, P, state, op, device=local, topo=T, ranks=(rank, world))
    maskInC, maskInW, maskInS = g.maskC[:, :, 0], g.maskW[:, :, 0], g.maskS[:, :, 0]
    nsteps = int(os.environ.get("DIST_CS_STEPS", "24"))
    for it in range(nsteps):
        r = m.step()
        check(abs(r["numIters"] - GOLD["cg2d_iters"][it]) <= 1, f"C iterations step {it}: {r['numIters']}")
        check(abs(r["firstResidual"] / float(GOLD["cg2d_init_res"][it]) - 1.0) <= 1e-10, f"C cg2d_init_res step {it}")
        eta, u, v, w = (gather_tiles(m.get(n)) for n in ("etaN", "uVel", "vVel", "wVel"))
        st = dict(eta=mon_stats(dG, eta[:, :, None], maskInC[:, :, None], maskInC, g.rA, [g.drF[0]]),
                  uvel=mon_stats(dG, u, g.hFacW, maskInW, g.rAw, g.drF),
                  vvel=mon_stats(dG, v, g.hFacS, maskInS, g.rAs, g.drF),
                  wvel=mon_stats(dG, w, g.maskC, maskInC, g.rA, g.drC[:1]))
        for f in ("eta", "uvel", "vvel", "wvel"):
            for s in ("max", "min", "sd"):
                ref = float(GOLD[f"dynstat_{f}_{s}"][it + 1])
                check(abs(st[f][s] - ref) <= 1e-9 * abs(ref) + 1e-30, f"C dynstat_{f}_{s} step {it}: {st[f][s]!r} vs {ref!r}")
    dist.barrier()
    m.close()
    dist.barrier()

t = torch.tensor([float(len(fails))], device="cuda")
dist.all_reduce(t)
if rank == 0:
    print("DIST_CS", "PASS" if t.item() == 0 else "FAIL", f"ranks={world} failures={int(t.item())}", flush=True)
dist.destroy_process_group()
sys.exit(0 if t.item() == 0 else 1)
