// cg2d.cu -- CG2D (model/src/cg2d.F:13-415) and CG2D_SR (model/src/cg2d_sr.F:13-458) as ONE
// persistent cooperative kernel per solve.
//
// The reference runs four sweeps per iteration with three global sums and two halo
// exchanges.  Here an iteration is two fused sweeps separated by grid-wide barriers:
//   phase B : s = z + beta*s (recomputed on the 5-point stencil from z and the old s, so no
//             separate s sweep and no s exchange), q = A s, partial <s,q>
//   phase CA: x += alpha s ; r = r - alpha q (recomputed on the stencil from the old r and q),
//             z = M r, partial <r,r> and <z,r>
// r and s are double-buffered so neighbours always read the previous iterate.  Edge points
// are "pushed" into the halo cell that mirrors them (ctx.pushTab), which replaces
// EXCH_S3D_RL (eesupp/src/exch_s3d_rx.template) without a separate pass.  Dot products are a
// fixed-shape reduction (lane accumulate -> warp shuffle -> CTA -> ordered sum over CTAs),
// replacing GLOBAL_SUM_TILE_RL (eesupp/src/global_sum_tile.F); run-to-run deterministic.
// Algorithmic traffic: 7 + 10 = 17 words = 136 B per point per iteration.
//
// Compiled with -fmad=false: every expression is evaluated in the reference's order without
// contraction, so single operator applies are bit-identical to the Fortran -ieee build; only
// the summation order of the dot products differs.
#include <cooperative_groups.h>
#include <cmath>
#include <cstring>
#include "context.h"

namespace cgrp = cooperative_groups;

namespace mg {

constexpr int CG_THREADS = 512;
constexpr int CG_WARPS = CG_THREADS / 32;
constexpr int MAX_PART = 2048;   // max CTAs in the cooperative grid

struct Cg2dOut {
  double firstResidual, minResidualSq, lastResidual, sumRHS, rhsMax;
  int numIters, nIterMin, pad;
};

struct Cg2dArgs {
  int sNx, sNy, OLx, OLy, PX, nTiles;
  size_t slab;
  int nIB, nJB, RY, nItems;
  const double *aW, *aS, *aC, *pW, *pS, *pC;
  double *b, *x;
  double *r[2], *s[2], *q, *z, *xmin, *v;   // v: extra vector of the SR variant
  const int *pushTab;
  double *partials;      // [4][MAX_PART]
  double *resid;         // per-iteration residual (sqrt(err_sq)), maxIters entries
  Cg2dOut *out;
  double cg2dNorm, tolSq;
  int normaliseRHS, maxIters, nIterMinIn;
};

struct Cg2dWs {
  double *r[2] = {nullptr, nullptr}, *s[2] = {nullptr, nullptr}, *q = nullptr, *z = nullptr, *xmin = nullptr,
         *v = nullptr;
  double *partials = nullptr, *resid = nullptr;
  Cg2dOut *out = nullptr;
  int residCap = 0;
  int maxBlocks = 0, maxBlocksSR = 0;
  double sumRHS = 0, rhsMax = 0;
  std::vector<double> residHost;
  int lastIters = 0;
};

void cg2d_free_workspace() {
  Ctx &c = ctx();
  if (!c.cg2d) return;
  Cg2dWs *w = c.cg2d;
  for (double *p : {w->r[0], w->r[1], w->s[0], w->s[1], w->q, w->z, w->xmin, w->v, w->partials, w->resid})
    if (p) cudaFree(p);
  if (w->out) cudaFree(w->out);
  delete w;
  c.cg2d = nullptr;
}

// ---- device helpers ---------------------------------------------------------------------

struct Item {
  int tile, i, j0, j1;    // Fortran indices; column i, rows j0..j1
  size_t base;            // flat index of (i, j0) in a tile2d array
  bool active;
};

__device__ __forceinline__ Item decode_item(const Cg2dArgs &a, int item, int lane) {
  Item it;
  int perTile = a.nIB * a.nJB;
  it.tile = item / perTile;
  int rem = item - it.tile * perTile;
  int jb = rem / a.nIB, ib = rem - jb * a.nIB;
  it.i = 1 + ib * 32 + lane;
  it.j0 = 1 + jb * a.RY;
  it.j1 = min(it.j0 + a.RY - 1, a.sNy);
  it.active = it.i <= a.sNx;
  it.base = (size_t)(it.i + a.OLx - 1) + (size_t)a.PX * (size_t)(it.j0 + a.OLy - 1) + a.slab * (size_t)it.tile;
  return it;
}

// Mirror an edge value into the halo cell(s) of the neighbouring tile(s).
__device__ __forceinline__ void push2(const Cg2dArgs &a, const Item &it, int j, double *f0, double v0,
                                      double *f1, double v1) {
  const int per = 2 * a.sNy + 2 * a.sNx;
  const int *t = a.pushTab + (size_t)per * it.tile;
  if (it.i == 1) { int d = t[j - 1]; f0[d] = v0; f1[d] = v1; }
  if (it.i == a.sNx) { int d = t[a.sNy + j - 1]; f0[d] = v0; f1[d] = v1; }
  if (j == 1) { int d = t[2 * a.sNy + it.i - 1]; f0[d] = v0; f1[d] = v1; }
  if (j == a.sNy) { int d = t[2 * a.sNy + a.sNx + it.i - 1]; f0[d] = v0; f1[d] = v1; }
}
__device__ __forceinline__ void push1(const Cg2dArgs &a, const Item &it, int j, double *f0, double v0) {
  const int per = 2 * a.sNy + 2 * a.sNx;
  const int *t = a.pushTab + (size_t)per * it.tile;
  if (it.i == 1) f0[t[j - 1]] = v0;
  if (it.i == a.sNx) f0[t[a.sNy + j - 1]] = v0;
  if (j == 1) f0[t[2 * a.sNy + it.i - 1]] = v0;
  if (j == a.sNy) f0[t[2 * a.sNy + a.sNx + it.i - 1]] = v0;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// CTA-level reduction of up to 3 sums, written to partials[k][blockIdx.x].
template <int N, bool MAXOP>
__device__ __forceinline__ void block_partials(const Cg2dArgs &a, double (&v)[N], double *sm) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < N; k++) v[k] = MAXOP ? warp_max(v[k]) : warp_sum(v[k]);
  if (lane == 0)
#pragma unroll
    for (int k = 0; k < N; k++) sm[k * CG_WARPS + w] = v[k];
  __syncthreads();
  if (threadIdx.x < N) {
    double t = sm[threadIdx.x * CG_WARPS];
    for (int i = 1; i < CG_WARPS; i++) t = MAXOP ? fmax(t, sm[threadIdx.x * CG_WARPS + i]) : t + sm[threadIdx.x * CG_WARPS + i];
    a.partials[threadIdx.x * MAX_PART + blockIdx.x] = t;
  }
  __syncthreads();
}

// After a grid barrier: every CTA forms the same ordered sum over the per-CTA partials.
template <int N, bool MAXOP>
__device__ __forceinline__ void grid_totals(const Cg2dArgs &a, double (&tot)[N], double *sm) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (w < N) {
    double t = MAXOP ? 0.0 : 0.0;
    for (int i = lane; i < (int)gridDim.x; i += 32) {
      double p = __ldcg(&a.partials[w * MAX_PART + i]);
      t = MAXOP ? fmax(t, p) : t + p;
    }
    t = MAXOP ? warp_max(t) : warp_sum(t);
    if (lane == 0) sm[w] = t;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < N; k++) tot[k] = sm[k];
  __syncthreads();
}

// ---- phases -----------------------------------------------------------------------------

// cg2d.F:105-133 first half: b *= cg2dNorm, rhsMax.
__device__ void phase_scale_b(const Cg2dArgs &a, double *sm) {
  const int lane = threadIdx.x & 31;
  const int gw = blockIdx.x * CG_WARPS + (threadIdx.x >> 5), nw = gridDim.x * CG_WARPS;
  double acc[1] = {0.0};
  for (int item = gw; item < a.nItems; item += nw) {
    Item it = decode_item(a, item, lane);
    if (!it.active) continue;
    size_t idx = it.base;
    for (int j = it.j0; j <= it.j1; j++, idx += a.PX) {
      double bv = a.b[idx] * a.cg2dNorm;
      a.b[idx] = bv;
      acc[0] = fmax(fabs(bv), acc[0]);
    }
  }
  block_partials<1, true>(a, acc, sm);
}

// cg2d.F:117-133 second half: b *= rhsNorm, x *= rhsNorm; ring-1 halo of x (EXCH_XY_RL :136).
__device__ void phase_normalise(const Cg2dArgs &a, double rhsNorm, bool normalise) {
  const int lane = threadIdx.x & 31;
  const int gw = blockIdx.x * CG_WARPS + (threadIdx.x >> 5), nw = gridDim.x * CG_WARPS;
  for (int item = gw; item < a.nItems; item += nw) {
    Item it = decode_item(a, item, lane);
    if (!it.active) continue;
    size_t idx = it.base;
    for (int j = it.j0; j <= it.j1; j++, idx += a.PX) {
      double xv = a.x[idx];
      if (normalise) {
        a.b[idx] = a.b[idx] * rhsNorm;
        xv = xv * rhsNorm;
        a.x[idx] = xv;
      }
      push1(a, it, j, a.x, xv);
    }
  }
}

// cg2d.F:141-178: r = b - A x, err = <r,r>, sumRHS = sum b, optional x_min = x.
__device__ void phase_residual(const Cg2dArgs &a, double *sm) {
  const int lane = threadIdx.x & 31;
  const int gw = blockIdx.x * CG_WARPS + (threadIdx.x >> 5), nw = gridDim.x * CG_WARPS;
  double acc[2] = {0.0, 0.0};
  double *r = a.r[0];
  for (int item = gw; item < a.nItems; item += nw) {
    Item it = decode_item(a, item, lane);
    if (!it.active) continue;
    size_t idx = it.base;
    for (int j = it.j0; j <= it.j1; j++, idx += a.PX) {
      double bv = a.b[idx];
      double rv = bv - (a.aW[idx] * a.x[idx - 1] + a.aW[idx + 1] * a.x[idx + 1] + a.aS[idx] * a.x[idx - a.PX] +
                        a.aS[idx + a.PX] * a.x[idx + a.PX] + a.aC[idx] * a.x[idx]);
      r[idx] = rv;
      push1(a, it, j, r, rv);
      if (a.nIterMinIn >= 0) a.xmin[idx] = a.x[idx];
      acc[0] += rv * rv;
      acc[1] += bv;
    }
  }
  block_partials<2, false>(a, acc, sm);
}

// Phase CA (cg2d.F:305-321 fused with :217-236 of the next iteration).
//   first == true : z = M r only (r already in rOld; q, s are zero)
__device__ void phase_ca(const Cg2dArgs &a, const double *rOld, double *rNew, const double *sCur, double alpha,
                         bool first, double *sm) {
  const int lane = threadIdx.x & 31;
  const int gw = blockIdx.x * CG_WARPS + (threadIdx.x >> 5), nw = gridDim.x * CG_WARPS;
  const int PX = a.PX;
  double acc[2] = {0.0, 0.0};   // err, eta
  for (int item = gw; item < a.nItems; item += nw) {
    Item it = decode_item(a, item, lane);
    if (!it.active) continue;
    size_t idx = it.base;
    double rS, rC;
    if (first) {
      rS = rOld[idx - PX];
      rC = rOld[idx];
    } else {
      rS = rOld[idx - PX] - alpha * a.q[idx - PX];
      rC = rOld[idx] - alpha * a.q[idx];
    }
    for (int j = it.j0; j <= it.j1; j++, idx += PX) {
      double rN, rW, rE;
      if (first) {
        rN = rOld[idx + PX]; rW = rOld[idx - 1]; rE = rOld[idx + 1];
      } else {
        rN = rOld[idx + PX] - alpha * a.q[idx + PX];
        rW = rOld[idx - 1] - alpha * a.q[idx - 1];
        rE = rOld[idx + 1] - alpha * a.q[idx + 1];
        a.x[idx] = a.x[idx] + alpha * sCur[idx];
      }
      double zv = a.pC[idx] * rC + a.pW[idx] * rW + a.pW[idx + 1] * rE + a.pS[idx] * rS + a.pS[idx + PX] * rN;
      rNew[idx] = rC;
      a.z[idx] = zv;
      push2(a, it, j, rNew, rC, a.z, zv);
      acc[0] += rC * rC;
      acc[1] += zv * rC;
      rS = rC;
      rC = rN;
    }
  }
  block_partials<2, false>(a, acc, sm);
}

// Phase B (cg2d.F:252-289): s = z + beta s on the stencil, q = A s, partial <s,q>.
// saveMin: store the lowest-residual solution first (cg2d.F:338-351).
__device__ void phase_b(const Cg2dArgs &a, const double *sOld, double *sNew, double beta, bool saveMin, double *sm) {
  const int lane = threadIdx.x & 31;
  const int gw = blockIdx.x * CG_WARPS + (threadIdx.x >> 5), nw = gridDim.x * CG_WARPS;
  const int PX = a.PX;
  double acc[1] = {0.0};
  for (int item = gw; item < a.nItems; item += nw) {
    Item it = decode_item(a, item, lane);
    if (!it.active) continue;
    size_t idx = it.base;
    double tS = a.z[idx - PX] + beta * sOld[idx - PX];
    double tC = a.z[idx] + beta * sOld[idx];
    for (int j = it.j0; j <= it.j1; j++, idx += PX) {
      double tN = a.z[idx + PX] + beta * sOld[idx + PX];
      double tW = a.z[idx - 1] + beta * sOld[idx - 1];
      double tE = a.z[idx + 1] + beta * sOld[idx + 1];
      double qv = a.aW[idx] * tW + a.aW[idx + 1] * tE + a.aS[idx] * tS + a.aS[idx + PX] * tN + a.aC[idx] * tC;
      sNew[idx] = tC;
      a.q[idx] = qv;
      push2(a, it, j, sNew, tC, a.q, qv);
      if (saveMin) a.xmin[idx] = a.x[idx];
      acc[0] += tC * qv;
      tS = tC;
      tC = tN;
    }
  }
  block_partials<1, false>(a, acc, sm);
}

// cg2d.F:358-384: restore the min-residual solution, un-normalise.
__device__ void phase_finish(const Cg2dArgs &a, bool useMin, bool saveMinPending, double rhsNorm) {
  const int lane = threadIdx.x & 31;
  const int gw = blockIdx.x * CG_WARPS + (threadIdx.x >> 5), nw = gridDim.x * CG_WARPS;
  (void)saveMinPending;
  for (int item = gw; item < a.nItems; item += nw) {
    Item it = decode_item(a, item, lane);
    if (!it.active) continue;
    size_t idx = it.base;
    for (int j = it.j0; j <= it.j1; j++, idx += a.PX) {
      double xv = useMin ? a.xmin[idx] : a.x[idx];
      if (a.normaliseRHS) xv = xv / rhsNorm;
      a.x[idx] = xv;
    }
  }
}

__global__ void __launch_bounds__(CG_THREADS) cg2d_kernel(Cg2dArgs a) {
  cgrp::grid_group grid = cgrp::this_grid();
  __shared__ double sm[4 * CG_WARPS];
  double t1[1], t2[2];

  phase_scale_b(a, sm);
  grid.sync();
  grid_totals<1, true>(a, t1, sm);
  const double rhsMax = t1[0];
  double rhsNorm = 1.0;
  if (a.normaliseRHS && rhsMax != 0.0) rhsNorm = 1.0 / rhsMax;
  phase_normalise(a, rhsNorm, a.normaliseRHS != 0);
  grid.sync();
  phase_residual(a, sm);
  grid.sync();
  grid_totals<2, false>(a, t2, sm);
  double err_sq = t2[0];
  const double sumRHS = t2[1];
  const double firstResidual = sqrt(err_sq);
  double minResidualSq = -1.0;
  int nIterMin = a.nIterMinIn;
  if (nIterMin >= 0) { nIterMin = 0; minResidualSq = err_sq; }
  int actualIts = 0;
  double eta_qrNM1 = 1.0;
  int cur = 0;              // r[cur] / s[cur] hold the current iterate
  bool saveMin = false;     // x_min = x is pending (done inside the next phase B)

  if (!(err_sq < a.tolSq)) {
    // z = M r for the first iteration
    phase_ca(a, a.r[0], a.r[1], a.s[0], 0.0, true, sm);
    cur = 1;   // r[1] now holds r (with halos); s[0] is the (zero) current s
    int scur = 0;
    grid.sync();
    grid_totals<2, false>(a, t2, sm);
    double eta_qrN = t2[1];
    for (int it2d = 1; it2d <= a.maxIters; it2d++) {
      const double cgBeta = eta_qrN / eta_qrNM1;
      eta_qrNM1 = eta_qrN;
      phase_b(a, a.s[scur], a.s[scur ^ 1], cgBeta, saveMin, sm);
      saveMin = false;
      scur ^= 1;
      grid.sync();
      grid_totals<1, false>(a, t1, sm);
      const double alpha = eta_qrN / t1[0];
      phase_ca(a, a.r[cur], a.r[cur ^ 1], a.s[scur], alpha, false, sm);
      cur ^= 1;
      grid.sync();
      grid_totals<2, false>(a, t2, sm);
      err_sq = t2[0];
      eta_qrN = t2[1];
      actualIts = it2d;
      if (blockIdx.x == 0 && threadIdx.x == 0) a.resid[it2d - 1] = sqrt(err_sq);
      if (err_sq < a.tolSq) break;
      if (err_sq < minResidualSq) {   // never true when nIterMin < 0 (minResidualSq = -1)
        minResidualSq = err_sq;
        nIterMin = it2d;
        saveMin = true;
      }
    }
  }
  // a pending x_min = x copy only matters if x_min is used, i.e. err_sq > minResidualSq, which
  // cannot hold for the iterate that set minResidualSq = err_sq; so it can be dropped.
  const bool useMin = (nIterMin >= 0 && err_sq > minResidualSq);
  phase_finish(a, useMin, saveMin, rhsNorm);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    a.out->firstResidual = firstResidual;
    a.out->minResidualSq = minResidualSq;
    a.out->lastResidual = sqrt(err_sq);
    a.out->sumRHS = sumRHS;
    a.out->rhsMax = rhsMax;
    a.out->numIters = actualIts;
    a.out->nIterMin = nIterMin;
  }
}

// ---- CG2D_SR (cg2d_sr.F): single-reduction recurrence -------------------------------------
// Phases per iteration (two barriers):
//   phase V : v = A y, partial <y,r>, <y,v>, <r,r>            (cg2d_sr.F:332-358)
//   phase U : s = y + beta s ; x += sigma s ; q = v + beta q ; r -= sigma q  (:390-405)
//   phase Y : y = M r                                           (:305-316)
__device__ void sr_phase_v(const Cg2dArgs &a, const double *y, const double *r, double *sm) {
  const int lane = threadIdx.x & 31;
  const int gw = blockIdx.x * CG_WARPS + (threadIdx.x >> 5), nw = gridDim.x * CG_WARPS;
  const int PX = a.PX;
  double acc[3] = {0.0, 0.0, 0.0};
  for (int item = gw; item < a.nItems; item += nw) {
    Item it = decode_item(a, item, lane);
    if (!it.active) continue;
    size_t idx = it.base;
    double yS = y[idx - PX], yC = y[idx];
    for (int j = it.j0; j <= it.j1; j++, idx += PX) {
      double yN = y[idx + PX];
      double vv = a.aW[idx] * y[idx - 1] + a.aW[idx + 1] * y[idx + 1] + a.aS[idx] * yS + a.aS[idx + PX] * yN +
                  a.aC[idx] * yC;
      double rv = r[idx];
      a.v[idx] = vv;
      acc[0] += yC * rv;
      acc[1] += yC * vv;
      acc[2] += rv * rv;
      yS = yC;
      yC = yN;
    }
  }
  block_partials<3, false>(a, acc, sm);
}

// y = M r (cg2d_sr.F:305-316); in the start-up iteration also s = y and eta = <y,r> (:220-242).
__device__ void sr_phase_y(const Cg2dArgs &a, const double *r, double *y, double *sCopy, double *sm, bool withEta) {
  const int lane = threadIdx.x & 31;
  const int gw = blockIdx.x * CG_WARPS + (threadIdx.x >> 5), nw = gridDim.x * CG_WARPS;
  const int PX = a.PX;
  double acc[1] = {0.0};
  for (int item = gw; item < a.nItems; item += nw) {
    Item it = decode_item(a, item, lane);
    if (!it.active) continue;
    size_t idx = it.base;
    double rS = r[idx - PX], rC = r[idx];
    for (int j = it.j0; j <= it.j1; j++, idx += PX) {
      double rN = r[idx + PX];
      double yv = a.pC[idx] * rC + a.pW[idx] * r[idx - 1] + a.pW[idx + 1] * r[idx + 1] + a.pS[idx] * rS +
                  a.pS[idx + PX] * rN;
      y[idx] = yv;
      if (sCopy) { sCopy[idx] = yv; push2(a, it, j, y, yv, sCopy, yv); }
      else push1(a, it, j, y, yv);
      acc[0] += yv * rC;
      rS = rC; rC = rN;
    }
  }
  if (withEta) block_partials<1, false>(a, acc, sm);
}

__device__ void sr_phase_as(const Cg2dArgs &a, const double *s, double *sm) {   // q = A s, <s,q>
  const int lane = threadIdx.x & 31;
  const int gw = blockIdx.x * CG_WARPS + (threadIdx.x >> 5), nw = gridDim.x * CG_WARPS;
  const int PX = a.PX;
  double acc[1] = {0.0};
  for (int item = gw; item < a.nItems; item += nw) {
    Item it = decode_item(a, item, lane);
    if (!it.active) continue;
    size_t idx = it.base;
    double sS = s[idx - PX], sC = s[idx];
    for (int j = it.j0; j <= it.j1; j++, idx += PX) {
      double sN = s[idx + PX];
      double qv = a.aW[idx] * s[idx - 1] + a.aW[idx + 1] * s[idx + 1] + a.aS[idx] * sS + a.aS[idx + PX] * sN +
                  a.aC[idx] * sC;
      a.q[idx] = qv;
      acc[0] += sC * qv;
      sS = sC; sC = sN;
    }
  }
  block_partials<1, false>(a, acc, sm);
}

// x += sigma s ; r -= sigma q with optional s = y + beta s, q = v + beta q first (in place,
// point-wise: no neighbours are read).  cg2d_sr.F:272-283 (startup) and :390-405.
__device__ void sr_phase_update(const Cg2dArgs &a, double *r, double beta, double sigma, bool startup, bool saveMin) {
  const int lane = threadIdx.x & 31;
  const int gw = blockIdx.x * CG_WARPS + (threadIdx.x >> 5), nw = gridDim.x * CG_WARPS;
  double *s = a.s[0];
  for (int item = gw; item < a.nItems; item += nw) {
    Item it = decode_item(a, item, lane);
    if (!it.active) continue;
    size_t idx = it.base;
    for (int j = it.j0; j <= it.j1; j++, idx += a.PX) {
      if (saveMin) a.xmin[idx] = a.x[idx];
      double sv = s[idx], qv = a.q[idx];
      if (!startup) {
        sv = a.z[idx] + beta * sv;
        qv = a.v[idx] + beta * qv;
        s[idx] = sv;
        a.q[idx] = qv;
      }
      a.x[idx] = a.x[idx] + sigma * sv;
      double rv = r[idx] - sigma * qv;
      r[idx] = rv;
      push1(a, it, j, r, rv);
    }
  }
}

__device__ void sr_phase_err(const Cg2dArgs &a, const double *r, double *sm) {
  const int lane = threadIdx.x & 31;
  const int gw = blockIdx.x * CG_WARPS + (threadIdx.x >> 5), nw = gridDim.x * CG_WARPS;
  double acc[1] = {0.0};
  for (int item = gw; item < a.nItems; item += nw) {
    Item it = decode_item(a, item, lane);
    if (!it.active) continue;
    size_t idx = it.base;
    for (int j = it.j0; j <= it.j1; j++, idx += a.PX) acc[0] += r[idx] * r[idx];
  }
  block_partials<1, false>(a, acc, sm);
}

__global__ void __launch_bounds__(CG_THREADS) cg2d_sr_kernel(Cg2dArgs a) {
  cgrp::grid_group grid = cgrp::this_grid();
  __shared__ double sm[4 * CG_WARPS];
  double t1[1], t2[2], t3[3];
  double *r = a.r[0], *y = a.z, *s = a.s[0];

  phase_scale_b(a, sm);
  grid.sync();
  grid_totals<1, true>(a, t1, sm);
  const double rhsMax = t1[0];
  double rhsNorm = 1.0;
  if (a.normaliseRHS && rhsMax != 0.0) rhsNorm = 1.0 / rhsMax;
  phase_normalise(a, rhsNorm, a.normaliseRHS != 0);
  grid.sync();
  phase_residual(a, sm);
  grid.sync();
  grid_totals<2, false>(a, t2, sm);
  double err_sq = t2[0];
  const double sumRHS = t2[1];
  const double firstResidual = sqrt(err_sq);
  double minResidualSq = -1.0;
  int nIterMin = a.nIterMinIn;
  if (nIterMin >= 0) { nIterMin = 0; minResidualSq = err_sq; }
  int it2d = 0;
  bool saveMin = false;

  if (!(err_sq < a.tolSq)) {
    // start-up iteration, cg2d_sr.F:220-291
    sr_phase_y(a, r, y, s, sm, true);
    grid.sync();
    grid_totals<1, false>(a, t1, sm);
    double eta_qrN = t1[0];
    double eta_qrNM1 = eta_qrN;
    sr_phase_as(a, s, sm);
    grid.sync();
    grid_totals<1, false>(a, t1, sm);
    double alpha = t1[0];
    double sigma = eta_qrN / alpha;
    sr_phase_update(a, r, 0.0, sigma, true, false);
    grid.sync();
    bool converged = false;
    for (it2d = 1; it2d <= a.maxIters - 1; it2d++) {
      sr_phase_y(a, r, y, nullptr, sm, false);
      grid.sync();
      sr_phase_v(a, y, r, sm);
      grid.sync();
      grid_totals<3, false>(a, t3, sm);
      eta_qrN = t3[0];
      const double delta = t3[1];
      err_sq = t3[2];
      if (blockIdx.x == 0 && threadIdx.x == 0) a.resid[it2d - 1] = sqrt(err_sq);
      if (err_sq < a.tolSq) { converged = true; break; }
      saveMin = false;
      if (err_sq < minResidualSq) { minResidualSq = err_sq; nIterMin = it2d; saveMin = true; }
      const double cgBeta = eta_qrN / eta_qrNM1;
      eta_qrNM1 = eta_qrN;
      alpha = delta - (cgBeta * cgBeta) * alpha;
      sigma = eta_qrN / alpha;
      sr_phase_update(a, r, cgBeta, sigma, false, saveMin);
      grid.sync();
    }
    if (!converged) {
      sr_phase_err(a, r, sm);
      grid.sync();
      grid_totals<1, false>(a, t1, sm);
      err_sq = t1[0];
    }
  }
  const bool useMin = (nIterMin >= 0 && err_sq > minResidualSq);
  phase_finish(a, useMin, false, rhsNorm);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    a.out->firstResidual = firstResidual;
    a.out->minResidualSq = minResidualSq;
    a.out->lastResidual = sqrt(err_sq);
    a.out->sumRHS = sumRHS;
    a.out->rhsMax = rhsMax;
    a.out->numIters = it2d;
    a.out->nIterMin = nIterMin;
  }
}

// ---- host side ------------------------------------------------------------------------------

static bool ensure_ws(int maxIters) {
  Ctx &c = ctx();
  if (!c.cg2d) {
    Cg2dWs *w = new Cg2dWs();
    c.cg2d = w;
    size_t bytes = c.g.n2 * sizeof(double);
    for (double **p : {&w->r[0], &w->r[1], &w->s[0], &w->s[1], &w->q, &w->z, &w->xmin, &w->v}) {
      MG_CUDA(cudaMalloc(p, bytes));
      MG_CUDA(cudaMemsetAsync(*p, 0, bytes, c.stream));
    }
    MG_CUDA(cudaMalloc(&w->partials, 4 * MAX_PART * sizeof(double)));
    MG_CUDA(cudaMalloc(&w->out, sizeof(Cg2dOut)));
    int nb = 0;
    MG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, cg2d_kernel, CG_THREADS, 0));
    w->maxBlocks = nb * c.numSMs;
    MG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, cg2d_sr_kernel, CG_THREADS, 0));
    w->maxBlocksSR = nb * c.numSMs;
  }
  Cg2dWs *w = c.cg2d;
  if (w->residCap < maxIters) {
    if (w->resid) cudaFree(w->resid);
    MG_CUDA(cudaMalloc(&w->resid, (size_t)(maxIters > 0 ? maxIters : 1) * sizeof(double)));
    w->residCap = maxIters > 0 ? maxIters : 1;
  }
  return true;
}

static bool cg2d_run(bool sr, double *cg2d_b, double *cg2d_x, double *firstResidual, double *minResidualSq,
                     double *lastResidual, int *numIters, int *nIterMin) {
  Ctx &c = ctx();
  if (!c.ready) return fail(30, "mitgcm_b200_init_ not called");
  if (!ensure_ws(*numIters)) return false;
  Cg2dWs *w = c.cg2d;
  const Geom &g = c.g;
  Cg2dArgs a;
  a.sNx = g.sNx; a.sNy = g.sNy; a.OLx = g.OLx; a.OLy = g.OLy; a.PX = g.PX; a.nTiles = g.nTiles;
  a.slab = g.slab;
  a.aW = field(MG_AW2D); a.aS = field(MG_AS2D); a.aC = field(MG_AC2D);
  a.pW = field(MG_PW); a.pS = field(MG_PS); a.pC = field(MG_PC);
  a.b = to_device(cg2d_b, g.n2, 0, true);
  a.x = to_device(cg2d_x, g.n2, 1, true);
  if (!a.b || !a.x || !a.aW || !a.pC) return false;
  a.r[0] = w->r[0]; a.r[1] = w->r[1]; a.s[0] = w->s[0]; a.s[1] = w->s[1];
  a.q = w->q; a.z = w->z; a.xmin = w->xmin; a.v = w->v;
  a.pushTab = c.pushTab; a.partials = w->partials; a.resid = w->resid; a.out = w->out;
  a.cg2dNorm = c.p.D(MP_CG2DNORM); a.tolSq = c.p.D(MP_CG2DTOLERANCE_SQ);
  a.normaliseRHS = c.p.I(MI_CG2DNORMALISERHS);
  a.maxIters = *numIters; a.nIterMinIn = *nIterMin;
  // work decomposition: warp items of 32 columns x RY rows
  a.nIB = (g.sNx + 31) / 32;
  const int maxBlocks = std::min(sr ? w->maxBlocksSR : w->maxBlocks, MAX_PART);
  const long totalWarps = (long)maxBlocks * CG_WARPS;
  long rows = (long)g.sNy * a.nIB * g.nTiles;     // warp-rows of work
  int RY = (int)std::min<long>(16, std::max<long>(1, rows / (2 * totalWarps)));
  a.RY = RY;
  a.nJB = (g.sNy + RY - 1) / RY;
  a.nItems = g.nTiles * a.nIB * a.nJB;
  int blocks = std::min(maxBlocks, (a.nItems + CG_WARPS - 1) / CG_WARPS);
  if (blocks < 1) blocks = 1;
  // zero-initialised work arrays incl. ring 0 / sN+1 (cg2d.F:142-147)
  size_t bytes = g.n2 * sizeof(double);
  for (double *p : {w->r[0], w->r[1], w->s[0], w->s[1], w->q, w->z, w->v}) MG_CUDA(cudaMemsetAsync(p, 0, bytes, c.stream));
  void *args[] = {&a};
  MG_CUDA(cudaLaunchCooperativeKernel(sr ? (void *)cg2d_sr_kernel : (void *)cg2d_kernel, dim3(blocks), dim3(CG_THREADS),
                                      args, 0, c.stream));
  Cg2dOut out;
  MG_CUDA(cudaMemcpyAsync(&out, w->out, sizeof(out), cudaMemcpyDeviceToHost, c.stream));
  if (!from_device(cg2d_b, a.b, g.n2)) return false;
  if (!from_device(cg2d_x, a.x, g.n2)) return false;
  MG_CUDA(cudaStreamSynchronize(c.stream));
  *firstResidual = out.firstResidual;
  *minResidualSq = out.minResidualSq;
  *lastResidual = out.lastResidual;
  *numIters = out.numIters;
  *nIterMin = out.nIterMin;
  w->sumRHS = out.sumRHS;
  w->rhsMax = out.rhsMax;
  w->lastIters = out.numIters;
  return true;
}

}  // namespace mg

extern "C" {

void cg2d_b200_(double *cg2d_b, double *cg2d_x, double *firstResidual, double *minResidualSq,
                double *lastResidual, int *numIters, int *nIterMin, const int *myThid) {
  (void)myThid;
  mg::ctx().lastError = 0;
  mg::cg2d_run(false, cg2d_b, cg2d_x, firstResidual, minResidualSq, lastResidual, numIters, nIterMin);
}

void cg2d_sr_b200_(double *cg2d_b, double *cg2d_x, double *firstResidual, double *minResidualSq,
                   double *lastResidual, int *numIters, int *nIterMin, const int *myThid) {
  (void)myThid;
  mg::ctx().lastError = 0;
  mg::cg2d_run(true, cg2d_b, cg2d_x, firstResidual, minResidualSq, lastResidual, numIters, nIterMin);
}

void mitgcm_b200_cg2d_stats_(double *sumRHS, double *rhsMax) {
  mg::Cg2dWs *w = mg::ctx().cg2d;
  *sumRHS = w ? w->sumRHS : 0.0;
  *rhsMax = w ? w->rhsMax : 0.0;
}

void mitgcm_b200_cg2d_residuals_(double *resid, const int *n) {
  mg::Ctx &c = mg::ctx();
  mg::Cg2dWs *w = c.cg2d;
  if (!w || *n <= 0) return;
  int m = *n < w->lastIters ? *n : w->lastIters;
  if (m > 0) cudaMemcpy(resid, w->resid, (size_t)m * sizeof(double), cudaMemcpyDeviceToHost);
}

}  // extern "C"
