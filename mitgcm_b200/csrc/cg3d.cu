// cg3d.cu -- CG3D drop-in (model/src/cg3d.F:13-545; caller solve_for_pressure.F:427): the preconditioned
// conjugate-gradient solver of the 3-D (non-hydrostatic) pressure equation, SURVEY.md section 8(f) rank 4.
// A x = b with the 7-point operator {aW3d(i), aW3d(i+1), aS3d(j), aS3d(j+1), aV3d(k), aV3d(k+1), aC3d}; the
// preconditioner is the vertical tridiagonal solve per column (zMC, zML, zMU of INI_CG3D), applied on the ring
// 0..sN+1 so that no exchange of the search direction is needed.
// One persistent cooperative kernel per solve (as cg2d.cu).
//
// cg3d_fused_kernel (default): TWO grid barriers per iteration, both phases are column marches
//   SA : s' = z + beta s recomputed at the 7 points of the stencil from z and the previous s (s is double-
//        buffered, so no sweep of its own and no barrier between "S" and "A"), q = A s', <s',q>; the column keeps
//        s'(k-1), s'(k), s'(k+1) in registers; edge columns also store s' into the ring cells next to them and
//        push q into the neighbouring tile's ring (width-1 push table shared with CG2D)
//        R{z, s, aW, aS, aV, aC} W{s', q}                                                       8 words
//   UM : the update of THIS iteration and the preconditioner of the NEXT in one march down and up every ring
//        column: r -= alpha q, <r,r>, z' = forward elimination, then z = back substitution, eta = <z,r>.  Ring
//        columns update their copy of r from the pushed q (same numbers as the owner: no exchange of r).
//        x += alpha s is applied two iterations at a time, x = (x + a1 s1) + a2 s2 (the reference's order of
//        additions; both s buffers are still in memory), so odd iterations touch neither x nor s.
//        R{r, q, zMC, zML | z', zMU, r} W{r, z' | z}  + x: 4 words every second iteration      10 + 2 words
//        (z' of the first C3_SMEM_LEVELS levels waits in shared memory instead of HBM)
//   = about 19.5 words = 156 B per cell and iteration (the back substitution's re-reads included; ncu) against
//   21 + 2 re-read for the four-sweep form, and 2 instead of 4 grid barriers.  Measured at 1024^2 x 50 on B200:
//   1525 us per iteration against 2131 (profiles/r02_cg3d_fused_sweeps.log).
//
// cg3d_kernel (MITGCM_B200_CG3D_UNFUSED=1): the four-sweep form, four grid barriers per iteration
//   M : q = M r down and up every column of the ring, eta = <q,r>            R{r,zMC,zML,zMU} W{q}
//   S : s = q + beta s on the ring                                           R{q,s} W{s}
//   A : q = A s, <s,q>                                                       R{s x7,aW,aS,aV,aC} W{q}
//   U : x += alpha s, r -= alpha q, <r,r>, edge values of r pushed into the neighbours' ring
//       (EXCH_S3D_RL( cg3d_r, Nr )) through the width-1 push table shared with CG2D          R{x,s,r,q} W{x,r}
// Both evaluate every expression of cg3d.F in its order (-fmad=false): x, r, s, q are the same numbers given the
// same alpha and beta; the dot products are summed lane -> warp shuffle -> CTA -> ordered sum over CTAs
// (deterministic run to run; the order differs from the reference's tile-ordered sum, as in CG2D).  Convergence
// is decided on the device.  Yardstick for "of peak": the four-sweep form's 21 words = 168 B per cell and iteration.
// Single rank, select_rStar = 0 (no surface term), as the oracle.
#include <cooperative_groups.h>
#include <algorithm>
#include <cstdlib>
#include "context.h"

namespace cg = cooperative_groups;

namespace mg {

constexpr int C3_THREADS = 256, C3_WARPS = C3_THREADS / 32, C3_MAXB = 1184;

struct Cg3dOut {
  double firstResidual, lastResidual, sumRHS, rhsMax;
  int numIters;
  unsigned long long nsSA, nsUM;     // fused kernel: time CTA 0 spent in the two phases (barriers included)
};

struct Cg3dArgs {
  int sNx, sNy, OLx, OLy, PX, PY, Nr, nTiles;
  size_t slab;
  const double *aW, *aS, *aV, *aC, *zMC, *zML, *zMU, *maskC;
  double *b, *x, *r, *q, *s;
  double *z, *s1;            // fused kernel only: z = M r (q holds A s), second buffer of s
  int nSm;                   // fused kernel only: levels of z' a CTA keeps in shared memory (nSm * C3_THREADS doubles)
  const int *pushTab;
  double *partials;          // [3][C3_MAXB]
  Cg3dOut *out;
  double cg3dNorm, tolSq;
  int normaliseRHS, maxIters;
};

__device__ __forceinline__ double w_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double w_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// CTA partial -> partials[slot][blockIdx.x]
template <bool MAXOP>
__device__ void put_partial(const Cg3dArgs &a, int slot, double v, double *sm) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  v = MAXOP ? w_max(v) : w_sum(v);
  if (lane == 0) sm[w] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = MAXOP ? 0. : 0.;
    for (int i = 0; i < C3_WARPS; i++) t = MAXOP ? fmax(t, sm[i]) : t + sm[i];
    a.partials[slot * C3_MAXB + blockIdx.x] = t;
  }
  __syncthreads();
}
// after a grid barrier: every CTA forms the same ordered total
template <bool MAXOP>
__device__ double get_total(const Cg3dArgs &a, int slot, double *sm) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (w == 0) {
    double t = 0.;
    for (int i = lane; i < (int)gridDim.x; i += 32) {
      const double p = __ldcg(&a.partials[slot * C3_MAXB + i]);
      t = MAXOP ? fmax(t, p) : t + p;
    }
    t = MAXOP ? w_max(t) : w_sum(t);
    if (lane == 0) sm[C3_WARPS] = t;
  }
  __syncthreads();
  const double t = sm[C3_WARPS];
  __syncthreads();
  return t;
}

// interior cell n (0 .. nTiles*Nr*sNy*sNx) -> i, j, k, flat index
struct Cell { int i, j, k, tile; size_t idx; };
__device__ __forceinline__ Cell cell_of(const Cg3dArgs &a, size_t n) {
  Cell c;
  c.i = 1 + (int)(n % a.sNx); n /= a.sNx;
  c.j = 1 + (int)(n % a.sNy); n /= a.sNy;
  c.k = 1 + (int)(n % a.Nr);
  c.tile = (int)(n / a.Nr);
  c.idx = (size_t)(c.i + a.OLx - 1) + (size_t)a.PX * (size_t)(c.j + a.OLy - 1) + a.slab * ((size_t)(c.k - 1) + (size_t)a.Nr * c.tile);
  return c;
}
// mirror an edge value of a tile3d field into the ring cell(s) of the neighbouring tile(s)
__device__ __forceinline__ void push3(const Cg3dArgs &a, const Cell &c, double *f, double v) {
  const int per = 2 * a.sNy + 2 * a.sNx;
  const int *t = a.pushTab + (size_t)per * c.tile;
  auto put = [&](int enc) {
    const size_t d = (size_t)(enc & 0x0FFFFFFF), dt = d / a.slab, cell = d - dt * a.slab;
    f[cell + a.slab * ((size_t)(c.k - 1) + (size_t)a.Nr * dt)] = v;
  };
  if (c.i == 1) put(t[c.j - 1]);
  if (c.i == a.sNx) put(t[a.sNy + c.j - 1]);
  if (c.j == 1) put(t[2 * a.sNy + c.i - 1]);
  if (c.j == a.sNy) put(t[2 * a.sNy + a.sNx + c.i - 1]);
}

// cg3d.F:121-230: b *= cg3dNorm*maskC, rhsMax, optional normalisation of b and x, _EXCH_XYZ_RL( cg3d_x ) as pushes,
// r = b - A x with its edge values pushed into the ring, err = <r,r>, sumRHS.  Shared by both kernels.
struct Cg3dStart { double rhsMax, rhsNorm, err_sq, sumRHS; };
__device__ Cg3dStart cg3d_start(const Cg3dArgs &a, cg::grid_group &grid, double *sm, size_t tid, size_t nthr, size_t nInt) {
  const size_t PX = a.PX, slab = a.slab;
  Cg3dStart o;
  double acc = 0.;
  for (size_t n = tid; n < nInt; n += nthr) {
    const Cell c = cell_of(a, n);
    const double bv = a.b[c.idx] * a.cg3dNorm * a.maskC[c.idx];
    a.b[c.idx] = bv;
    acc = fmax(fabs(bv), acc);
  }
  put_partial<true>(a, 0, acc, sm);
  grid.sync();
  o.rhsMax = get_total<true>(a, 0, sm);
  o.rhsNorm = 1.;
  if (a.normaliseRHS) {     // :135-153
    if (o.rhsMax != 0.) o.rhsNorm = 1. / o.rhsMax;
    for (size_t n = tid; n < nInt; n += nthr) {
      const Cell c = cell_of(a, n);
      a.b[c.idx] = a.b[c.idx] * o.rhsNorm;
      const double xv = a.x[c.idx] * o.rhsNorm;
      a.x[c.idx] = xv;
      push3(a, c, a.x, xv);       // _EXCH_XYZ_RL( cg3d_x ): the ring is all the residual reads
    }
  } else {
    for (size_t n = tid; n < nInt; n += nthr) {
      const Cell c = cell_of(a, n);
      push3(a, c, a.x, a.x[c.idx]);
    }
  }
  grid.sync();
  // :160-230 r = b - A x, err, sumRHS (s is zero-initialised by the host)
  double e = 0., sb = 0.;
  for (size_t n = tid; n < nInt; n += nthr) {
    const Cell c = cell_of(a, n);
    const size_t q = c.idx;
    const int km1 = max(c.k - 1, 1), kp1 = min(c.k + 1, a.Nr);
    const double maskM1 = c.k == 1 ? 0. : 1., maskP1 = c.k == a.Nr ? 0. : 1.;
    const size_t qm = q - slab * (size_t)(c.k - km1), qp = q + slab * (size_t)(kp1 - c.k);
    const double bv = a.b[q];
    const double rv = bv - (0. + a.aW[q] * a.x[q - 1] + a.aW[q + 1] * a.x[q + 1] + a.aS[q] * a.x[q - PX] + a.aS[q + PX] * a.x[q + PX] +
                            a.aV[q] * a.x[qm] * maskM1 + a.aV[qp] * a.x[qp] * maskP1 + a.aC[q] * a.x[q]);
    a.r[q] = rv;
    push3(a, c, a.r, rv);
    e += rv * rv;
    sb += bv;
  }
  put_partial<false>(a, 0, e, sm);
  put_partial<false>(a, 1, sb, sm);
  grid.sync();
  o.err_sq = get_total<false>(a, 0, sm);
  o.sumRHS = get_total<false>(a, 1, sm);
  return o;
}

// :481-495 un-normalise x; results
__device__ void cg3d_finish(const Cg3dArgs &a, cg::grid_group &grid, const Cg3dStart &st, double err_sq, int actualIts,
                            size_t tid, size_t nthr, size_t nInt) {
  if (a.normaliseRHS) {
    grid.sync();
    for (size_t n = tid; n < nInt; n += nthr) {
      const Cell c = cell_of(a, n);
      a.x[c.idx] = a.x[c.idx] / st.rhsNorm;
    }
  }
  if (tid == 0) {
    a.out->firstResidual = sqrt(st.err_sq);
    a.out->lastResidual = sqrt(err_sq);
    a.out->sumRHS = st.sumRHS;
    a.out->rhsMax = st.rhsMax;
    a.out->numIters = actualIts;
  }
}

// ---- four-sweep form ------------------------------------------------------------------------------------------
// Latency-bound streaming phases: occupancy is what counts (measured at 1024^2 x 50: 2 CTAs/SM 3955 us/iteration,
// 3: 2928, 4: 2521, 6: 2266, 8 (32 registers, 80 B spilled): 2149).
#ifndef C3_MINB
#define C3_MINB 8
#endif
__global__ void __launch_bounds__(C3_THREADS, C3_MINB) cg3d_kernel(Cg3dArgs a) {
  cg::grid_group grid = cg::this_grid();
  __shared__ double sm[C3_WARPS + 1];
  const size_t tid = (size_t)blockIdx.x * C3_THREADS + threadIdx.x, nthr = (size_t)gridDim.x * C3_THREADS;
  const size_t nInt = (size_t)a.nTiles * a.Nr * a.sNy * a.sNx;
  const int RX = a.sNx + 2, RY = a.sNy + 2;
  const size_t nCol = (size_t)a.nTiles * RY * RX;
  const size_t PX = a.PX, slab = a.slab;

  const Cg3dStart st = cg3d_start(a, grid, sm, tid, nthr, nInt);
  double err_sq = st.err_sq;
  int actualIts = 0;
  double eta_qrNM1 = 1.;
  if (!(err_sq < a.tolSq)) {
    for (int it3d = 1; it3d <= a.maxIters; it3d++) {
      grid.sync();      // the totals above were read by every CTA before the partials are overwritten
      // ---- M: q = M r on the ring (cg3d.F:258-303), eta = <q,r> over the interior
      double eta = 0.;
      for (size_t n = tid; n < nCol; n += nthr) {
        const int i = (int)(n % RX), j = (int)((n / RX) % RY), tile = (int)(n / ((size_t)RX * RY));
        const bool inner = i >= 1 && i <= a.sNx && j >= 1 && j <= a.sNy;
        size_t q = (size_t)(i + a.OLx - 1) + PX * (size_t)(j + a.OLy - 1) + slab * (size_t)a.Nr * tile;
        double qk = a.zMC[q] * a.r[q];
        a.q[q] = qk;
        for (int k = 2; k <= a.Nr; k++) {
          q += slab;
          qk = a.zMC[q] * (a.r[q] - a.zML[q] * qk);
          a.q[q] = qk;
        }
        if (inner) eta += qk * a.r[q];
        for (int k = a.Nr - 1; k >= 1; k--) {
          q -= slab;
          qk = a.q[q] - a.zMU[q] * qk;
          a.q[q] = qk;
          if (inner) eta += qk * a.r[q];
        }
      }
      put_partial<false>(a, 0, eta, sm);
      grid.sync();
      const double eta_qrN = get_total<false>(a, 0, sm);
      const double cgBeta = eta_qrN / eta_qrNM1;
      eta_qrNM1 = eta_qrN;
      // ---- S: s = q + beta s on the ring (:309-321)
      for (size_t n = tid; n < nCol * (size_t)a.Nr; n += nthr) {
        const int i = (int)(n % RX), j = (int)((n / RX) % RY);
        const size_t rest = n / ((size_t)RX * RY);
        const int k = 1 + (int)(rest % a.Nr), tile = (int)(rest / a.Nr);
        const size_t q = (size_t)(i + a.OLx - 1) + PX * (size_t)(j + a.OLy - 1) + slab * ((size_t)(k - 1) + (size_t)a.Nr * tile);
        a.s[q] = a.q[q] + cgBeta * a.s[q];
      }
      grid.sync();
      // ---- A: q = A s, <s,q> (:352-436)
      double al = 0.;
      for (size_t n = tid; n < nInt; n += nthr) {
        const Cell c = cell_of(a, n);
        const size_t q = c.idx;
        double v = a.aW[q] * a.s[q - 1] + a.aW[q + 1] * a.s[q + 1] + a.aS[q] * a.s[q - PX] + a.aS[q + PX] * a.s[q + PX];
        if (c.k > 1) v = v + a.aV[q] * a.s[q - slab];
        if (c.k < a.Nr) v = v + a.aV[q + slab] * a.s[q + slab];
        v = v + a.aC[q] * a.s[q];
        a.q[q] = v;
        al += a.s[q] * v;
      }
      put_partial<false>(a, 1, al, sm);
      grid.sync();
      const double alpha = eta_qrN / get_total<false>(a, 1, sm);
      // ---- U: x += alpha s, r -= alpha q, <r,r> (:441-462), EXCH_S3D_RL( cg3d_r ) as pushes
      double er = 0.;
      for (size_t n = tid; n < nInt; n += nthr) {
        const Cell c = cell_of(a, n);
        const size_t q = c.idx;
        a.x[q] = a.x[q] + alpha * a.s[q];
        const double rv = a.r[q] - alpha * a.q[q];
        a.r[q] = rv;
        push3(a, c, a.r, rv);
        er += rv * rv;
      }
      put_partial<false>(a, 2, er, sm);
      grid.sync();
      actualIts = it3d;
      err_sq = get_total<false>(a, 2, sm);
      if (err_sq < a.tolSq) break;
    }
  }
  cg3d_finish(a, grid, st, err_sq, actualIts, tid, nthr, nInt);
}

// ---- fused form -----------------------------------------------------------------------------------------------
#ifndef C3_DEFAULT_VARIANT
#define C3_DEFAULT_VARIANT 1
#endif
#ifndef C3_SMEM_LEVELS
#define C3_SMEM_LEVELS 32
#endif
#ifndef C3_PATCH_X
#define C3_PATCH_X 32
#endif
constexpr int C3_TX = C3_PATCH_X, C3_TY = C3_THREADS / C3_TX;      // SA: columns of one CTA = a C3_TX x C3_TY patch of a tile

// Streaming hints (-DC3_CS_HINTS=1, off by default): what a phase touches once (operators, q, x, s, the finished z)
// loaded / stored with .cs so that the L2 would keep what the back substitution reads again (z' and r of the columns
// in flight).  Measured slower at 1024^2 x 50: UM 789 -> 881 us per iteration; kept for A/B.
#ifndef C3_CS_HINTS
#define C3_CS_HINTS 0
#endif
__device__ __forceinline__ double ld_once(const double *p) { return C3_CS_HINTS ? __ldcs(p) : *p; }
__device__ __forceinline__ void st_once(double *p, double v) { if (C3_CS_HINTS) __stcs(p, v); else *p = v; }
__device__ __forceinline__ unsigned long long now_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// UM for one ring column.  MODE -1: preconditioner only (before the first iteration); 0: r only (x waits);
// 1: x += alpha s; 2: x = (x + alphaPrev sPrev) + alpha s.  U levels are loaded before they are used.
// The forward-elimination values z' of levels 0 .. nSm-1 wait for the back substitution in shared memory
// (zs[level * C3_THREADS], this thread's slots) instead of going to HBM and back; the last level's z' is final.
template <int U, int MODE>
__device__ __forceinline__ void um_column(const Cg3dArgs &a, size_t q0, bool inner, double alpha, double alphaPrev,
                                          const double *__restrict__ sCur, const double *__restrict__ sPrev,
                                          double *__restrict__ zs, int nSm, double &er, double &eta) {
  const size_t slab = a.slab;
  const int Nr = a.Nr;
  const double *__restrict__ zMC = a.zMC, *__restrict__ zML = a.zML, *__restrict__ zMU = a.zMU, *__restrict__ qA = a.q;
  double *__restrict__ r = a.r, *__restrict__ z = a.z, *__restrict__ x = a.x;
  const bool doX = MODE > 0 && inner;
  double zprev = 0., rlast = 0.;
  size_t q = q0;
  int k = 0;      // 0-based level of q
  auto level = [&](int kk, size_t idx, double rv, double qv, double mc, double ml, double xv, double sv, double pv) {
    if (MODE >= 0) {
      rv = rv - alpha * qv;                      // cg3d.F:449
      r[idx] = rv;
      if (inner) er += rv * rv;
      if (MODE == 1 && doX) st_once(&x[idx], xv + alpha * sv);                          // :447
      if (MODE == 2 && doX) st_once(&x[idx], (xv + alphaPrev * pv) + alpha * sv);       // two iterations of :447
    }
    const double zk = kk == 0 ? mc * rv : mc * (rv - ml * zprev);              // :262-283
    if (kk < nSm && kk != Nr - 1) zs[kk * C3_THREADS] = zk;
    else z[idx] = zk;
    zprev = zk;
    rlast = rv;
  };
  for (; k + U <= Nr; k += U, q += (size_t)U * slab) {
    double rv[U], qv[U], mc[U], ml[U], xv[U], sv[U], pv[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      const size_t idx = q + (size_t)u * slab;
      rv[u] = r[idx];
      qv[u] = MODE >= 0 ? ld_once(&qA[idx]) : 0.;
      mc[u] = ld_once(&zMC[idx]);
      ml[u] = ld_once(&zML[idx]);
      xv[u] = doX ? ld_once(&x[idx]) : 0.;
      sv[u] = doX ? ld_once(&sCur[idx]) : 0.;
      pv[u] = (MODE == 2 && doX) ? ld_once(&sPrev[idx]) : 0.;
    }
#pragma unroll
    for (int u = 0; u < U; u++) level(k + u, q + (size_t)u * slab, rv[u], qv[u], mc[u], ml[u], xv[u], sv[u], pv[u]);
  }
  for (; k < Nr; k++, q += slab)
    level(k, q, r[q], MODE >= 0 ? ld_once(&qA[q]) : 0., ld_once(&zMC[q]), ld_once(&zML[q]), doX ? ld_once(&x[q]) : 0.,
          doX ? ld_once(&sCur[q]) : 0., (MODE == 2 && doX) ? ld_once(&sPrev[q]) : 0.);
  // back substitution (:287-303); q is one level past the bottom
  if (inner) eta += zprev * rlast;
  k = Nr - 2;        // next level to finish, 0-based
  q -= 2 * slab;     // its index (unused when Nr == 1)
  for (; k - U + 1 >= 0; k -= U, q -= (size_t)U * slab) {
    double zv[U], mu[U], rv[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      const size_t idx = q - (size_t)u * slab;
      zv[u] = (k - u) < nSm ? zs[(k - u) * C3_THREADS] : z[idx];
      mu[u] = ld_once(&zMU[idx]);
      rv[u] = inner ? r[idx] : 0.;
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
      const double zk = zv[u] - mu[u] * zprev;
      st_once(&z[q - (size_t)u * slab], zk);
      zprev = zk;
      if (inner) eta += zk * rv[u];
    }
  }
  for (; k >= 0; k--, q -= slab) {
    const double zk = (k < nSm ? zs[k * C3_THREADS] : z[q]) - ld_once(&zMU[q]) * zprev;
    st_once(&z[q], zk);
    zprev = zk;
    if (inner) eta += zk * r[q];
  }
}

// UM over all ring columns: 256 consecutive columns of the linearised ring (i fastest) per CTA
template <int U, int MODE>
__device__ __forceinline__ void um_phase(const Cg3dArgs &a, double alpha, double alphaPrev, const double *sCur,
                                         const double *sPrev, double *zsm, double &er, double &eta) {
  const int RX = a.sNx + 2, RY = a.sNy + 2;
  const size_t nCol = (size_t)a.nTiles * RY * RX;
  for (size_t n = (size_t)blockIdx.x * C3_THREADS + threadIdx.x; n < nCol; n += (size_t)gridDim.x * C3_THREADS) {
    const int i = (int)(n % RX), j = (int)((n / RX) % RY), tile = (int)(n / ((size_t)RX * RY));
    const bool inner = i >= 1 && i <= a.sNx && j >= 1 && j <= a.sNy;
    const size_t q0 = (size_t)(i + a.OLx - 1) + (size_t)a.PX * (size_t)(j + a.OLy - 1) + a.slab * (size_t)a.Nr * tile;
    um_column<U, MODE>(a, q0, inner, alpha, alphaPrev, sCur, sPrev, zsm + threadIdx.x, a.nSm, er, eta);
  }
}

// SA: what one level of one column reads (its own k+1 values included)
struct SaOps { double zW, zE, zS, zN, oW, oE, oS, oN, zU, oU, aW0, aW1, aS0, aS1, aVp, aC; };

// SA for the interior columns: so = previous s (read), sn = new s (written, ring cells included).  The operands of
// level k+1 are requested before level k is evaluated (two register sets), so a column has two levels in flight.
__device__ __forceinline__ void sa_phase(const Cg3dArgs &a, double beta, const double *__restrict__ so,
                                         double *__restrict__ sn, double &al) {
  const size_t PX = a.PX, slab = a.slab;
  const int Nr = a.Nr;
  const double *__restrict__ z = a.z, *__restrict__ aW = a.aW, *__restrict__ aS = a.aS, *__restrict__ aV = a.aV,
               *__restrict__ aC = a.aC;
  double *__restrict__ qA = a.q;
  const int npx = (a.sNx + C3_TX - 1) / C3_TX, npy = (a.sNy + C3_TY - 1) / C3_TY;
  const size_t nPatch = (size_t)a.nTiles * npx * npy;
  const int lx = threadIdx.x % C3_TX, ly = threadIdx.x / C3_TX;
  for (size_t p = blockIdx.x; p < nPatch; p += gridDim.x) {
    Cell c;
    c.i = 1 + (int)(p % npx) * C3_TX + lx;
    c.j = 1 + (int)((p / npx) % npy) * C3_TY + ly;
    c.tile = (int)(p / ((size_t)npx * npy));
    if (c.i > a.sNx || c.j > a.sNy) continue;
    const bool eW = c.i == 1, eE = c.i == a.sNx, eS = c.j == 1, eN = c.j == a.sNy, edge = eW || eE || eS || eN;
    size_t q = (size_t)(c.i + a.OLx - 1) + PX * (size_t)(c.j + a.OLy - 1) + slab * (size_t)a.Nr * c.tile;
    double sm1 = 0., sc = z[q] + beta * so[q];       // s'(k-1), s'(k) of this column
    double aVk = 0.;                                 // aV3d(k)
    auto load = [&](size_t qq, bool up) {
      SaOps o;
      o.zW = z[qq - 1]; o.oW = so[qq - 1]; o.zE = z[qq + 1]; o.oE = so[qq + 1];
      o.zS = z[qq - PX]; o.oS = so[qq - PX]; o.zN = z[qq + PX]; o.oN = so[qq + PX];
      o.aW0 = aW[qq]; o.aW1 = aW[qq + 1]; o.aS0 = aS[qq]; o.aS1 = aS[qq + PX]; o.aC = aC[qq];
      if (up) { o.zU = z[qq + slab]; o.oU = so[qq + slab]; o.aVp = aV[qq + slab]; }
      else { o.zU = 0.; o.oU = 0.; o.aVp = 0.; }
      return o;
    };
    auto eval = [&](const SaOps o, int k, size_t qq) {
      const bool up = k < Nr;
      const double sW = o.zW + beta * o.oW, sE = o.zE + beta * o.oE;       // cg3d.F:313-317 at the neighbours
      const double sS = o.zS + beta * o.oS, sN = o.zN + beta * o.oN;
      const double sp1 = o.zU + beta * o.oU;                               // s'(k+1); unused at the bottom
      double v = o.aW0 * sW + o.aW1 * sE + o.aS0 * sS + o.aS1 * sN;        // :376-404
      if (k > 1) v = v + aVk * sm1;
      if (up) v = v + o.aVp * sp1;
      v = v + o.aC * sc;
      st_once(&qA[qq], v);
      st_once(&sn[qq], sc);
      al += sc * v;
      if (edge) {
        if (eW) sn[qq - 1] = sW;
        if (eE) sn[qq + 1] = sE;
        if (eS) sn[qq - PX] = sS;
        if (eN) sn[qq + PX] = sN;
        c.k = k;
        push3(a, c, qA, v);      // the ring columns of UM update their r from it
      }
      sm1 = sc;
      sc = sp1;
      aVk = o.aVp;
    };
    SaOps A = load(q, 1 < Nr), B = A;
    int k = 1;
    while (true) {
      if (k < Nr) B = load(q + slab, k + 1 < Nr);
      eval(A, k, q);
      if (++k > Nr) break;
      q += slab;
      if (k < Nr) A = load(q + slab, k + 1 < Nr);
      eval(B, k, q);
      if (++k > Nr) break;
      q += slab;
    }
  }
}

// MINB CTAs per SM, U levels in flight per column in UM (registers: about 14 per level)
template <int MINB, int U>
__global__ void __launch_bounds__(C3_THREADS, MINB) cg3d_fused_kernel(Cg3dArgs a) {
  cg::grid_group grid = cg::this_grid();
  __shared__ double sm[C3_WARPS + 1];
  extern __shared__ double zsm[];      // [nSm][C3_THREADS]
  const size_t tid = (size_t)blockIdx.x * C3_THREADS + threadIdx.x, nthr = (size_t)gridDim.x * C3_THREADS;
  const size_t nInt = (size_t)a.nTiles * a.Nr * a.sNy * a.sNx;

  const Cg3dStart st = cg3d_start(a, grid, sm, tid, nthr, nInt);
  double err_sq = st.err_sq;
  int actualIts = 0;
  if (!(err_sq < a.tolSq) && a.maxIters > 0) {
    double *sb[2] = {a.s, a.s1};
    int cur = 0;                       // sb[cur] holds the latest search direction
    double eta_qrNM1 = 1., er = 0., eta = 0., alphaPend = 0.;
    bool pending = false;              // x has not received alphaPend * sb[cur] yet
    um_phase<U, -1>(a, 0., 0., nullptr, nullptr, zsm, er, eta);      // z = M r, eta = <z,r> (cg3d.F:258-303)
    put_partial<false>(a, 0, eta, sm);
    grid.sync();
    unsigned long long t0 = now_ns(), nsSA = 0, nsUM = 0;
    for (int it3d = 1; it3d <= a.maxIters; it3d++) {
      const double eta_qrN = get_total<false>(a, 0, sm);
      const double cgBeta = eta_qrN / eta_qrNM1;      // :305-307
      eta_qrNM1 = eta_qrN;
      double al = 0.;
      sa_phase(a, cgBeta, sb[cur], sb[cur ^ 1], al);
      cur ^= 1;
      put_partial<false>(a, 1, al, sm);
      grid.sync();
      unsigned long long t1 = now_ns();
      nsSA += t1 - t0;
      const double alpha = eta_qrN / get_total<false>(a, 1, sm);      // :438-439
      er = 0.; eta = 0.;
      if (pending) {
        um_phase<U, 2>(a, alpha, alphaPend, sb[cur], sb[cur ^ 1], zsm, er, eta);
        pending = false;
      } else if (it3d < a.maxIters) {
        um_phase<U, 0>(a, alpha, 0., nullptr, nullptr, zsm, er, eta);
        pending = true;
        alphaPend = alpha;
      } else {
        um_phase<U, 1>(a, alpha, 0., sb[cur], nullptr, zsm, er, eta);
      }
      put_partial<false>(a, 2, er, sm);
      put_partial<false>(a, 0, eta, sm);
      grid.sync();
      t0 = now_ns();
      nsUM += t0 - t1;
      actualIts = it3d;
      err_sq = get_total<false>(a, 2, sm);
      if (err_sq < a.tolSq) break;
    }
    if (tid == 0) { a.out->nsSA = nsSA; a.out->nsUM = nsUM; }
    if (pending) {      // the solve ended on an odd iteration: its x update is still due
      const double *sC = sb[cur];
      for (size_t n = tid; n < nInt; n += nthr) {
        const Cell c = cell_of(a, n);
        a.x[c.idx] = a.x[c.idx] + alphaPend * sC[c.idx];
      }
    }
  }
  cg3d_finish(a, grid, st, err_sq, actualIts, tid, nthr, nInt);
}

struct Cg3dWs {
  double *r = nullptr, *q = nullptr, *s = nullptr, *z = nullptr, *s1 = nullptr, *partials = nullptr;
  Cg3dOut *out = nullptr;
  double sumRHS = 0., rhsMax = 0.;
};
static Cg3dWs g_cg3d;

void cg3d_free_workspace() {
  for (double *p : {g_cg3d.r, g_cg3d.q, g_cg3d.s, g_cg3d.z, g_cg3d.s1, g_cg3d.partials})
    if (p) cudaFree(p);
  if (g_cg3d.out) cudaFree(g_cg3d.out);
  g_cg3d = Cg3dWs{};
}

}  // namespace mg

using namespace mg;

extern "C" void cg3d_b200_(double *cg3d_b, double *cg3d_x, double *firstResidual, double *lastResidual, int *numIters,
                           const int *myIter, const int *myThid) {
  (void)myIter; (void)myThid;
  Ctx &c = ctx();
  c.lastError = 0;
  if (!c.ready) { fail(30, "mitgcm_b200_init_ not called"); return; }
  const Geom &g = c.g;
  if (g.nPx * g.nPy > 1) { fail(80, "cg3d_b200_: single rank only"); return; }
  if (c.p.I(MI_SELECT_RSTAR) != 0) { fail(80, "cg3d_b200_: select_rStar != 0 (surface term) is not on the B200 path"); return; }
  if (g.n3 >= ((size_t)1 << 31)) { fail(80, "cg3d_b200_: tile3d array too large"); return; }
  const int ids[8] = {MG_AW3D, MG_AS3D, MG_AV3D, MG_AC3D, MG_ZMC, MG_ZML, MG_ZMU, MG_MASKC};
  const double *op[8];
  for (int n = 0; n < 8; n++) {
    op[n] = field(ids[n], false);
    if (!op[n]) { fail(43, "cg3d_b200_: operator mirrors (MG_AW3D .. MG_ZMU, MG_MASKC) not set"); return; }
  }
  // MITGCM_B200_CG3D_UNFUSED=1: the four-sweep kernel; MITGCM_B200_CG3D_VARIANT=0..3: (CTAs per SM, levels in
  // flight) of the fused kernel (A/B measurements; every variant computes the same numbers)
  const char *ev = getenv("MITGCM_B200_CG3D_UNFUSED");
  const bool fused = !(ev && atoi(ev) != 0);
  ev = getenv("MITGCM_B200_CG3D_VARIANT");
  const int variant = ev ? atoi(ev) : C3_DEFAULT_VARIANT;
  if (variant < 0 || variant > 3) { fail(80, "cg3d_b200_: MITGCM_B200_CG3D_VARIANT must be 0..3"); return; }
  Cg3dWs &w = g_cg3d;
  auto alloc = [&](double **p, size_t n) { return *p || cudaMalloc(p, n * sizeof(double)) == cudaSuccess; };
  if (!alloc(&w.r, g.n3) || !alloc(&w.q, g.n3) || !alloc(&w.s, g.n3) || !alloc(&w.partials, 3 * C3_MAXB) ||
      (fused && (!alloc(&w.z, g.n3) || !alloc(&w.s1, g.n3))) ||
      (!w.out && cudaMalloc(&w.out, sizeof(Cg3dOut)) != cudaSuccess)) {
    cudaGetLastError();
    cg3d_free_workspace();      // never leave a half-allocated workspace behind
    fail(3, "cg3d_b200_: cudaMalloc failed");
    return;
  }
  Cg3dArgs a{};
  a.sNx = g.sNx; a.sNy = g.sNy; a.OLx = g.OLx; a.OLy = g.OLy; a.PX = g.PX; a.PY = g.PY; a.Nr = g.Nr; a.nTiles = g.nTiles;
  a.slab = g.slab;
  a.aW = op[0]; a.aS = op[1]; a.aV = op[2]; a.aC = op[3]; a.zMC = op[4]; a.zML = op[5]; a.zMU = op[6]; a.maskC = op[7];
  a.b = to_device(cg3d_b, g.n3, 60, true);
  a.x = to_device(cg3d_x, g.n3, 61, true);
  if (!a.b || !a.x) return;
  a.r = w.r; a.q = w.q; a.s = w.s; a.z = w.z; a.s1 = w.s1; a.pushTab = c.pushTab; a.partials = w.partials; a.out = w.out;
  a.cg3dNorm = c.p.D(MP_CG3DNORM); a.tolSq = c.p.D(MP_CG3DTOLERANCE_SQ);
  a.normaliseRHS = c.p.I(MI_CG3DNORMALISERHS); a.maxIters = *numIters;
  // zero-initialised work arrays incl. the ring (ini_cg3d.F:70-76, cg3d.F:223-227)
  for (double *p : {w.r, w.q, w.s, fused ? w.z : nullptr, fused ? w.s1 : nullptr})
    if (p && cudaMemsetAsync(p, 0, g.n3 * sizeof(double), c.stream) != cudaSuccess) { fail(4, "cg3d_b200_: memset failed"); return; }
  void *kern = (void *)cg3d_kernel;
  size_t dynSmem = 0;
  if (fused) {
    void *const variants[4] = {(void *)cg3d_fused_kernel<3, 3>, (void *)cg3d_fused_kernel<2, 5>,
                               (void *)cg3d_fused_kernel<2, 4>, (void *)cg3d_fused_kernel<1, 10>};
    const int minb[4] = {3, 2, 2, 1};
    kern = variants[variant];
    // z' in shared memory: as many levels as MINB co-resident CTAs can hold (at most C3_SMEM_LEVELS; MITGCM_B200_CG3D_SMEM_LEVELS overrides, 0 = none)
    int smPerSM = 0, smPerBlock = 0;
    cudaDeviceGetAttribute(&smPerSM, cudaDevAttrMaxSharedMemoryPerMultiprocessor, c.device);
    cudaDeviceGetAttribute(&smPerBlock, cudaDevAttrMaxSharedMemoryPerBlockOptin, c.device);
    long budget = std::min<long>((long)smPerBlock, (long)smPerSM / minb[variant] - 1024 - (long)sizeof(double) * (C3_WARPS + 1));
    const int fit = (int)std::max<long>(0, budget / (long)(C3_THREADS * sizeof(double)));
    int levels = fit;
    // measured at 1024^2 x 50 (2 CTAs/SM): UM 750 us/iteration with 0 levels, 727 with 16, 695 with 32; with all 50
    // (2 x 100 KB of the SM's 256 KB) the L1 is too small for SA's neighbour reuse: SA 757 -> 1378 us
    levels = std::min(levels, C3_SMEM_LEVELS);
    ev = getenv("MITGCM_B200_CG3D_SMEM_LEVELS");
    if (ev) levels = std::min(fit, std::max(0, atoi(ev)));
    a.nSm = std::min(levels, g.Nr);
    dynSmem = (size_t)a.nSm * C3_THREADS * sizeof(double);
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dynSmem) != cudaSuccess) {
      fail(5, std::string("cg3d_b200_: cudaFuncSetAttribute: ") + cudaGetErrorString(cudaGetLastError()));
      return;
    }
  }
  int nb = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, C3_THREADS, dynSmem) != cudaSuccess || nb < 1) { fail(5, "cg3d_b200_: occupancy query failed"); return; }
  const size_t nInt = (size_t)g.nTiles * g.Nr * g.sNy * g.sNx;
  // work items: cells (four-sweep form) / ring columns (fused form)
  const size_t nWork = fused ? (size_t)g.nTiles * (g.sNy + 2) * (g.sNx + 2) : nInt;
  int blocks = std::min({nb * c.numSMs, C3_MAXB, (int)std::min<size_t>((nWork + C3_THREADS - 1) / C3_THREADS, (size_t)C3_MAXB)});
  if (blocks < 1) blocks = 1;
  void *args[] = {&a};
  c.launches++;
  if (cudaLaunchCooperativeKernel(kern, dim3(blocks), dim3(C3_THREADS), args, dynSmem, c.stream) != cudaSuccess) {
    fail(5, std::string("cg3d_b200_: launch failed: ") + cudaGetErrorString(cudaGetLastError()));
    return;
  }
  Cg3dOut out;
  if (cudaMemcpyAsync(&out, w.out, sizeof(out), cudaMemcpyDeviceToHost, c.stream) != cudaSuccess) { fail(4, "cg3d_b200_: D2H failed"); return; }
  if (!from_device(cg3d_b, a.b, g.n3) || !from_device(cg3d_x, a.x, g.n3)) return;
  if (cudaStreamSynchronize(c.stream) != cudaSuccess) { fail(6, std::string("cg3d_b200_: ") + cudaGetErrorString(cudaGetLastError())); return; }
  *firstResidual = out.firstResidual;
  *lastResidual = out.lastResidual;
  *numIters = out.numIters;
  if (fused && out.numIters > 0 && getenv("MITGCM_B200_CG3D_TIMING"))
    fprintf(stderr, "cg3d_fused_kernel variant %d, %d levels of z' in shared memory: SA %.1f us, UM %.1f us per iteration (CTA 0, barriers included)\n",
            variant, a.nSm, out.nsSA * 1e-3 / out.numIters, out.nsUM * 1e-3 / out.numIters);
  w.sumRHS = out.sumRHS; w.rhsMax = out.rhsMax;
}

// the values cg3d.F:243-244 prints inside the solver (`cg3d: Sum(rhs),rhsMax`)
extern "C" void mitgcm_b200_cg3d_rhs_stats_(double *sumRHS, double *rhsMax) {
  *sumRHS = g_cg3d.sumRHS;
  *rhsMax = g_cg3d.rhsMax;
}
