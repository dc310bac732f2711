// cg3d.cu -- CG3D drop-in (model/src/cg3d.F:13-545; caller solve_for_pressure.F:427): the preconditioned
// conjugate-gradient solver of the 3-D (non-hydrostatic) pressure equation, SURVEY.md section 8(f) rank 4.
// A x = b with the 7-point operator {aW3d(i), aW3d(i+1), aS3d(j), aS3d(j+1), aV3d(k), aV3d(k+1), aC3d}; the
// preconditioner is the vertical tridiagonal solve per column (zMC, zML, zMU of INI_CG3D), applied on the ring
// 0..sN+1 so that no exchange of the search direction is needed.
// One persistent cooperative kernel per solve (as cg2d.cu): four grid barriers per iteration
//   M : q = M r down and up every column of the ring, eta = <q,r>            R{r,zMC,zML,zMU} W{q}
//   S : s = q + beta s on the ring                                           R{q,s} W{s}
//   A : q = A s, <s,q>                                                       R{s x7,aW,aS,aV,aC} W{q}
//   U : x += alpha s, r -= alpha q, <r,r>, edge values of r pushed into the neighbours' ring
//       (EXCH_S3D_RL( cg3d_r, Nr )) through the width-1 push table shared with CG2D          R{x,s,r,q} W{x,r}
// Dot products: lane -> warp shuffle -> CTA -> ordered sum over CTAs (deterministic run to run; the order differs
// from the reference's tile-ordered sum, as in CG2D).  Convergence is decided on the device.
// Algorithmic bytes per cell per iteration: M 6, S 3, A 6 (+ neighbours from L1/L2), U 6 words = 21 words = 168 B.
// The four sweeps are not fused yet (S could ride on A by recomputing s at the 7 points).
// Single rank, select_rStar = 0 (no surface term), as the oracle.  -fmad=false, reference operation order.
#include <cooperative_groups.h>
#include <algorithm>
#include "context.h"

namespace cg = cooperative_groups;

namespace mg {

constexpr int C3_THREADS = 256, C3_WARPS = C3_THREADS / 32, C3_MAXB = 1184;

struct Cg3dOut {
  double firstResidual, lastResidual, sumRHS, rhsMax;
  int numIters;
};

struct Cg3dArgs {
  int sNx, sNy, OLx, OLy, PX, PY, Nr, nTiles;
  size_t slab;
  const double *aW, *aS, *aV, *aC, *zMC, *zML, *zMU, *maskC;
  double *b, *x, *r, *q, *s;
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

  // cg3d.F:121-133: b *= cg3dNorm*maskC, rhsMax
  double acc = 0.;
  for (size_t n = tid; n < nInt; n += nthr) {
    const Cell c = cell_of(a, n);
    const double bv = a.b[c.idx] * a.cg3dNorm * a.maskC[c.idx];
    a.b[c.idx] = bv;
    acc = fmax(fabs(bv), acc);
  }
  put_partial<true>(a, 0, acc, sm);
  grid.sync();
  const double rhsMax = get_total<true>(a, 0, sm);
  double rhsNorm = 1.;
  if (a.normaliseRHS) {     // :135-153
    if (rhsMax != 0.) rhsNorm = 1. / rhsMax;
    for (size_t n = tid; n < nInt; n += nthr) {
      const Cell c = cell_of(a, n);
      a.b[c.idx] = a.b[c.idx] * rhsNorm;
      const double xv = a.x[c.idx] * rhsNorm;
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
  double err_sq = get_total<false>(a, 0, sm);
  const double sumRHS = get_total<false>(a, 1, sm);
  const double firstResidual = sqrt(err_sq);
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
  if (a.normaliseRHS) {     // :481-495
    grid.sync();
    for (size_t n = tid; n < nInt; n += nthr) {
      const Cell c = cell_of(a, n);
      a.x[c.idx] = a.x[c.idx] / rhsNorm;
    }
  }
  if (tid == 0) {
    a.out->firstResidual = firstResidual;
    a.out->lastResidual = sqrt(err_sq);
    a.out->sumRHS = sumRHS;
    a.out->rhsMax = rhsMax;
    a.out->numIters = actualIts;
  }
}

struct Cg3dWs {
  double *r = nullptr, *q = nullptr, *s = nullptr, *partials = nullptr;
  Cg3dOut *out = nullptr;
  double sumRHS = 0., rhsMax = 0.;
};
static Cg3dWs g_cg3d;

void cg3d_free_workspace() {
  for (double *p : {g_cg3d.r, g_cg3d.q, g_cg3d.s, g_cg3d.partials})
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
  Cg3dWs &w = g_cg3d;
  if (!w.r) {
    if (cudaMalloc(&w.r, g.n3 * sizeof(double)) != cudaSuccess || cudaMalloc(&w.q, g.n3 * sizeof(double)) != cudaSuccess ||
        cudaMalloc(&w.s, g.n3 * sizeof(double)) != cudaSuccess || cudaMalloc(&w.partials, 3 * C3_MAXB * sizeof(double)) != cudaSuccess ||
        cudaMalloc(&w.out, sizeof(Cg3dOut)) != cudaSuccess) { fail(3, "cg3d_b200_: cudaMalloc failed"); return; }
  }
  Cg3dArgs a{};
  a.sNx = g.sNx; a.sNy = g.sNy; a.OLx = g.OLx; a.OLy = g.OLy; a.PX = g.PX; a.PY = g.PY; a.Nr = g.Nr; a.nTiles = g.nTiles;
  a.slab = g.slab;
  a.aW = op[0]; a.aS = op[1]; a.aV = op[2]; a.aC = op[3]; a.zMC = op[4]; a.zML = op[5]; a.zMU = op[6]; a.maskC = op[7];
  a.b = to_device(cg3d_b, g.n3, 60, true);
  a.x = to_device(cg3d_x, g.n3, 61, true);
  if (!a.b || !a.x) return;
  a.r = w.r; a.q = w.q; a.s = w.s; a.pushTab = c.pushTab; a.partials = w.partials; a.out = w.out;
  a.cg3dNorm = c.p.D(MP_CG3DNORM); a.tolSq = c.p.D(MP_CG3DTOLERANCE_SQ);
  a.normaliseRHS = c.p.I(MI_CG3DNORMALISERHS); a.maxIters = *numIters;
  // zero-initialised work arrays incl. the ring (ini_cg3d.F:70-76, cg3d.F:223-227)
  for (double *p : {w.r, w.q, w.s})
    if (cudaMemsetAsync(p, 0, g.n3 * sizeof(double), c.stream) != cudaSuccess) { fail(4, "cg3d_b200_: memset failed"); return; }
  int nb = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, cg3d_kernel, C3_THREADS, 0) != cudaSuccess || nb < 1) { fail(5, "cg3d_b200_: occupancy query failed"); return; }
  const size_t nInt = (size_t)g.nTiles * g.Nr * g.sNy * g.sNx;
  int blocks = std::min({nb * c.numSMs, C3_MAXB, (int)((nInt + C3_THREADS - 1) / C3_THREADS)});
  if (blocks < 1) blocks = 1;
  void *args[] = {&a};
  c.launches++;
  if (cudaLaunchCooperativeKernel((void *)cg3d_kernel, dim3(blocks), dim3(C3_THREADS), args, 0, c.stream) != cudaSuccess) {
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
  w.sumRHS = out.sumRHS; w.rhsMax = out.rhsMax;
}

// the values cg3d.F:243-244 prints inside the solver (`cg3d: Sum(rhs),rhsMax`)
extern "C" void mitgcm_b200_cg3d_rhs_stats_(double *sumRHS, double *rhsMax) {
  *sumRHS = g_cg3d.sumRHS;
  *rhsMax = g_cg3d.rhsMax;
}
