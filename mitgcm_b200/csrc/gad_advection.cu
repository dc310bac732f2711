// gad_advection.cu -- GAD_ADVECTION drop-in (pkg/generic_advdiff/gad_advection.F:11-1097): the
// multi-dimensional ("direct space-time") advection of one tracer on one tile, all levels per call,
// with the reference argument list (callers temp_integrate.F:283, salt_integrate.F:275,
// ptracers_integrate.F).  The reference sweeps the halo'd slab level by level: an X pass and a Y pass
// that update a local copy of the tracer in sequence (each pass = one flux sweep + one update sweep),
// then a vertical pass k = Nr..1 with a two-slab flux ring.  Here every level is independent in the
// horizontal passes and the vertical flux at both interfaces of a cell is recomputed by its own
// thread, so the whole routine is three launches over (i, j, k):
//   md_pass_kernel<0>  T0 -> T1 (+ local volume)   fluxes from the tracer                 R{T,u,hFacW,hFacC}  W{T1,V1}
//   md_pass_kernel<1>  T1 -> T2 (+ local volume)   fluxes from the X-updated field        R{T1,V1,v,hFacS}    W{T2,V2}
//   md_vert_kernel     gTracer from T2             4..8-level stencil in k                R{T2,V2,w,T,maskC}  W{gTracer}
// Cubed sphere (pkg/exch2 topology set): the three facet-dependent passes of gad_advection.F:339-812 -- which
// direction a facet sweeps in which pass, interior-only / overlap-only update ranges, FILL_CS_CORNER_TR_RL as a
// read map of the pass input (and of its output corners), FILL_CS_CORNER_UV_RS on copies of the masks.
// Flux formulas are the leaves of gad.cuh shared with GAD_CALC_RHS
// (schemes 1, 20, 77, 30, 33, 7).  MI_GAD_MULTIDIM_COMPRESSIBLE selects the GAD_MULTIDIM_COMPRESSIBLE
// build variant (gad_advection.F:480-490, :1018-1032) at run time.  -fmad=false, reference operation
// order: bit-identical to the oracle, which is pinned to verification/advect_xy and verification/advect_cs.
#include "gad.cuh"

namespace mg {

// accessor the flux leaves read through: tracer values from the current pass's input, transports
// derived from the velocity of level k
// FILL_CS_CORNER_TR_RL (eesupp/src/fill_cs_corner_tr_rl.F:74-156) as an index map: the cell whose value (i,j) holds
// after the facet corners in `corners` (1 SW, 2 SE, 4 NE, 8 NW) were refilled for direction dir (1: x, 2: y).
__device__ __forceinline__ bool cs_corner_src(const TileGrid &g, int corners, int dir, int i, int j, int &ii, int &jj) {
  if (!dir || !corners) return false;
  const int sNx = g.sNx, sNy = g.sNy;
  if (i < 1 && j < 1 && (corners & 1)) {
    const int a = 1 - i, b = 1 - j;
    if (dir == 1) { ii = 1 - b; jj = a; } else { ii = b; jj = 1 - a; }
    return true;
  }
  if (i > sNx && j < 1 && (corners & 2)) {
    const int a = i - sNx, b = 1 - j;
    if (dir == 1) { ii = sNx + b; jj = a; } else { ii = sNx + 1 - b; jj = 1 - a; }
    return true;
  }
  if (i < 1 && j > sNy && (corners & 8)) {
    const int a = 1 - i, b = j - sNy;
    if (dir == 1) { ii = 1 - b; jj = sNy + 1 - a; } else { ii = b; jj = sNy + a; }
    return true;
  }
  if (i > sNx && j > sNy && (corners & 4)) {
    const int a = i - sNx, b = j - sNy;
    if (dir == 1) { ii = sNx + b; jj = sNy + 1 - a; } else { ii = sNx + 1 - b; jj = sNy + a; }
    return true;
  }
  return false;
}

struct MdAcc {
  TileGrid g;
  const double *T_, *u, *v, *w;   // per-tile (slab, Nr)
  int k;
  int corners, dirIn;             // cube: the pass input is read through the corner fill of direction dirIn (0: as stored)
  __device__ __forceinline__ double TA(int i, int j, int kk) const {
    int ii, jj;
    if (cs_corner_src(g, corners, dirIn, i, j, ii, jj)) return T_[g.s3(ii, jj, kk)];
    return T_[g.s3(i, j, kk)];
  }
  __device__ __forceinline__ double T(int i, int j, int kk) const { return TA(i, j, kk); }
  __device__ __forceinline__ double xA(int i, int j) const { return g.dyG[g.s(i, j)] * 1. * g.drF[k - 1] * g.hFacW[g.s3(i, j, k)]; }
  __device__ __forceinline__ double yA(int i, int j) const { return g.dxG[g.s(i, j)] * 1. * g.drF[k - 1] * g.hFacS[g.s3(i, j, k)]; }
  __device__ __forceinline__ double uFld(int i, int j) const { return u[g.s3(i, j, k)]; }
  __device__ __forceinline__ double vFld(int i, int j) const { return v[g.s3(i, j, k)]; }
  __device__ __forceinline__ double wFld(int i, int j) const { return w[g.s3(i, j, k)]; }
  __device__ __forceinline__ double uTrans(int i, int j) const { return u[g.s3(i, j, k)] * xA(i, j) * 1.; }
  __device__ __forceinline__ double vTrans(int i, int j) const { return v[g.s3(i, j, k)] * yA(i, j) * 1.; }
  __device__ __forceinline__ double rTrans(int i, int j) const {
    return k == 1 ? 0. : w[g.s3(i, j, k)] * g.rA[g.s(i, j)] * 1. * 1. * g.maskC[g.s3(i, j, k - 1)];
  }
  __device__ __forceinline__ double maskUp(int, int) const { return 0.; }
  __device__ __forceinline__ double rTransKp1(int, int) const { return 0.; }
  __device__ __forceinline__ double KappaR(int, int) const { return 0.; }
};

// One horizontal pass (DIR 0: X, gad_advection.F:376-560; DIR 1: Y, :597-790) over the halo'd slab of
// every level.  Tin / Vin: tracer and local volume before the pass (Vin == nullptr: the volume of the
// undisturbed cell, first pass); tracer0: the tracer at time n (the -T*div(U) correction of the default form).
struct MdUpd {
  int cube, overlapOnly, interiorOnly, N, S, E, W;   // gad_advection.F:339-362
  int dirOut;                                        // corner fill applied to the field after the fluxes (ipass = 1), 0: none
};

#ifndef MD_MINB
#define MD_MINB 16   // latency-bound point-wise kernels: full occupancy (32 registers) measured best: 32.0 -> 25.5 ms for the thermodynamics phase at 2048^2 x 50
#endif
template <int DIR>
__global__ void __launch_bounds__(128, MD_MINB) md_pass_kernel(TileGrid g, MdAcc a, GadPar p, const double *__restrict__ Vin,
                                                      const double *__restrict__ tracer0, double *__restrict__ Tout,
                                                      double *__restrict__ Vout, int compressible, const double *dTLev,
                                                      MdUpd q) {
  const int i = 1 - g.OLx + blockIdx.x * 32 + threadIdx.x;
  const int j = 1 - g.OLy + blockIdx.y * 4 + threadIdx.y;
  const int k = 1 + blockIdx.z;
  if (i > g.sNx + g.OLx || j > g.sNy + g.OLy) return;
  a.k = k; p.k = k; p.deltaT = dTLev[k - 1];
  const size_t s3 = g.s3(i, j, k);
  double T = a.TA(i, j, k);
  // the local volume is carried only by the GAD_MULTIDIM_COMPRESSIBLE form; the default form never reads it
  double V = 0.;
  if (compressible) V = Vin ? Vin[s3] : g.rA[g.s(i, j)] * 1. * 1. * g.drF[k - 1] * g.hFacC[s3] + (1. - g.maskC[s3]);
  // update range of this pass (gad_advection.F:471-600 for X, :692-812 for Y)
  bool upd;
  if (DIR == 0) {
    if (q.overlapOnly) {
      const int iLo = q.W ? 1 : 2 - g.OLx, iHi = q.E ? g.sNx : g.sNx + g.OLx - 1;
      upd = i >= iLo && i <= iHi && ((q.S && j <= 0) || (q.N && j >= g.sNy + 1));
    } else {
      const int jLo = (q.interiorOnly && q.S) ? 1 : 1 - g.OLy, jHi = (q.interiorOnly && q.N) ? g.sNy : g.sNy + g.OLy;
      upd = j >= jLo && j <= jHi && i >= 2 - g.OLx && i <= g.sNx + g.OLx - 1;
    }
  } else {
    if (q.overlapOnly) {
      const int jLo = q.S ? 1 : 2 - g.OLy, jHi = q.N ? g.sNy : g.sNy + g.OLy - 1;
      upd = j >= jLo && j <= jHi && ((q.W && i <= 0) || (q.E && i >= g.sNx + 1));
    } else {
      const int iLo = (q.interiorOnly && q.W) ? 1 : 1 - g.OLx, iHi = (q.interiorOnly && q.E) ? g.sNx : g.sNx + g.OLx;
      upd = i >= iLo && i <= iHi && j >= 2 - g.OLy && j <= g.sNy + g.OLy - 1;
    }
  }
  if (upd) {
    const int di = DIR == 0, dj = DIR == 1;
    const double af0 = gad_adv_h(g, a, p, DIR, i, j), af1 = gad_adv_h(g, a, p, DIR, i + di, j + dj);
    const double tr0 = DIR == 0 ? a.uTrans(i, j) : a.vTrans(i, j);
    const double tr1 = DIR == 0 ? a.uTrans(i + 1, j) : a.vTrans(i, j + 1);
    if (compressible) {
      const double tmpTrac = T * V - p.deltaT * (af1 - af0) * 1.;
      V = V - p.deltaT * (tr1 - tr0) * 1.;
      T = tmpTrac / V;
    } else {
      T = T - p.deltaT * 1. * g.recip_hFacC[s3] * g.recip_drF[k - 1] * g.recip_rA[g.s(i, j)] * 1. *
                  (af1 - af0 - tracer0[s3] * (tr1 - tr0)) * 1.;
    }
  } else if (q.dirOut) {     // FILL_CS_CORNER_TR_RL for the other direction after the fluxes (:456-459, :677-680)
    int ii, jj;
    if (cs_corner_src(g, a.corners, q.dirOut, i, j, ii, jj)) T = a.T_[g.s3(ii, jj, k)];
  }
  Tout[s3] = T;
  if (compressible) Vout[s3] = V;
}

// The same pass with every face flux and transport evaluated once: the thread of cell (i,j) evaluates the face it owns
// (west face for DIR 0, south face for DIR 1) and takes the far face from its neighbour -- by warp shuffle along x
// (warps overlap by one cell: 31 cells per 32 lanes, so no lane evaluates a second face), through shared memory
// along y (7 cell rows per 8 thread rows).  Same leaf, same arguments: bit-identical to md_pass_kernel.
#ifndef MDS_ROWS
#define MDS_ROWS 8      // 16 measured the same (thermodynamics phase 16.79 vs 16.76 ms at 2048^2 x 50)
#endif
template <int DIR>
__global__ void __launch_bounds__(32 * MDS_ROWS, 64 / MDS_ROWS) md_pass_share_kernel(TileGrid g, MdAcc a, GadPar p, const double *__restrict__ Vin,
                                                      const double *__restrict__ tracer0, double *__restrict__ Tout,
                                                      double *__restrict__ Vout, int compressible, const double *dTLev,
                                                      MdUpd q) {
  __shared__ double sAf[DIR == 1 ? MDS_ROWS : 1][32], sTr[DIR == 1 ? MDS_ROWS : 1][32];
  const int i = 1 - g.OLx + blockIdx.x * (DIR == 0 ? 31 : 32) + threadIdx.x;
  const int j = 1 - g.OLy + blockIdx.y * (DIR == 1 ? MDS_ROWS - 1 : MDS_ROWS) + threadIdx.y;
  const int k = 1 + blockIdx.z;
  const bool inb = i <= g.sNx + g.OLx && j <= g.sNy + g.OLy;
  // the last lane (DIR 0) / thread row (DIR 1) only supplies its face to the neighbour; its cell belongs to the next block
  const bool owner = inb && (DIR == 0 ? threadIdx.x < 31 : threadIdx.y < MDS_ROWS - 1);
  a.k = k; p.k = k; p.deltaT = dTLev[k - 1];
  // cells this pass updates (as md_pass_kernel), split into the part along the sweep and the part across it
  bool along, across;
  if (DIR == 0) {
    if (q.overlapOnly) {
      const int iLo = q.W ? 1 : 2 - g.OLx, iHi = q.E ? g.sNx : g.sNx + g.OLx - 1;
      along = i >= iLo && i <= iHi; across = (q.S && j <= 0) || (q.N && j >= g.sNy + 1);
    } else {
      const int jLo = (q.interiorOnly && q.S) ? 1 : 1 - g.OLy, jHi = (q.interiorOnly && q.N) ? g.sNy : g.sNy + g.OLy;
      across = j >= jLo && j <= jHi; along = i >= 2 - g.OLx && i <= g.sNx + g.OLx - 1;
    }
  } else {
    if (q.overlapOnly) {
      const int jLo = q.S ? 1 : 2 - g.OLy, jHi = q.N ? g.sNy : g.sNy + g.OLy - 1;
      along = j >= jLo && j <= jHi; across = (q.W && i <= 0) || (q.E && i >= g.sNx + 1);
    } else {
      const int iLo = (q.interiorOnly && q.W) ? 1 : 1 - g.OLx, iHi = (q.interiorOnly && q.E) ? g.sNx : g.sNx + g.OLx;
      across = i >= iLo && i <= iHi; along = j >= 2 - g.OLy && j <= g.sNy + g.OLy - 1;
    }
  }
  const bool upd = along && across;
  // own face: needed by this cell or by the previous one along the sweep; faces 2-OL .. sN+OL are the ones any update reads
  const int c = DIR == 0 ? i : j, cHi = DIR == 0 ? g.sNx + g.OLx : g.sNy + g.OLy, OL = DIR == 0 ? g.OLx : g.OLy;
  double af0 = 0., tr0 = 0.;
  if (inb && across && c >= 2 - OL && c <= cHi) {
    af0 = gad_adv_h(g, a, p, DIR, i, j);
    tr0 = DIR == 0 ? a.uTrans(i, j) : a.vTrans(i, j);
  }
  double af1, tr1;
  if (DIR == 0) {
    af1 = __shfl_down_sync(0xffffffffu, af0, 1);
    tr1 = __shfl_down_sync(0xffffffffu, tr0, 1);
  } else {
    sAf[threadIdx.y][threadIdx.x] = af0; sTr[threadIdx.y][threadIdx.x] = tr0;
    __syncthreads();
    const int ty = min((int)threadIdx.y + 1, MDS_ROWS - 1);
    af1 = sAf[ty][threadIdx.x]; tr1 = sTr[ty][threadIdx.x];
  }
  if (!owner) return;
  const size_t s3 = g.s3(i, j, k);
  double T = a.TA(i, j, k);
  double V = 0.;
  if (compressible) V = Vin ? Vin[s3] : g.rA[g.s(i, j)] * 1. * 1. * g.drF[k - 1] * g.hFacC[s3] + (1. - g.maskC[s3]);
  if (upd) {
    if (compressible) {
      const double tmpTrac = T * V - p.deltaT * (af1 - af0) * 1.;
      V = V - p.deltaT * (tr1 - tr0) * 1.;
      T = tmpTrac / V;
    } else {
      T = T - p.deltaT * 1. * g.recip_hFacC[s3] * g.recip_drF[k - 1] * g.recip_rA[g.s(i, j)] * 1. *
                  (af1 - af0 - tracer0[s3] * (tr1 - tr0)) * 1.;
    }
  } else if (q.dirOut) {
    int ii, jj;
    if (cs_corner_src(g, a.corners, q.dirOut, i, j, ii, jj)) T = a.T_[g.s3(ii, jj, k)];
  }
  Tout[s3] = T;
  if (compressible) Vout[s3] = V;
}

// FILL_CS_CORNER_UV_RS (eesupp/src/fill_cs_corner_uv_rs.F:46-108, withSigns = .FALSE.) on copies of maskW, maskS of
// one tile (gad_advection.F:329-334); every source is a non-corner cell, so the fill is a pure gather.
__global__ void md_mask_kernel(TileGrid g, int corners, double *__restrict__ mW, double *__restrict__ mS) {
  const int i = 1 - g.OLx + blockIdx.x * 32 + threadIdx.x;
  const int j = 1 - g.OLy + blockIdx.y * 4 + threadIdx.y;
  const int k = 1 + blockIdx.z;
  if (i > g.sNx + g.OLx || j > g.sNy + g.OLy) return;
  const size_t s3 = g.s3(i, j, k);
  double w = g.maskW[s3], s = g.maskS[s3];
  const int sNx = g.sNx, sNy = g.sNy;
  auto W = [&](int ii, int jj) { return g.maskW[g.s3(ii, jj, k)]; };
  auto S = [&](int ii, int jj) { return g.maskS[g.s3(ii, jj, k)]; };
  if (i < 1 && j < 1 && (corners & 1)) {
    const int a = 1 - i, b = 1 - j;
    w = S(1 - b, 1 + a); s = W(1 + b, 1 - a);
  } else if (i > sNx && j < 1 && (corners & 2)) {
    const int a = i - sNx, b = 1 - j;
    if (a >= 2) w = S(sNx + b, a);
    s = W(sNx + 1 - b, 1 - a);
  } else if (i < 1 && j > sNy && (corners & 8)) {
    const int a = 1 - i, b = j - sNy;
    w = S(1 - b, sNy + 1 - a);
    if (b >= 2) s = W(b, sNy + a);
  } else if (i > sNx && j > sNy && (corners & 4)) {
    const int a = i - sNx, b = j - sNy;
    if (a >= 2) w = S(sNx + b, sNy + 2 - a);
    if (b >= 2) s = W(sNx + 2 - b, sNy + a);
  }
  mW[s3] = w; mS[s3] = s;
}

// X+Y passes only (implicitAdvection): gTracer = (T2 - tracer)/deltaT (gad_advection.F:815-823)
__global__ void md_implicit_kernel(TileGrid g, const double *__restrict__ T2, const double *__restrict__ tracer0,
                                   double *__restrict__ gTracer, const double *dTLev) {
  const int i = 1 - g.OLx + blockIdx.x * 32 + threadIdx.x;
  const int j = 1 - g.OLy + blockIdx.y * 4 + threadIdx.y;
  const int k = 1 + blockIdx.z;
  if (i > g.sNx + g.OLx || j > g.sNy + g.OLy) return;
  const size_t s3 = g.s3(i, j, k);
  gTracer[s3] = (T2[s3] - tracer0[s3]) / dTLev[k - 1];
}

// Vertical pass (gad_advection.F:886-1050): fVerT at the upper (k) and lower (k+1) interface of the
// cell from the Y-updated field, then the tendency.
__global__ void __launch_bounds__(128, MD_MINB) md_vert_kernel(TileGrid g, MdAcc a, GadPar p, const double *__restrict__ V2,
                                                      const double *__restrict__ tracer0, double *__restrict__ gTracer,
                                                      int compressible, const double *dTLev) {
  const int i = 1 - g.OLx + blockIdx.x * 32 + threadIdx.x;
  const int j = 1 - g.OLy + blockIdx.y * 4 + threadIdx.y;
  const int k = 1 + blockIdx.z;
  if (i > g.sNx + g.OLx || j > g.sNy + g.OLy) return;
  const size_t s3 = g.s3(i, j, k);
  p.deltaT = dTLev[k - 1];
  // upper interface (kUp): zero at k = 1
  a.k = k; p.k = k;
  double fUp = 0., rTrans = 0.;
  if (k >= 2) { rTrans = a.rTrans(i, j); fUp = gad_adv_r(g, a, p, i, j); }
  // lower interface (kDown) = upper interface of level k+1, left by the previous level of the k = Nr..1 march;
  // its CFL number uses deltaTLev(k+1) (the leaf is called for level k+1)
  double fDn = 0., rTransKp = 0.;
  if (k < g.Nr) {
    MdAcc b = a;
    GadPar q = p;
    b.k = k + 1; q.k = k + 1; q.deltaT = dTLev[k];
    rTransKp = b.rTrans(i, j);
    fDn = gad_adv_r(g, b, q, i, j);
  }
  const double T2 = a.T_[s3];
  if (compressible) {
    const double tmpTrac = T2 * V2[s3] - p.deltaT * (fDn - fUp) * p.rkSign * 1.;
    const double vol = V2[s3] - p.deltaT * (rTransKp - rTrans) * p.rkSign * 1.;
    gTracer[s3] = (tmpTrac - tracer0[s3] * vol) * g.recip_rA[g.s(i, j)] * 1. * g.recip_drF[k - 1] * g.recip_hFacC[s3] * 1. /
                  p.deltaT;
  } else {
    const double lt = T2 - p.deltaT * 1. * g.recip_hFacC[s3] * g.recip_drF[k - 1] * g.recip_rA[g.s(i, j)] * 1. *
                               (fDn - fUp - tracer0[s3] * (rTransKp - rTrans)) * p.rkSign * 1.;
    gTracer[s3] = (lt - tracer0[s3]) / p.deltaT;
  }
}

// The same vertical pass as a column march: one thread per (i,j), k = 1..Nr.  The flux and transport of the lower
// interface are carried to the next level as its upper interface (same leaf call, same deltaTLev: bit-identical to
// md_vert_kernel), so every interface is evaluated once and the k-stencil re-reads hit L1 / L2 instead of DRAM
// (ncu at 1024^2 x 50: the plane-parallel kernel moved 130 B/cell against 48 B/cell algorithmic).
__global__ void __launch_bounds__(128, MD_MINB) md_vert_col_kernel(TileGrid g, MdAcc a, GadPar p, const double *__restrict__ V2,
                                                          const double *__restrict__ tracer0, double *__restrict__ gTracer,
                                                          int compressible, const double *dTLev) {
  const int i = 1 - g.OLx + blockIdx.x * 32 + threadIdx.x;
  const int j = 1 - g.OLy + blockIdx.y * 4 + threadIdx.y;
  if (i > g.sNx + g.OLx || j > g.sNy + g.OLy) return;
  const double recip_rA = g.recip_rA[g.s(i, j)];
  double fUp = 0., rTrans = 0.;
  for (int k = 1; k <= g.Nr; k++) {
    const size_t s3 = g.s3(i, j, k);
    const double dT = dTLev[k - 1];
    double fDn = 0., rTransKp = 0.;
    if (k < g.Nr) {
      a.k = k + 1; p.k = k + 1; p.deltaT = dTLev[k];
      rTransKp = a.rTrans(i, j);
      fDn = gad_adv_r(g, a, p, i, j);
    }
    const double T2 = a.T_[s3];
    if (compressible) {
      const double tmpTrac = T2 * V2[s3] - dT * (fDn - fUp) * p.rkSign * 1.;
      const double vol = V2[s3] - dT * (rTransKp - rTrans) * p.rkSign * 1.;
      gTracer[s3] = (tmpTrac - tracer0[s3] * vol) * recip_rA * 1. * g.recip_drF[k - 1] * g.recip_hFacC[s3] * 1. / dT;
    } else {
      const double lt = T2 - dT * 1. * g.recip_hFacC[s3] * g.recip_drF[k - 1] * recip_rA * 1. *
                                 (fDn - fUp - tracer0[s3] * (rTransKp - rTrans)) * p.rkSign * 1.;
      gTracer[s3] = (lt - tracer0[s3]) / dT;
    }
    fUp = fDn; rTrans = rTransKp;
  }
}

static bool md_scheme(int s) {
  return s == ADV_UPWIND_1RST || s == ADV_DST2 || s == ADV_FLUX_LIMIT || s == ADV_DST3 || s == ADV_DST3_FLUX_LIMIT ||
         s == ADV_OS7MP;
}

}  // namespace mg

using namespace mg;

namespace mg {

// GAD_ADVECTION for one tile on device pointers: u, v, w, tr, gT are the (slab, Nr) arrays of the tile, dT the
// deltaTLev(Nr) array.  Shared by the per-call entry point and the resident step (multi-dimensional advection of theta).
bool gad_advection_tile(TileGrid tg, size_t tile, int advScheme, int vertScheme, int implicitAdvection, const double *u,
                        const double *v, const double *w, const double *tr, double *gT, const double *dT) {
  Ctx &c = ctx();
  const Geom &g = c.g;
  const int compressible = c.p.I(MI_GAD_MULTIDIM_COMPRESSIBLE);
  if (!md_scheme(advScheme) || (!implicitAdvection && !md_scheme(vertScheme)))
    return fail(42, "gad_advection: advection scheme incompatible with multi-dim advection");
  if (implicitAdvection && compressible) return fail(42, "gad_advection: implicitAdvection with GAD_MULTIDIM_COMPRESSIBLE");
  if (advScheme == ADV_OS7MP && (g.OLx < 4 || g.OLy < 4)) return fail(42, "gad_advection: OS7MP needs OLx, OLy >= 4");
  const size_t n3 = g.slab * g.Nr;
  // cubed sphere: facet number and facet edges of this tile (gad_advection.F:249-258)
  const bool cube = exch2_active();
  int nCFace = 0, edges = 0;
  if (cube) {
    if ((int)c.csFace.size() != g.nTiles || (int)c.csEdges.size() != g.nTiles) return fail(44, "gad_advection: mitgcm_b200_set_cs_tiles_ not called");
    if (g.OLx != g.OLy) return fail(44, "gad_advection: the cubed-sphere form needs OLx = OLy");
    nCFace = c.csFace[tile]; edges = c.csEdges[tile];
  }
  const int eN = edges & 1, eS = (edges >> 1) & 1, eE = (edges >> 2) & 1, eW = (edges >> 3) & 1;
  const int corners = (eW && eS ? 1 : 0) | (eE && eS ? 2 : 0) | (eE && eN ? 4 : 0) | (eW && eN ? 8 : 0);
  double *Tb[2] = {to_device(nullptr, n3, 45, false), to_device(nullptr, n3, 47, false)};
  double *Vb[2] = {to_device(nullptr, n3, 46, false), to_device(nullptr, n3, 48, false)};
  if (!Tb[0] || !Vb[0] || !Tb[1] || !Vb[1]) return false;
  dim3 blk(32, 4), grd((g.PX + 31) / 32, (g.PY + 3) / 4, g.Nr);
  if (cube && corners) {     // maskLocW / maskLocS with their facet corners filled
    double *mW = to_device(nullptr, n3, 54, false), *mS = to_device(nullptr, n3, 55, false);
    if (!mW || !mS) return false;
    c.launches++;
    md_mask_kernel<<<grd, blk, 0, c.stream>>>(tg, corners, mW, mS);
    tg.maskW = mW; tg.maskS = mS;
  }
  GadPar p{};
  p.advScheme = advScheme;
  // GAD_DST2U1_ADV_R is handed advectionScheme, not vertAdvecScheme (gad_advection.F:976)
  p.vertAdvScheme = (vertScheme == ADV_UPWIND_1RST || vertScheme == ADV_DST2) ? advScheme : vertScheme;
  p.calcAdvection = 1; p.rkSign = c.p.D(MP_RKSIGN);
  MdAcc a{tg, tr, u, v, w, 1, corners, 0};
  const double *Tin = tr, *Vin = nullptr;
  int cur = 0;
  const int npass = cube ? 3 : 2;
  const bool noShare = getenv("MITGCM_B200_MD_NOSHARE") != nullptr;      // per-cell evaluation of both faces (md_pass_kernel)
  for (int ipass = 1; ipass <= npass; ipass++) {
    MdUpd q{};
    bool fluxX, fluxY;
    if (cube) {
      q.cube = 1; q.N = eN; q.S = eS; q.E = eE; q.W = eW;
      if (ipass == 1) {
        q.overlapOnly = nCFace % 3 == 0; q.interiorOnly = nCFace % 3 != 0;
        fluxX = nCFace == 6 || nCFace == 1 || nCFace == 2; fluxY = nCFace == 3 || nCFace == 4 || nCFace == 5;
      } else if (ipass == 2) {
        q.overlapOnly = nCFace % 3 == 2; q.interiorOnly = nCFace % 3 == 1;
        fluxX = nCFace == 2 || nCFace == 3 || nCFace == 4; fluxY = nCFace == 5 || nCFace == 6 || nCFace == 1;
      } else {
        q.interiorOnly = 1;
        fluxX = nCFace == 5 || nCFace == 6; fluxY = nCFace == 2 || nCFace == 3;
      }
    } else { fluxX = ipass % 2 == 1; fluxY = !fluxX; }
    for (int dir = 0; dir < 2; dir++) {
      if (!(dir == 0 ? fluxX : fluxY)) continue;
      // overlap-only passes touch nothing unless the tile lies on the relevant facet edges (:386, :607)
      const bool edgeOk = dir == 0 ? (eN || eS) : (eE || eW);
      if (q.overlapOnly && !edgeOk) continue;
      a.T_ = Tin;
      a.dirIn = q.overlapOnly ? (dir == 0 ? 1 : 2) : 0;
      q.dirOut = (q.overlapOnly && ipass == 1) ? (dir == 0 ? 2 : 1) : 0;
      c.launches++;
      if (noShare) {
        if (dir == 0) md_pass_kernel<0><<<grd, blk, 0, c.stream>>>(tg, a, p, Vin, tr, Tb[cur], Vb[cur], compressible, dT, q);
        else md_pass_kernel<1><<<grd, blk, 0, c.stream>>>(tg, a, p, Vin, tr, Tb[cur], Vb[cur], compressible, dT, q);
      } else {
        const dim3 sblk(32, MDS_ROWS);
        if (dir == 0) md_pass_share_kernel<0><<<dim3((g.PX + 30) / 31, (g.PY + MDS_ROWS - 1) / MDS_ROWS, g.Nr), sblk, 0, c.stream>>>(
            tg, a, p, Vin, tr, Tb[cur], Vb[cur], compressible, dT, q);
        else md_pass_share_kernel<1><<<dim3((g.PX + 31) / 32, (g.PY + MDS_ROWS - 2) / (MDS_ROWS - 1), g.Nr), sblk, 0, c.stream>>>(
            tg, a, p, Vin, tr, Tb[cur], Vb[cur], compressible, dT, q);
      }
      Tin = Tb[cur]; Vin = Vb[cur];
      cur ^= 1;
    }
  }
  a.T_ = Tin; a.dirIn = 0;
  c.launches++;
  if (implicitAdvection) md_implicit_kernel<<<grd, blk, 0, c.stream>>>(tg, Tin, tr, gT, dT);
  else if (getenv("MITGCM_B200_MD_VERT_PLANE")) md_vert_kernel<<<grd, blk, 0, c.stream>>>(tg, a, p, Vin, tr, gT, compressible, dT);
  else md_vert_col_kernel<<<dim3(grd.x, grd.y, 1), blk, 0, c.stream>>>(tg, a, p, Vin, tr, gT, compressible, dT);
  if (cudaGetLastError() != cudaSuccess) return fail(5, "gad_advection: launch failed");
  return true;
}

}  // namespace mg

extern "C" void gad_advection_b200_(const int *implicitAdvection, const int *advectionSchArg, const int *vertAdvecSchArg,
                                    const int *trIdentity, const double *deltaTLev, const double *uFld,
                                    const double *vFld, const double *wFld, const double *tracer, double *gTracer,
                                    const int *bi, const int *bj, const double *myTime, const int *myIter,
                                    const int *myThid) {
  (void)trIdentity; (void)myTime; (void)myIter; (void)myThid;
  Ctx &c = ctx();
  c.lastError = 0;
  if (!c.ready) { fail(30, "mitgcm_b200_init_ not called"); return; }
  const Geom &g = c.g;
  TileGrid tg;
  if (!make_tile_grid(*bi, *bj, tg)) { if (!c.lastError) fail(43, "grid mirrors not set"); return; }
  const size_t n3 = g.slab * g.Nr;
  const size_t tile = (size_t)(*bi - 1) + (size_t)g.nSx * (size_t)(*bj - 1);
  const double *u = to_device(uFld, n3, 40, true), *v = to_device(vFld, n3, 41, true), *w = to_device(wFld, n3, 42, true);
  // tracer is the full (.., Nr, nSx, nSy) array of the caller; only this tile is needed
  const double *tr = is_device_ptr(tracer) ? tracer + n3 * tile : to_device(tracer + n3 * tile, n3, 43, true);
  double *gT = is_device_ptr(gTracer) ? gTracer : to_device(gTracer, n3, 44, false);
  double *dT = to_device(deltaTLev, (size_t)g.Nr, 49, true);
  if (!u || !v || !w || !tr || !gT || !dT) return;
  if (!gad_advection_tile(tg, tile, *advectionSchArg, *vertAdvecSchArg, *implicitAdvection, u, v, w, tr, gT, dT)) return;
  if (!from_device(gTracer, gT, n3)) return;
  if (cudaStreamSynchronize(c.stream) != cudaSuccess) fail(6, "gad_advection_b200_: stream error");
}
