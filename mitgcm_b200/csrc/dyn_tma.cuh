// dyn_tma.cuh -- MOM_FLUXFORM + TIMESTEP + ADAMS_BASHFORTH2 of the resident step with every per-level operand
// brought into shared memory by the bulk-tensor copy engine (TMA) and an mbarrier ring.
//
// Same decomposition and the same arithmetic, expression for expression, as dyn_pipe_kernel (step_fast.cuh): a CTA
// of 32 x 8 columns marches k = 1..Nr over its 34 x 10 patch.  What changed is how operands arrive:
//   * one elected thread issues 16 cp.async.bulk.tensor copies per level -- the patch boxes of u, v, hFacW, hFacS,
//     hFacC (level k), w, maskC (level k+1), phiHyd (level k) and the 34 x 8 own-column boxes of kappaRU/V(k+1),
//     recip_hFacW/S, guNm1, gvNm1 (level k), maskW/S (level k+1) -- into the ring slot the CTA is NOT computing on;
//     the copies complete on the slot's mbarrier (complete_tx), which every thread polls once per level;
//   * so the level loop holds no per-element cp.async, no 64-bit address arithmetic for 14 arrays and only two
//     global loads per cell (u, v of the own column at k+1, L2 hits), against 14 LDGSTS + 13 LDG before;
//   * the upper vertical viscous flux is the lower one of the level above (same expression, carried), and the
//     identically-zero biharmonic terms of the harmonic-only fast path are not evaluated.
// Results are identical to dyn_pipe_kernel / dyn_kernel<0> (tests/test_step_gpu.py compares the three).
// Fast-path conditions: those of dyn_fast_ok, momAdvection and momViscosity on, PX and OLx even (16-byte aligned box
// rows), Nr < 160.
#pragma once
#include "step_fast.cuh"
#include "tma.cuh"
#include <type_traits>

namespace mg {

constexpr int DT_PATCH_D = 352;            // 34 * 10 = 340 doubles, padded to a multiple of 128 bytes
// The copy engine wants the first element of every box row 16-byte aligned: the inner coordinate must be even.  The
// patch origin x0 = 32*blockIdx.x + OLx - 2 is even for even OLx, the own-column boxes therefore start at x0 too
// (34 columns x the 8 own rows) and a thread reads its column at tx + 1.
constexpr int DT_OWN_W = FT_W, DT_OWN_D = DT_OWN_W * FT_Y;      // 34 x 8 = 272 doubles = 2176 bytes (17 x 128)
constexpr uint32_t DT_PATCH_BYTES = FT_N * 8, DT_OWN_BYTES = DT_OWN_D * 8;
enum { DP_U = 0, DP_V, DP_HW, DP_HS, DP_HC, DP_W, DP_MC, DP_PHI, DP_NPATCH };
enum { DO_KU = 0, DO_KV, DO_RHW, DO_RHS, DO_GUO, DO_GVO, DO_MW, DO_MS, DO_NOWN };

struct DynTmaStage {
  double patch[DP_NPATCH][DT_PATCH_D];
  double own[DO_NOWN][DT_OWN_D];
};
template <int NST>
struct DynTmaSmemN {
  static constexpr bool CG = false;
  DynTmaStage st[NST];
  double uT[FT_N], vT[FT_N], wA[FT_N], hZ[FT_N], mCk[FT_N], dyG[FT_N], dxG[FT_N], rA[FT_N];
  VertSmem vs;
  uint64_t full[NST];
};
typedef DynTmaSmemN<2> DynTmaSmem;
struct DynTmaMaps {
  CUtensorMap u, v, hW, hS, hC, w, mC, phi;      // box 34 x 10 x 1
  CUtensorMap kU, kV, rhW, rhS, guO, gvO, mW, mS;   // box 34 x 8 x 1
};

// level k (1-based) of the CTA's patch into stage s; x0, y0 = array coordinates of the patch origin
__device__ __forceinline__ void dyn_tma_issue(const DynTmaMaps &m, DynTmaStage &s, uint64_t *bar, int x0, int y0, int k, int Nr,
                                              bool hasPhi) {
  const bool below = k + 1 <= Nr;
  mbar_expect_tx(bar, 5 * DT_PATCH_BYTES + 6 * DT_OWN_BYTES + (below ? 2 * DT_PATCH_BYTES + 2 * DT_OWN_BYTES : 0) +
                          (hasPhi ? DT_PATCH_BYTES : 0));
  tma_load3(s.patch[DP_U], &m.u, x0, y0, k - 1, bar);
  tma_load3(s.patch[DP_V], &m.v, x0, y0, k - 1, bar);
  tma_load3(s.patch[DP_HW], &m.hW, x0, y0, k - 1, bar);
  tma_load3(s.patch[DP_HS], &m.hS, x0, y0, k - 1, bar);
  tma_load3(s.patch[DP_HC], &m.hC, x0, y0, k - 1, bar);
  if (hasPhi) tma_load3(s.patch[DP_PHI], &m.phi, x0, y0, k - 1, bar);
  tma_load3(s.own[DO_KU], &m.kU, x0, y0 + 1, k, bar);          // kappaRU(k+1): (Nr+1)-level array
  tma_load3(s.own[DO_KV], &m.kV, x0, y0 + 1, k, bar);
  tma_load3(s.own[DO_RHW], &m.rhW, x0, y0 + 1, k - 1, bar);
  tma_load3(s.own[DO_RHS], &m.rhS, x0, y0 + 1, k - 1, bar);
  tma_load3(s.own[DO_GUO], &m.guO, x0, y0 + 1, k - 1, bar);
  tma_load3(s.own[DO_GVO], &m.gvO, x0, y0 + 1, k - 1, bar);
  if (below) {
    tma_load3(s.patch[DP_W], &m.w, x0, y0, k, bar);
    tma_load3(s.patch[DP_MC], &m.mC, x0, y0, k, bar);
    tma_load3(s.own[DO_MW], &m.mW, x0, y0 + 1, k, bar);
    tma_load3(s.own[DO_MS], &m.mS, x0, y0 + 1, k, bar);
  }
}

#ifndef DYNT_MINB
#define DYNT_MINB 2
#endif
__global__ void __launch_bounds__(FT_X *FT_Y, DYNT_MINB)
    dyn_tma_kernel(const __grid_constant__ DynTmaMaps maps, TileGrid g, MomState st, MomPar p, const double *__restrict__ sfU,
                   const double *__restrict__ sfV, double *__restrict__ gU, double *__restrict__ gV, double *__restrict__ guNm1,
                   double *__restrict__ gvNm1, double deltaTMom, double abFac, int momForcing, int dissInAB, int hasPhi) {
  extern __shared__ __align__(1024) unsigned char dyn_tma_smem_raw[];      // TMA destinations need 128-byte alignment
  DynTmaSmem &sm = *reinterpret_cast<DynTmaSmem *>(dyn_tma_smem_raw);
  const int tx = threadIdx.x, ty = threadIdx.y, t = ty * FT_X + tx;
  const int i0 = blockIdx.x * FT_X, j0 = blockIdx.y * FT_Y;     // output range 0..sN+1 (dynamics.F:191-192)
  const int i = i0 + tx, j = j0 + ty;
  const bool active = i <= g.sNx + 1 && j <= g.sNy + 1;
  const int c = (ty + 1) * FT_W + (tx + 1);                     // my cell in the patch
  const int co = ty * DT_OWN_W + (tx + 1);                      // ... and in the own-column boxes
  const int x0 = i0 - 1 + g.OLx - 1, y0 = j0 - 1 + g.OLy - 1;   // patch origin in array coordinates (>= 0: OLx, OLy >= 2)
  if (t == 0) {
    mbar_init(&sm.full[0], 1);
    mbar_init(&sm.full[1], 1);
    mbar_fence_init();
  }
  stage_vert(sm.vs, g, t, FT_X * FT_Y);
  // k-invariant metrics of the patch cells (clamped into the slab; cells past it feed inactive threads only)
  bool hzok[2];
#pragma unroll
  for (int r = 0; r < 2; r++) {
    const int e = t + r * FT_X * FT_Y;
    hzok[r] = false;
    if (e < FT_N) {
      const int li = e % FT_W, lj = e / FT_W;
      hzok[r] = li >= 1 && lj >= 1;
      const int gi = min(i0 - 1 + li, g.sNx + g.OLx), gj = min(j0 - 1 + lj, g.sNy + g.OLy);
      const size_t q = g.s(gi, gj);
      sm.dyG[e] = g.dyG[q]; sm.dxG[e] = g.dxG[q]; sm.rA[e] = g.rA[q];
    }
  }
  __syncthreads();      // barriers initialised, k-invariants staged
  if (t == 0) {
    // pseudo-level 0 in slot 0: w and maskC of level 1 (surface interface, mom_fluxform.F:384-417) and maskW/S(1)
    mbar_expect_tx(&sm.full[0], 2 * DT_PATCH_BYTES + 2 * DT_OWN_BYTES);
    tma_load3(sm.st[0].patch[DP_W], &maps.w, x0, y0, 0, &sm.full[0]);
    tma_load3(sm.st[0].patch[DP_MC], &maps.mC, x0, y0, 0, &sm.full[0]);
    tma_load3(sm.st[0].own[DO_MW], &maps.mW, x0, y0 + 1, 0, &sm.full[0]);
    tma_load3(sm.st[0].own[DO_MS], &maps.mS, x0, y0 + 1, 0, &sm.full[0]);
    dyn_tma_issue(maps, sm.st[1], &sm.full[1], x0, y0, 1, g.Nr, hasPhi != 0);
  }
  // k-invariant metrics of my column
  const size_t s = active ? g.s(i, j) : g.s(0, 0);
  const int PX = g.PX;
  const double r_rAw = g.recip_rAw[s], r_rAs = g.recip_rAs[s], rAw = g.rAw[s], rAs = g.rAs[s];
  const double dyF00 = g.dyF[s], dyFm0 = g.dyF[s - 1], rdxF00 = g.recip_dxF[s], rdxFm0 = g.recip_dxF[s - 1];
  const double dxV00 = g.dxV[s], dxV01 = g.dxV[s + PX], rdyU00 = g.recip_dyU[s], rdyU01 = g.recip_dyU[s + PX];
  const double dyU00 = g.dyU[s], dyU10 = g.dyU[s + 1], rdxV00 = g.recip_dxV[s], rdxV10 = g.recip_dxV[s + 1];
  const double dxF00 = g.dxF[s], dxF0m = g.dxF[s - PX], rdyF00 = g.recip_dyF[s], rdyF0m = g.recip_dyF[s - PX];
  const double cfU = g.cosFacU[j + g.OLy - 1 < g.PY ? j + g.OLy - 1 : 0], cfV = g.cosFacV[j + g.OLy - 1 < g.PY ? j + g.OLy - 1 : 0];
  const double fC00 = g.fCori[s], fCm0 = g.fCori[s - 1], fC0m = g.fCori[s - PX];
  const double uDudxFac = p.afFacMom, AhFac = p.vfFacMom, ArFac = p.implicitViscosity ? 0. : p.vfFacMom;
  // CALC_GRAD_PHI_HYD (calc_grad_phi_hyd.F:150-165) is defined on i >= iMin+1, j >= jMin+1
  const double gpx = (hasPhi && i >= 1) ? g.recip_dxC[s] : 0., gpy = (hasPhi && j >= 1) ? g.recip_dyC[s] : 0.;
  const size_t slab = g.slab;
  const double *__restrict__ uCol = st.u + s, *__restrict__ vCol = st.v + s;     // own column, advanced level by level
  // wait for pseudo-level 0 and level 1
  mbar_wait(&sm.full[0], 0);
  mbar_wait(&sm.full[1], 0);
  unsigned ph = 3;                  // bit s = parity the NEXT wait on slot s expects
  double uKm1 = 0., vKm1 = 0.;
  double uK = sm.st[1].patch[DP_U][c], vK = sm.st[1].patch[DP_V][c];
  double mWk = sm.st[0].own[DO_MW][co], mSk = sm.st[0].own[DO_MS][co];
  double kapUk = st.kapU[s], kapVk = st.kapV[s];
  (void)kapUk; (void)kapVk;
#pragma unroll
  for (int r = 0; r < 2; r++) {
    const int e = t + r * FT_X * FT_Y;
    if (e < FT_N) sm.wA[e] = sm.st[0].patch[DP_W][e] * sm.rA[e];
  }
  __syncthreads();
  double ukm = 0., vkm = 0.;
  if (p.momAdvection && !p.rigidLid) {
    ukm = (0.5 * (sm.wA[c - 1] + sm.wA[c])) * uK;
    vkm = (0.5 * (sm.wA[c - FT_W] + sm.wA[c])) * vK;
  }
  double fVrUp = 0., gVrUp = 0.;      // vertical viscous fluxes at the upper interface = the lower ones of the level above
  __syncthreads();
  for (int k = 1; k <= g.Nr; k++) {
    const int rb = k & 1;
    const DynTmaStage &S = sm.st[rb];
    const double *__restrict__ mCk1 = sm.st[rb ^ 1].patch[DP_MC];      // maskC(k), fetched with level k-1
    const double drFk = sm.vs.drF[k - 1], rdrF = sm.vs.rdrF[k - 1];
    const bool below = k + 1 <= g.Nr;
    double uKp1 = 0., vKp1 = 0.;
    if (below) { uKp1 = uCol[slab]; vKp1 = vCol[slab]; }
    uCol += slab; vCol += slab;
    if (k > 1) {      // level k was requested one level ago
      mbar_wait(&sm.full[rb], (ph >> rb) & 1);
      ph ^= 1u << rb;
      __syncthreads();      // every thread is done with the derived arrays of level k-1
    }
    // ---- derive what the fluxes share ----
#pragma unroll
    for (int r = 0; r < 2; r++) {
      const int e = t + r * FT_X * FT_Y;
      if (e < FT_N) {
        const double hW = S.patch[DP_HW][e], hS = S.patch[DP_HS][e];
        sm.uT[e] = S.patch[DP_U][e] * (sm.dyG[e] * drFk * hW);
        sm.vT[e] = S.patch[DP_V][e] * (sm.dxG[e] * drFk * hS);
        if (below) sm.wA[e] = S.patch[DP_W][e] * sm.rA[e];
        sm.mCk[e] = mCk1[e];
        if (hzok[r]) {                               // MOM_CALC_HFACZ at the south-west corner
          // MIN of finite open-water fractions: a compare + select (fmin's NaN handling costs 7 instructions a piece)
          const double hWs = S.patch[DP_HW][e - FT_W], hSw = S.patch[DP_HS][e - 1];
          double h = hW < hWs ? hW : hWs;
          h = hS < h ? hS : h;
          h = hSw < h ? hSw : h;
          sm.hZ[e] = h;
        }
      }
    }
    __syncthreads();      // derived arrays visible; slot rb^1 (level k-1) is no longer read by anybody
    if (t == 0 && below) dyn_tma_issue(maps, sm.st[rb ^ 1], &sm.full[rb ^ 1], x0, y0, k + 1, g.Nr, hasPhi != 0);
    const double mWkp1 = below ? S.own[DO_MW][co] : 0., mSkp1 = below ? S.own[DO_MS][co] : 0.;
    const double kapUkp1 = S.own[DO_KU][co], kapVkp1 = S.own[DO_KV][co];
    if (active) {
#define PU(di, dj) S.patch[DP_U][c + (dj)*FT_W + (di)]
#define PV(di, dj) S.patch[DP_V][c + (dj)*FT_W + (di)]
#define PHC(di, dj) S.patch[DP_HC][c + (dj)*FT_W + (di)]
#define PMC1(di, dj) S.patch[DP_MC][c + (dj)*FT_W + (di)]
#define DUT(di, dj) sm.uT[c + (dj)*FT_W + (di)]
#define DVT(di, dj) sm.vT[c + (dj)*FT_W + (di)]
#define DHZ(di, dj) sm.hZ[c + (dj)*FT_W + (di)]
#define DWA(di, dj) sm.wA[c + (dj)*FT_W + (di)]
#define DMC(di, dj) sm.mCk[c + (dj)*FT_W + (di)]
      const double rhW = S.own[DO_RHW][co], rhS = S.own[DO_RHS][co];
      const double guOld = S.own[DO_GUO][co], gvOld = S.own[DO_GVO][co];
      double dpx = 0., dpy = 0.;
      if (hasPhi) {
        const double phc = S.patch[DP_PHI][c];
        dpx = gpx * 1. * (phc - S.patch[DP_PHI][c - 1]) * 1.;
        dpy = gpy * 1. * (phc - S.patch[DP_PHI][c - FT_W]) * 1.;
      }
      const double u00 = PU(0, 0), v00 = PV(0, 0);
      const double uE = PU(1, 0), uW = PU(-1, 0), uN = PU(0, 1), uS = PU(0, -1);
      const double vE = PV(1, 0), vW = PV(-1, 0), vN = PV(0, 1), vS = PV(0, -1);
      // vertical advective fluxes at interface k+1 (MOM_U_ADV_WU / MOM_V_ADV_WV)
      double ukp = 0., vkp = 0.;
      if (p.momAdvection && below) {
        const double wA00 = DWA(0, 0), wAm0 = DWA(-1, 0), wA0m = DWA(0, -1);
        const double rTU = 0.5 * (wAm0 + wA00), rTV = 0.5 * (wA0m + wA00);
        ukp = rTU * 0.5 * (uKp1 + uK);
        vkp = rTV * 0.5 * (vKp1 + vK);
        if (!p.rigidLid) {
          const double d00 = PMC1(0, 0) - DMC(0, 0);
          ukp = ukp + 0.25 * (wA00 * d00 + wAm0 * (PMC1(-1, 0) - DMC(-1, 0))) * uKp1;
          vkp = vkp + 0.25 * (wA00 * d00 + wA0m * (PMC1(0, -1) - DMC(0, -1))) * vKp1;
        }
      }
      double gu = 0., gv = 0., guD = 0., gvD = 0.;
      if (p.momAdvection) {
        const double uT00 = DUT(0, 0), uT10 = DUT(1, 0), vT00 = DVT(0, 0), vT01 = DVT(0, 1);
        const double fzU1 = 0.25 * (uT00 + uT10) * (u00 + uE);
        const double fzU0 = 0.25 * (DUT(-1, 0) + uT00) * (uW + u00);
        const double fmU1 = 0.25 * (vT01 + DVT(-1, 1)) * (uN + u00);
        const double fmU0 = 0.25 * (vT00 + DVT(-1, 0)) * (u00 + uS);
        gu = -rhW * rdrF * r_rAw * ((fzU1 - fzU0) * uDudxFac + (fmU1 - fmU0) * uDudxFac + (ukp - ukm) * p.rkSign * uDudxFac);
        const double fzV1 = 0.25 * (uT10 + DUT(1, -1)) * (vE + v00);
        const double fzV0 = 0.25 * (uT00 + DUT(0, -1)) * (v00 + vW);
        const double fmV1 = 0.25 * (vT00 + vT01) * (v00 + vN);
        const double fmV0 = 0.25 * (DVT(0, -1) + vT00) * (vS + v00);
        gv = -rhS * rdrF * r_rAs * ((fzV1 - fzV0) * uDudxFac + (fmV1 - fmV0) * uDudxFac + (vkp - vkm) * p.rkSign * uDudxFac);
      }
      if (p.momViscosity) {
        const double hZ00 = DHZ(0, 0), hZ01 = DHZ(0, 1), hZ10 = DHZ(1, 0), hC00 = PHC(0, 0);
        // U: MOM_U_XVISCFLUX, MOM_U_YVISCFLUX, MOM_U_RVISCFLUX (harmonic only: the del2 terms of the generic form are 0)
        const double xv1 = dyF00 * drFk * hC00 * (-p.viscAhD * (uE - u00) * cfU) * rdxF00;
        const double xv0 = dyFm0 * drFk * PHC(-1, 0) * (-p.viscAhD * (u00 - uW) * cfU) * rdxFm0;
        const double yv1 = dxV01 * drFk * hZ01 * (-p.viscAhZ * (uN - u00)) * rdyU01;
        const double yv0 = dxV00 * drFk * hZ00 * (-p.viscAhZ * (u00 - uS)) * rdyU00;
        double fVrDw = 0.;
        if (!p.implicitViscosity && below) fVrDw = -kapUkp1 * rAw * (uKp1 - uK) * p.rkSign * sm.vs.rdrC[k] * mWkp1 * mWk;
        guD = -rhW * rdrF * r_rAw * ((xv1 - xv0) * AhFac + (yv1 - yv0) * AhFac + (fVrDw - fVrUp) * p.rkSign * ArFac);
        fVrUp = fVrDw;
        // V: MOM_V_XVISCFLUX, MOM_V_YVISCFLUX, MOM_V_RVISCFLUX
        const double xw1 = dyU10 * drFk * hZ10 * (-p.viscAhZ * (vE - v00) * cfV) * rdxV10;
        const double xw0 = dyU00 * drFk * hZ00 * (-p.viscAhZ * (v00 - vW) * cfV) * rdxV00;
        const double yw1 = dxF00 * drFk * hC00 * (-p.viscAhD * (vN - v00)) * rdyF00;
        const double yw0 = dxF0m * drFk * PHC(0, -1) * (-p.viscAhD * (v00 - vS)) * rdyF0m;
        double gVrDw = 0.;
        if (!p.implicitViscosity && below) gVrDw = -kapVkp1 * rAs * (vKp1 - vK) * p.rkSign * sm.vs.rdrC[k] * mSkp1 * mSk;
        gvD = -rhS * rdrF * r_rAs * ((xw1 - xw0) * AhFac + (yw1 - yw0) * AhFac + (gVrDw - gVrUp) * p.rkSign * ArFac);
        gVrUp = gVrDw;
        if (p.no_slip_sides) {   // MOM_U_SIDEDRAG / MOM_V_SIDEDRAG
          const double hWc = S.patch[DP_HW][c], hSc = S.patch[DP_HS][c];
          const double tu = p.viscAhZ * u00;
          guD = guD + (-rhW * rdrF * r_rAw * ((hWc - hZ00) * dxV00 * rdyU00 * tu + (hWc - hZ01) * dxV01 * rdyU01 * tu) * drFk * p.sideDragFactor);
          const double tv = p.viscAhZ * v00 * cfV;
          gvD = gvD + (-rhS * rdrF * r_rAs * ((hSc - hZ00) * dyU00 * rdxV00 * tv + (hSc - hZ10) * dyU10 * rdxV10 * tv) * drFk * p.sideDragFactor);
        }
        if (p.bottomDragTerms) {  // MOM_{U,V}_BOTDRAG_COEFF with selectBotDragQuadr = -1
          const double viscFac = p.no_slip_bottom ? 2. : 0.;
          const double recDrC = (k == g.Nr) ? rdrF : sm.vs.rdrC[k];
          double cu = p.bottomDragLinear * 1., cv = p.bottomDragLinear * 1.;
          if (p.no_slip_bottom && p.bottomVisc_pCell) { cu = cu + kapUkp1 * recDrC * viscFac * rhW; cv = cv + kapVkp1 * recDrC * viscFac * rhS; }
          else if (p.no_slip_bottom) { cu = cu + kapUkp1 * recDrC * viscFac; cv = cv + kapVkp1 * recDrC * viscFac; }
          if (k == g.Nr) { cu = cu * mWk; cv = cv * mSk; }
          else { cu = cu * mWk * (1. - mWkp1); cv = cv * mSk * (1. - mSkp1); }
          guD = guD - cu * u00 * rhW * rdrF;
          gvD = gvD - cv * v00 * rhS * rdrF;
        }
      }
      if (!p.useCDscheme) {      // MOM_U_CORIOLIS / MOM_V_CORIOLIS
        double uCf, vCf;
        if (p.selectCoriScheme >= 2) {
          uCf = 0.5 * (fC00 * 0.5 * (v00 + vN) + fCm0 * 0.5 * (vW + PV(-1, 1)));
          vCf = -0.5 * (fC00 * 0.5 * (u00 + uE) + fC0m * 0.5 * (uS + PU(1, -1)));
        } else {
          uCf = 0.5 * (fC00 + fCm0) * 0.25 * (v00 + vN + vW + PV(-1, 1));
          vCf = -0.5 * (fC00 + fC0m) * 0.25 * (u00 + uE + uS + PU(1, -1));
        }
        gu = gu + p.cfFacMom * uCf;
        gv = gv + p.cfFacMom * vCf;
      }
      gu = gu * mWk; guD = guD * mWk; gv = gv * mSk; gvD = gvD * mSk;   // mom_fluxform.F:1044-1051
      // ---- TIMESTEP (timestep.F:95-385), as in dyn_kernel ----
      gu = gu - 1. * dpx; gv = gv - 1. * dpy;      // timestep.F:120-121, phFac = pfFacMom = 1
      if (p.momViscosity && dissInAB) { gu = gu + guD; gv = gv + gvD; }
      if (momForcing) {
        double ge = 0., he = 0.;
        if (k == 1) {
          if (i >= 1 && i <= g.sNx + 1) ge = 0. + sfU[s] * sm.vs.rdrF[0] * rhW;
          if (j >= 1 && j <= g.sNy + 1) he = 0. + sfV[s] * sm.vs.rdrF[0] * rhS;
        }
        gu = gu + ge; gv = gv + he;
      }
      const size_t s3 = s + slab * (size_t)(k - 1);
      double ab = abFac * (gu - guOld);
      guNm1[s3] = gu;
      gu = gu + ab;
      ab = abFac * (gv - gvOld);
      gvNm1[s3] = gv;
      gv = gv + ab;
      if (p.momViscosity && !dissInAB) { gu = gu + guD; gv = gv + gvD; }
      gU[s3] = uK + deltaTMom * (gu + 0.) * mWk;
      gV[s3] = vK + deltaTMom * (gv + 0.) * mSk;
      ukm = ukp; vkm = vkp;
#undef PU
#undef PV
#undef PHC
#undef PMC1
#undef DUT
#undef DVT
#undef DHZ
#undef DWA
#undef DMC
    }
    uKm1 = uK; uK = uKp1; vKm1 = vK; vK = vKp1; mWk = mWkp1; mSk = mSkp1;
    (void)uKm1; (void)vKm1;
  }
}

// ---- column-geometry variant of the staging (colgeom.cu): the eight geometry boxes per level (hFacW/S/C, maskC and the
// own-column recip_hFacW/S, maskW/S) are not fetched at all; the CTA keeps (kLow, hLow) of its patch for the three
// point types in shared memory and rebuilds the level values by compare + select.  Half the TMA traffic and half the
// ring: 20 KB per stage, so three stages fit twice per SM.
enum { CP_U = 0, CP_V, CP_W, CP_PHI, CP_NPATCH };
enum { CO_KU = 0, CO_KV, CO_GUO, CO_GVO, CO_NOWN };
struct DynTmaStageCG {
  double patch[CP_NPATCH][DT_PATCH_D];
  double own[CO_NOWN][DT_OWN_D];
};
template <int NST>
struct DynTmaSmemCG {
  static constexpr bool CG = true;
  DynTmaStageCG st[NST];
  double uT[FT_N], vT[FT_N], wA[FT_N], hZ[FT_N], hC[FT_N], dyG[FT_N], dxG[FT_N], rA[FT_N];
  double hLW[FT_N], hLS[FT_N], hLC[FT_N];
  int kLW[FT_N], kLS[FT_N], kLC[FT_N];
  // CALC_PHI_HYD fused (PHIF): the PHI slot of the ring carries theta, every patch cell integrates its own column
  double phi[FT_N], dM[FT_NRMAX], dP[FT_NRMAX], tRf[FT_NRMAX];
  VertSmem vs;
  uint64_t full[NST];
};
// linear equation of state + the vertical grid CALC_PHI_HYD integrates over (phys.cuh: phihyd_kernel, same expressions)
struct PhiFuse {
  double rhoNil, dRho, tAlpha, gravity, recip_rhoConst;
  const double *tRef, *rF, *rC;
};
__device__ __forceinline__ void dyn_tma_issue(const DynTmaMaps &m, DynTmaStageCG &s, uint64_t *bar, int x0, int y0, int k, int Nr,
                                              bool hasPhi) {
  const bool below = k + 1 <= Nr;
  mbar_expect_tx(bar, 2 * DT_PATCH_BYTES + 4 * DT_OWN_BYTES + (below ? DT_PATCH_BYTES : 0) + (hasPhi ? DT_PATCH_BYTES : 0));
  tma_load3(s.patch[CP_U], &m.u, x0, y0, k - 1, bar);
  tma_load3(s.patch[CP_V], &m.v, x0, y0, k - 1, bar);
  if (hasPhi) tma_load3(s.patch[CP_PHI], &m.phi, x0, y0, k - 1, bar);
  tma_load3(s.own[CO_KU], &m.kU, x0, y0 + 1, k, bar);
  tma_load3(s.own[CO_KV], &m.kV, x0, y0 + 1, k, bar);
  tma_load3(s.own[CO_GUO], &m.guO, x0, y0 + 1, k - 1, bar);
  tma_load3(s.own[CO_GVO], &m.gvO, x0, y0 + 1, k - 1, bar);
  if (below) tma_load3(s.patch[CP_W], &m.w, x0, y0, k, bar);
}
// pseudo-level 0 in slot 0: what level 1 needs from "the level above" (w of level 1; the 3-D-array form also maskC, maskW/S)
__device__ __forceinline__ void dyn_tma_issue0(const DynTmaMaps &m, DynTmaStageCG &s, uint64_t *bar, int x0, int y0) {
  mbar_expect_tx(bar, DT_PATCH_BYTES);
  tma_load3(s.patch[CP_W], &m.w, x0, y0, 0, bar);
}
__device__ __forceinline__ void dyn_tma_issue0(const DynTmaMaps &m, DynTmaStage &s, uint64_t *bar, int x0, int y0) {
  mbar_expect_tx(bar, 2 * DT_PATCH_BYTES + 2 * DT_OWN_BYTES);
  tma_load3(s.patch[DP_W], &m.w, x0, y0, 0, bar);
  tma_load3(s.patch[DP_MC], &m.mC, x0, y0, 0, bar);
  tma_load3(s.own[DO_MW], &m.mW, x0, y0 + 1, 0, bar);
  tma_load3(s.own[DO_MS], &m.mS, x0, y0 + 1, 0, bar);
}

// ---- role-split variant: 512 threads, warps 0-7 compute the U tendency, warps 8-15 the V tendency ----------------
// dyn_tma_kernel is issue-bound at 4 warps per scheduler (128 registers x 256 threads x 2 CTAs fill the register
// file; ncu: 25 % warps active, 36 % issue-active, FP64 pipe 32 %).  Splitting the two momentum components over two
// thread groups of the SAME CTA halves what a thread has to keep alive (k-invariant metrics, carried vertical
// fluxes, own-column operands): twice the warps on the same staged shared memory, same TMA traffic, same
// expressions (bit-identical output; tests/test_step_gpu.py).  The component is a template parameter of the level
// loop (ROLE 0: U, 1: V), so array choices and neighbour offsets are immediates, not selects: the first version with
// a run-time role spent 22 % of its instructions on IMAD / ISETP / FSEL (profiles/r02_dyn_tma_uv_kernel_*.txt).
// NST = depth of the TMA ring (2: two CTAs per SM; 3-5: one CTA per SM with NST - 1 levels in flight).
__device__ __forceinline__ void dyn_uv_bar() { asm volatile("bar.sync 1, %0;" ::"n"(2 * FT_X * FT_Y) : "memory"); }

// hFacC of a patch cell, and maskC(k+1) - maskC(k), from the staged arrays or from the column geometry
template <bool CG, class SM, class ST>
__device__ __forceinline__ double dyn_hc(const SM &sm, const ST &S, int e) {
  if constexpr (CG) return sm.hC[e];
  else return S.patch[DP_HC][e];
}
template <bool CG, class SM, class ST>
__device__ __forceinline__ double dyn_dmask(const SM &sm, const ST &S, int e, int k) {
  if constexpr (CG) return sm.kLC[e] == k ? -1. : 0.;      // 1 - 1, 0 - 1, 0 - 0
  else return S.patch[DP_MC][e] - sm.mCk[e];
}

template <int NST, int ROLE, bool PHIF, class SM>
__device__ __forceinline__ void dyn_uv_levels(const PhiFuse &pf, SM &sm, const DynTmaMaps &maps, const TileGrid &g, const MomState &st,
                                              const MomPar &p, const double *__restrict__ sf, double *__restrict__ gOut,
                                              double *__restrict__ gNm1, double deltaTMom, double abFac, int momForcing,
                                              int dissInAB, int hasPhi, int t, int i, int j, bool active, bool hzok, int c,
                                              int co, int x0, int y0) {
  constexpr bool CG = SM::CG;
  constexpr int dP = ROLE ? FT_W : 1;              // patch offset towards "my" neighbour: west for U, south for V
  constexpr int P_U = CG ? (int)CP_U : (int)DP_U, P_V = CG ? (int)CP_V : (int)DP_V, P_W = CG ? (int)CP_W : (int)DP_W;
  constexpr int P_PHI = CG ? (int)CP_PHI : (int)DP_PHI;
  constexpr int P_F = ROLE ? P_V : P_U, O_M = ROLE ? DO_MS : DO_MW;
  constexpr int O_K = CG ? (ROLE ? (int)CO_KV : (int)CO_KU) : (ROLE ? (int)DO_KV : (int)DO_KU);
  constexpr int O_RH = ROLE ? DO_RHS : DO_RHW, O_G = CG ? (ROLE ? (int)CO_GVO : (int)CO_GUO) : (ROLE ? (int)DO_GVO : (int)DO_GUO);
  const size_t s = active ? g.s(i, j) : g.s(0, 0);
  const int PX = g.PX;
  const int dn = ROLE ? PX : 1;
  // U: dyF00, dyFm0, rdxF00, rdxFm0, dxV00, dxV01, rdyU00, rdyU01;  V: dxF00, dxF0m, rdyF00, rdyF0m, dyU00, dyU10, rdxV00, rdxV10
  const double r_rA = ROLE ? g.recip_rAs[s] : g.recip_rAw[s], rAx = ROLE ? g.rAs[s] : g.rAw[s];
  const double mA0 = ROLE ? g.dxF[s] : g.dyF[s], mA1 = ROLE ? g.dxF[s - PX] : g.dyF[s - 1];
  const double mB0 = ROLE ? g.recip_dyF[s] : g.recip_dxF[s], mB1 = ROLE ? g.recip_dyF[s - PX] : g.recip_dxF[s - 1];
  const double mC0 = ROLE ? g.dyU[s] : g.dxV[s], mC1 = ROLE ? g.dyU[s + 1] : g.dxV[s + PX];
  const double mD0 = ROLE ? g.recip_dxV[s] : g.recip_dyU[s], mD1 = ROLE ? g.recip_dxV[s + 1] : g.recip_dyU[s + PX];
  const int jc = j + g.OLy - 1 < g.PY ? j + g.OLy - 1 : 0;
  const double cf = ROLE ? g.cosFacV[jc] : g.cosFacU[jc];
  const double fC00 = g.fCori[s], fCn = g.fCori[s - dn];
  const double uDudxFac = p.afFacMom, AhFac = p.vfFacMom, ArFac = p.implicitViscosity ? 0. : p.vfFacMom;
  const double gp = ROLE ? ((hasPhi && j >= 1) ? g.recip_dyC[s] : 0.) : ((hasPhi && i >= 1) ? g.recip_dxC[s] : 0.);
  const size_t slab = g.slab;
  const double *__restrict__ fCol = (ROLE ? st.v : st.u) + s;      // own column of my component, advanced level by level
  const bool forcRange = ROLE ? (j >= 1 && j <= g.sNy + 1) : (i >= 1 && i <= g.sNx + 1);
  mbar_wait(&sm.full[0], 0);
  mbar_wait(&sm.full[1], 0);
  unsigned ph = 3;
  double fK = sm.st[1].patch[P_F][c];
  // column geometry of my own point (CG): deepest wet level, hFac and recip_hFac there
  int kLo = 0;
  double hLo = 0., rhLo = 0.;
  double mk;
  if constexpr (CG) {
    kLo = ROLE ? g.kLowS[s] : g.kLowW[s];
    hLo = ROLE ? g.hLowS[s] : g.hLowW[s];
    rhLo = ROLE ? g.rhLowS[s] : g.rhLowW[s];
    mk = cg_mask(1, kLo);
  } else
    mk = sm.st[0].own[O_M][co];
  if (t < FT_N) sm.wA[t] = sm.st[0].patch[P_W][t] * sm.rA[t];
  dyn_uv_bar();
  double fkm = 0.;
  if (!p.rigidLid) fkm = (0.5 * (sm.wA[c - dP] + sm.wA[c])) * fK;
  double fVrUp = 0.;
  double phiF = 0.;                 // PHIF: potential at the upper face of level k of MY patch cell (thread t < FT_N)
  dyn_uv_bar();
  int rb = 0;                       // ring slot of level k; rp = slot of level k-1 (freed after the derive phase)
  for (int k = 1; k <= g.Nr; k++) {
    const int rp = rb;
    rb = rb + 1 == NST ? 0 : rb + 1;
    const auto &S = sm.st[rb];
    const double drFk = sm.vs.drF[k - 1], rdrF = sm.vs.rdrF[k - 1];
    const bool below = k + 1 <= g.Nr;
    double fKp1 = 0.;
    if (below) fKp1 = fCol[slab];
    fCol += slab;
    if (k > 1) {
      mbar_wait(&sm.full[rb], (ph >> rb) & 1);
      ph ^= 1u << rb;
      dyn_uv_bar();      // every thread is done with the derived arrays of level k-1
    }
    if (t < FT_N) {
      const int e = t;
      double hW, hS;
      if constexpr (CG) {
        hW = cg_hfac(k, sm.kLW[e], sm.hLW[e]); hS = cg_hfac(k, sm.kLS[e], sm.hLS[e]);
        sm.hC[e] = cg_hfac(k, sm.kLC[e], sm.hLC[e]);
      } else {
        hW = S.patch[DP_HW][e]; hS = S.patch[DP_HS][e];
        sm.mCk[e] = sm.st[rp].patch[DP_MC][e];      // maskC(k), fetched with level k-1
      }
      sm.uT[e] = S.patch[P_U][e] * (sm.dyG[e] * drFk * hW);
      sm.vT[e] = S.patch[P_V][e] * (sm.dxG[e] * drFk * hS);
      if (below) sm.wA[e] = S.patch[P_W][e] * sm.rA[e];
      if constexpr (PHIF) {      // CALC_PHI_HYD (calc_phi_hyd.F:240-262), the expressions of phihyd_kernel
        const double a_ = pf.rhoNil * (0. - pf.tAlpha * (S.patch[P_PHI][e] - sm.tRf[k - 1])) + pf.dRho;
        const double phiC = phiF + sm.dM[k - 1] * pf.gravity * a_ * pf.recip_rhoConst;
        phiF = phiC + sm.dP[k - 1] * pf.gravity * a_ * pf.recip_rhoConst;
        sm.phi[e] = phiC;
      }
      if (hzok) {
        double hWs, hSw;
        if constexpr (CG) { hWs = cg_hfac(k, sm.kLW[e - FT_W], sm.hLW[e - FT_W]); hSw = cg_hfac(k, sm.kLS[e - 1], sm.hLS[e - 1]); }
        else { hWs = S.patch[DP_HW][e - FT_W]; hSw = S.patch[DP_HS][e - 1]; }
        double h = hW < hWs ? hW : hWs;
        h = hS < h ? hS : h;
        h = hSw < h ? hSw : h;
        sm.hZ[e] = h;
      }
    }
    dyn_uv_bar();      // derived arrays visible; slot rp (level k-1) is no longer read by anybody
    if (t == 0 && k + NST - 1 <= g.Nr) dyn_tma_issue(maps, sm.st[rp], &sm.full[rp], x0, y0, k + NST - 1, g.Nr, hasPhi != 0);
    double mkp1;
    if constexpr (CG) mkp1 = below ? cg_mask(k + 1, kLo) : 0.;
    else mkp1 = below ? S.own[O_M][co] : 0.;
    const double kapkp1 = S.own[O_K][co];
    if (active) {
#define PU(di, dj) S.patch[P_U][c + (dj)*FT_W + (di)]
#define PV(di, dj) S.patch[P_V][c + (dj)*FT_W + (di)]
#define PHC(di, dj) dyn_hc<CG>(sm, S, c + (dj)*FT_W + (di))
#define DUT(di, dj) sm.uT[c + (dj)*FT_W + (di)]
#define DVT(di, dj) sm.vT[c + (dj)*FT_W + (di)]
#define DHZ(di, dj) sm.hZ[c + (dj)*FT_W + (di)]
#define DWA(di, dj) sm.wA[c + (dj)*FT_W + (di)]
      double rh;
      if constexpr (CG) rh = cg_hfac(k, kLo, rhLo);
      else rh = S.own[O_RH][co];
      const double gOld = S.own[O_G][co];
      double dp = 0.;
      if constexpr (PHIF) dp = gp * 1. * (sm.phi[c] - sm.phi[c - dP]) * 1.;
      else if (hasPhi) dp = gp * 1. * (S.patch[P_PHI][c] - S.patch[P_PHI][c - dP]) * 1.;
      double fkp = 0.;
      if (below) {      // vertical advective flux at interface k+1 (MOM_U_ADV_WU / MOM_V_ADV_WV)
        const double wA00 = DWA(0, 0), wAn = sm.wA[c - dP];
        const double rT = 0.5 * (wAn + wA00);
        fkp = rT * 0.5 * (fKp1 + fK);
        if (!p.rigidLid) {      // maskC(k+1) - maskC(k) at my cell and at my neighbour
          const double d00 = dyn_dmask<CG>(sm, S, c, k), dn0 = dyn_dmask<CG>(sm, S, c - dP, k);
          fkp = fkp + 0.25 * (wA00 * d00 + wAn * dn0) * fKp1;
        }
      }
      double gt, gD;
      if (ROLE == 0) {
        const double u00 = PU(0, 0), uE = PU(1, 0), uW = PU(-1, 0), uN = PU(0, 1), uS = PU(0, -1);
        {
          const double uT00 = DUT(0, 0);
          const double fzU1 = 0.25 * (uT00 + DUT(1, 0)) * (u00 + uE);
          const double fzU0 = 0.25 * (DUT(-1, 0) + uT00) * (uW + u00);
          const double fmU1 = 0.25 * (DVT(0, 1) + DVT(-1, 1)) * (uN + u00);
          const double fmU0 = 0.25 * (DVT(0, 0) + DVT(-1, 0)) * (u00 + uS);
          gt = -rh * rdrF * r_rA * ((fzU1 - fzU0) * uDudxFac + (fmU1 - fmU0) * uDudxFac + (fkp - fkm) * p.rkSign * uDudxFac);
        }
        {
          const double hZ00 = DHZ(0, 0), hZ01 = DHZ(0, 1);
          const double xv1 = mA0 * drFk * PHC(0, 0) * (-p.viscAhD * (uE - u00) * cf) * mB0;
          const double xv0 = mA1 * drFk * PHC(-1, 0) * (-p.viscAhD * (u00 - uW) * cf) * mB1;
          const double yv1 = mC1 * drFk * hZ01 * (-p.viscAhZ * (uN - u00)) * mD1;
          const double yv0 = mC0 * drFk * hZ00 * (-p.viscAhZ * (u00 - uS)) * mD0;
          double fVrDw = 0.;
          if (!p.implicitViscosity && below) fVrDw = -kapkp1 * rAx * (fKp1 - fK) * p.rkSign * sm.vs.rdrC[k] * mkp1 * mk;
          gD = -rh * rdrF * r_rA * ((xv1 - xv0) * AhFac + (yv1 - yv0) * AhFac + (fVrDw - fVrUp) * p.rkSign * ArFac);
          fVrUp = fVrDw;
          if (p.no_slip_sides) {
            double hWc;
            if constexpr (CG) hWc = cg_hfac(k, kLo, hLo);
            else hWc = S.patch[DP_HW][c];
            const double tu = p.viscAhZ * u00;
            gD = gD + (-rh * rdrF * r_rA * ((hWc - hZ00) * mC0 * mD0 * tu + (hWc - hZ01) * mC1 * mD1 * tu) * drFk * p.sideDragFactor);
          }
          if (p.bottomDragTerms) {
            const double viscFac = p.no_slip_bottom ? 2. : 0.;
            const double recDrC = (k == g.Nr) ? rdrF : sm.vs.rdrC[k];
            double cu = p.bottomDragLinear * 1.;
            if (p.no_slip_bottom && p.bottomVisc_pCell) cu = cu + kapkp1 * recDrC * viscFac * rh;
            else if (p.no_slip_bottom) cu = cu + kapkp1 * recDrC * viscFac;
            if (k == g.Nr) cu = cu * mk;
            else cu = cu * mk * (1. - mkp1);
            gD = gD - cu * u00 * rh * rdrF;
          }
        }
        if (!p.useCDscheme) {
          const double v00 = PV(0, 0), vN = PV(0, 1), vW = PV(-1, 0);
          double uCf;
          if (p.selectCoriScheme >= 2) uCf = 0.5 * (fC00 * 0.5 * (v00 + vN) + fCn * 0.5 * (vW + PV(-1, 1)));
          else uCf = 0.5 * (fC00 + fCn) * 0.25 * (v00 + vN + vW + PV(-1, 1));
          gt = gt + p.cfFacMom * uCf;
        }
      } else {
        const double v00 = PV(0, 0), vE = PV(1, 0), vW = PV(-1, 0), vN = PV(0, 1), vS = PV(0, -1);
        {
          const double vT00 = DVT(0, 0);
          const double fzV1 = 0.25 * (DUT(1, 0) + DUT(1, -1)) * (vE + v00);
          const double fzV0 = 0.25 * (DUT(0, 0) + DUT(0, -1)) * (v00 + vW);
          const double fmV1 = 0.25 * (vT00 + DVT(0, 1)) * (v00 + vN);
          const double fmV0 = 0.25 * (DVT(0, -1) + vT00) * (vS + v00);
          gt = -rh * rdrF * r_rA * ((fzV1 - fzV0) * uDudxFac + (fmV1 - fmV0) * uDudxFac + (fkp - fkm) * p.rkSign * uDudxFac);
        }
        {
          const double hZ00 = DHZ(0, 0), hZ10 = DHZ(1, 0);
          const double xw1 = mC1 * drFk * hZ10 * (-p.viscAhZ * (vE - v00) * cf) * mD1;
          const double xw0 = mC0 * drFk * hZ00 * (-p.viscAhZ * (v00 - vW) * cf) * mD0;
          const double yw1 = mA0 * drFk * PHC(0, 0) * (-p.viscAhD * (vN - v00)) * mB0;
          const double yw0 = mA1 * drFk * PHC(0, -1) * (-p.viscAhD * (v00 - vS)) * mB1;
          double gVrDw = 0.;
          if (!p.implicitViscosity && below) gVrDw = -kapkp1 * rAx * (fKp1 - fK) * p.rkSign * sm.vs.rdrC[k] * mkp1 * mk;
          gD = -rh * rdrF * r_rA * ((xw1 - xw0) * AhFac + (yw1 - yw0) * AhFac + (gVrDw - fVrUp) * p.rkSign * ArFac);
          fVrUp = gVrDw;
          if (p.no_slip_sides) {
            double hSc;
            if constexpr (CG) hSc = cg_hfac(k, kLo, hLo);
            else hSc = S.patch[DP_HS][c];
            const double tv = p.viscAhZ * v00 * cf;
            gD = gD + (-rh * rdrF * r_rA * ((hSc - hZ00) * mC0 * mD0 * tv + (hSc - hZ10) * mC1 * mD1 * tv) * drFk * p.sideDragFactor);
          }
          if (p.bottomDragTerms) {
            const double viscFac = p.no_slip_bottom ? 2. : 0.;
            const double recDrC = (k == g.Nr) ? rdrF : sm.vs.rdrC[k];
            double cv = p.bottomDragLinear * 1.;
            if (p.no_slip_bottom && p.bottomVisc_pCell) cv = cv + kapkp1 * recDrC * viscFac * rh;
            else if (p.no_slip_bottom) cv = cv + kapkp1 * recDrC * viscFac;
            if (k == g.Nr) cv = cv * mk;
            else cv = cv * mk * (1. - mkp1);
            gD = gD - cv * v00 * rh * rdrF;
          }
        }
        if (!p.useCDscheme) {
          const double u00 = PU(0, 0), uE = PU(1, 0), uS = PU(0, -1);
          double vCf;
          if (p.selectCoriScheme >= 2) vCf = -0.5 * (fC00 * 0.5 * (u00 + uE) + fCn * 0.5 * (uS + PU(1, -1)));
          else vCf = -0.5 * (fC00 + fCn) * 0.25 * (u00 + uE + uS + PU(1, -1));
          gt = gt + p.cfFacMom * vCf;
        }
      }
      gt = gt * mk; gD = gD * mk;      // mom_fluxform.F:1044-1051
      gt = gt - 1. * dp;               // timestep.F:120-121, phFac = pfFacMom = 1
      if (dissInAB) gt = gt + gD;
      if (momForcing) {
        double ge = 0.;
        if (k == 1 && forcRange) ge = 0. + sf[s] * sm.vs.rdrF[0] * rh;
        gt = gt + ge;
      }
      const size_t s3 = s + slab * (size_t)(k - 1);
      const double ab = abFac * (gt - gOld);
      gNm1[s3] = gt;
      gt = gt + ab;
      if (!dissInAB) gt = gt + gD;
      gOut[s3] = fK + deltaTMom * (gt + 0.) * mk;
      fkm = fkp;
#undef PU
#undef PV
#undef PHC
#undef DUT
#undef DVT
#undef DHZ
#undef DWA
    }
    fK = fKp1; mk = mkp1;
  }
}

template <int NST, int MINB, bool CG = false, bool PHIF = false>
__global__ void __launch_bounds__(2 * FT_X *FT_Y, MINB)
    dyn_tma_uv_kernel(const __grid_constant__ DynTmaMaps maps, TileGrid g, MomState st, MomPar p, const double *__restrict__ sfU,
                      const double *__restrict__ sfV, double *__restrict__ gU, double *__restrict__ gV, double *__restrict__ guNm1,
                      double *__restrict__ gvNm1, double deltaTMom, double abFac, int momForcing, int dissInAB, int hasPhi,
                      PhiFuse pf) {
  static_assert(!PHIF || CG, "the fused CALC_PHI_HYD exists for the column-geometry form only");
  extern __shared__ __align__(1024) unsigned char dyn_tma_smem_raw[];
  typedef typename std::conditional<CG, DynTmaSmemCG<NST>, DynTmaSmemN<NST>>::type SM;
  SM &sm = *reinterpret_cast<SM *>(dyn_tma_smem_raw);
  constexpr int NT = 2 * FT_X * FT_Y;
  const int tx = threadIdx.x, ty = threadIdx.y, role = threadIdx.z;      // role 0: U, role 1: V (warp-uniform)
  const int t = (role * FT_Y + ty) * FT_X + tx;
  const int i0 = blockIdx.x * FT_X, j0 = blockIdx.y * FT_Y;
  const int i = i0 + tx, j = j0 + ty;
  const bool active = i <= g.sNx + 1 && j <= g.sNy + 1;
  const int c = (ty + 1) * FT_W + (tx + 1);
  const int co = ty * DT_OWN_W + (tx + 1);
  const int x0 = i0 - 1 + g.OLx - 1, y0 = j0 - 1 + g.OLy - 1;
  if (t == 0) {
#pragma unroll
    for (int q = 0; q < NST; q++) mbar_init(&sm.full[q], 1);
    mbar_fence_init();
  }
  stage_vert(sm.vs, g, t, NT);
  if constexpr (PHIF)
    for (int k = 1 + t; k <= g.Nr; k += NT) {      // dRlocM, dRlocP of phihyd_kernel (integr_GeoPot = 2)
      double dRlocM = 0.5 * g.drC[k - 1] * 1.;
      if (k == 1) dRlocM = (pf.rF[0] - pf.rC[0]) * 1.;
      double dRlocP;
      if (k == g.Nr) dRlocP = (pf.rC[k - 1] - pf.rF[k]) * 1.;
      else dRlocP = 0.5 * g.drC[k] * 1.;
      sm.dM[k - 1] = dRlocM; sm.dP[k - 1] = dRlocP; sm.tRf[k - 1] = pf.tRef[k - 1];
    }
  bool hzok = false;
  if (t < FT_N) {
    const int li = t % FT_W, lj = t / FT_W;
    hzok = li >= 1 && lj >= 1;
    const int gi = min(i0 - 1 + li, g.sNx + g.OLx), gj = min(j0 - 1 + lj, g.sNy + g.OLy);
    const size_t q = g.s(gi, gj);
    sm.dyG[t] = g.dyG[q]; sm.dxG[t] = g.dxG[q]; sm.rA[t] = g.rA[q];
    if constexpr (CG) {
      sm.kLW[t] = g.kLowW[q]; sm.hLW[t] = g.hLowW[q]; sm.kLS[t] = g.kLowS[q]; sm.hLS[t] = g.hLowS[q];
      sm.kLC[t] = g.kLowC[q]; sm.hLC[t] = g.hLowC[q];
    }
  }
  __syncthreads();
  if (t == 0) {
    dyn_tma_issue0(maps, sm.st[0], &sm.full[0], x0, y0);
#pragma unroll
    for (int q = 1; q < NST; q++)
      if (q <= g.Nr) dyn_tma_issue(maps, sm.st[q], &sm.full[q], x0, y0, q, g.Nr, hasPhi != 0);
  }
  // the two components run the same level loop, compiled once per component; both sides meet at the named barrier
  if (role == 0)
    dyn_uv_levels<NST, 0, PHIF>(pf, sm, maps, g, st, p, sfU, gU, guNm1, deltaTMom, abFac, momForcing, dissInAB, hasPhi, t, i, j, active, hzok,
                          c, co, x0, y0);
  else
    dyn_uv_levels<NST, 1, PHIF>(pf, sm, maps, g, st, p, sfV, gV, gvNm1, deltaTMom, abFac, momForcing, dissInAB, hasPhi, t, i, j, active, hzok,
                          c, co, x0, y0);
}

// The TMA path additionally needs both horizontal terms on and a 16-byte row pitch.
inline bool dyn_tma_ok(const Geom &g, const MomPar &p) {
  return dyn_fast_ok(g, p) && p.momAdvection && p.momViscosity && g.PX % 2 == 0 && g.OLx % 2 == 0 && tmap_encoder() != nullptr;
}

}  // namespace mg
