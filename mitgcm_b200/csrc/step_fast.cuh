// step_fast.cuh -- shared-memory staged versions of the two 3-D kernels of the resident step.
//
// The generic kernels (step.cu: dyn_kernel / thermo_kernel) re-evaluate every face flux from
// global memory through the point-wise functions of mom.cuh / gad.cuh: correct for every option,
// but ~1900 instructions and ~230 (mostly L1-hit) global loads per cell, latency-bound at 6 % of
// the HBM roofline (profiles/r01_step_kernels_1024x1024x50_baseline.txt).  Here a CTA of 32 x 8
// columns marches in k and stages each level once in shared memory with a one-point rim:
// u, v, the transports uTrans = u*xA, vTrans = v*yA (CALC_ADV_FLOW / mom_fluxform.F:287-327),
// hFacW/S/C, hFacZ (MOM_CALC_HFACZ), w*rA and maskC of the interface below.  Every thread then
// forms its fluxes from shared memory; k-invariant metrics sit in registers; values of its own
// column at k-1 / k+1 are carried between levels.  Each expression is evaluated in exactly the
// order of mom.cuh / gad.cuh, so the results are bit-identical to the generic kernels.
//
// Fast path conditions (checked on the host, otherwise the generic kernels run):
//   dyn   : no biharmonic viscosity, selectBotDragQuadr = -1, selectCoriScheme in {0, 2},
//           no spherical metric terms, OLx, OLy >= 2
//   thermo: centred 2nd-order advection (tempAdvScheme = 2) in all directions, diffK4 = 0
#pragma once
#include <cuda_pipeline.h>
#include "mom.cuh"

namespace mg {

#ifndef FT_YV
#define FT_YV 8
#endif
#ifndef DYNF_MINB
#define DYNF_MINB 2
#endif
constexpr int FT_X = 32, FT_Y = FT_YV, FT_W = FT_X + 2, FT_H = FT_Y + 2, FT_N = FT_W * FT_H;

// Vertical grid arrays staged once per CTA (they were an exposed L2 round trip per level).
constexpr int FT_NRMAX = 160;
struct VertSmem { double drF[FT_NRMAX], rdrF[FT_NRMAX], rdrC[FT_NRMAX + 1]; };
__device__ __forceinline__ void stage_vert(VertSmem &vs, const TileGrid &g, int t, int nthreads) {
  for (int k = t; k < g.Nr; k += nthreads) { vs.drF[k] = g.drF[k]; vs.rdrF[k] = g.recip_drF[k]; }
  for (int k = t; k <= g.Nr; k += nthreads) vs.rdrC[k] = g.recip_drC[k];
}

struct DynSmem {
  double u[FT_N], v[FT_N], uT[FT_N], vT[FT_N], hW[FT_N], hS[FT_N], hC[FT_N], hZ[FT_N], wA[FT_N], mC[2][FT_N];
};

__global__ void __launch_bounds__(FT_X *FT_Y, DYNF_MINB)
    dyn_fast_kernel(TileGrid g, MomState st, MomPar p, const double *__restrict__ sfU, const double *__restrict__ sfV,
                    double *__restrict__ gU, double *__restrict__ gV, double *__restrict__ guNm1,
                    double *__restrict__ gvNm1, double deltaTMom, double abFac, int momForcing, int dissInAB) {
  __shared__ DynSmem sm;
  const int tx = threadIdx.x, ty = threadIdx.y, t = ty * FT_X + tx;
  const int i0 = blockIdx.x * FT_X, j0 = blockIdx.y * FT_Y;     // output range 0..sN+1 (dynamics.F:191-192)
  const int i = i0 + tx, j = j0 + ty;
  const bool active = i <= g.sNx + 1 && j <= g.sNy + 1;
  const int c = (ty + 1) * FT_W + (tx + 1);                     // my cell in the staged tile
#define SM(a, di, dj) sm.a[c + (dj)*FT_W + (di)]
  // staging map: entries e = t and t + 256 of the 34 x 10 tile, clamped into the halo'd slab
  int se[2];
  size_t sg[2];
  bool sv_[2];
#pragma unroll
  for (int r = 0; r < 2; r++) {
    int e = t + r * FT_X * FT_Y;
    sv_[r] = e < FT_N;
    int li = sv_[r] ? e % FT_W : 0, lj = sv_[r] ? e / FT_W : 0;
    int gi = min(i0 - 1 + li, g.sNx + g.OLx), gj = min(j0 - 1 + lj, g.sNy + g.OLy);
    se[r] = e;
    sg[r] = g.s(gi, gj);
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
  const double sfu = sfU[s], sfv = sfV[s];
  const double uDudxFac = p.afFacMom, AhFac = p.vfFacMom, ArFac = p.implicitViscosity ? 0. : p.vfFacMom;

  // own-column values carried between levels
  const size_t slab = g.slab;
  double uKm1 = 0., vKm1 = 0., mWkm1 = 0., mSkm1 = 0.;
  double uK = st.u[s], vK = st.v[s], mWk = g.maskW[s], mSk = g.maskS[s];
  double kapUk = st.kapU[s], kapVk = st.kapV[s];
  // interface k = 1: stage w*rA and maskC of level 1 to form the surface flux (mom_fluxform.F:384-417)
#pragma unroll
  for (int r = 0; r < 2; r++)
    if (sv_[r]) { sm.wA[se[r]] = st.w[sg[r]] * g.rA[sg[r]]; sm.mC[0][se[r]] = g.maskC[sg[r]]; }
  __syncthreads();
  double ukm = 0., vkm = 0.;
  if (p.momAdvection && !p.rigidLid) {
    ukm = (0.5 * (SM(wA, -1, 0) + SM(wA, 0, 0))) * uK;
    vkm = (0.5 * (SM(wA, 0, -1) + SM(wA, 0, 0))) * vK;
  }
  __syncthreads();
  int cur = 0;   // sm.mC[cur] = maskC(k), sm.mC[cur^1] = maskC(k+1)
  for (int k = 1; k <= g.Nr; k++) {
    const size_t ko = slab * (size_t)(k - 1);
    const double drFk = g.drF[k - 1], rdrF = g.recip_drF[k - 1];
    const bool below = k + 1 <= g.Nr;
    // own column: issue every global load of this level up front, so one memory latency is exposed
    // per level (the compiler cannot hoist loads over the barriers below)
    const size_t s3 = s + ko;
    double uKp1 = 0., vKp1 = 0., mWkp1 = 0., mSkp1 = 0.;
    const double kapUkp1 = st.kapU[s3 + slab], kapVkp1 = st.kapV[s3 + slab];
    if (below) { uKp1 = st.u[s3 + slab]; vKp1 = st.v[s3 + slab]; mWkp1 = g.maskW[s3 + slab]; mSkp1 = g.maskS[s3 + slab]; }
    const double rhW = g.recip_hFacW[s3], rhS = g.recip_hFacS[s3];
    const double guOld = guNm1[s3], gvOld = gvNm1[s3];
    // ---- stage level k (and interface k+1) ----
#pragma unroll
    for (int r = 0; r < 2; r++)
      if (sv_[r]) {
        const size_t q = sg[r] + ko;
        const double u = st.u[q], v = st.v[q], hW = g.hFacW[q], hS = g.hFacS[q];
        sm.u[se[r]] = u; sm.v[se[r]] = v; sm.hW[se[r]] = hW; sm.hS[se[r]] = hS; sm.hC[se[r]] = g.hFacC[q];
        sm.uT[se[r]] = u * (g.dyG[sg[r]] * drFk * hW);
        sm.vT[se[r]] = v * (g.dxG[sg[r]] * drFk * hS);
        if (below) { sm.wA[se[r]] = st.w[q + slab] * g.rA[sg[r]]; sm.mC[cur ^ 1][se[r]] = g.maskC[q + slab]; }
      }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 2; r++)
      if (sv_[r]) {     // MOM_CALC_HFACZ at the south-west corner of every staged cell that has its neighbours
        int e = se[r], li = e % FT_W, lj = e / FT_W;
        if (li >= 1 && lj >= 1) {
          double h = fmin(sm.hW[e], sm.hW[e - FT_W]);
          h = fmin(sm.hS[e], h);
          h = fmin(sm.hS[e - 1], h);
          sm.hZ[e] = h;
        }
      }
    __syncthreads();
    if (active) {
      const double u00 = SM(u, 0, 0), v00 = SM(v, 0, 0);
      // vertical advective fluxes at interface k+1 (MOM_U_ADV_WU / MOM_V_ADV_WV)
      double ukp = 0., vkp = 0.;
      if (p.momAdvection && below) {
        const double rTU = 0.5 * (SM(wA, -1, 0) + SM(wA, 0, 0)), rTV = 0.5 * (SM(wA, 0, -1) + SM(wA, 0, 0));
        ukp = rTU * 0.5 * (uKp1 + uK);
        vkp = rTV * 0.5 * (vKp1 + vK);
        if (!p.rigidLid) {
          const double d00 = sm.mC[cur ^ 1][c] - sm.mC[cur][c];
          ukp = ukp + 0.25 * (SM(wA, 0, 0) * d00 + SM(wA, -1, 0) * (sm.mC[cur ^ 1][c - 1] - sm.mC[cur][c - 1])) * uKp1;
          vkp = vkp + 0.25 * (SM(wA, 0, 0) * d00 + SM(wA, 0, -1) * (sm.mC[cur ^ 1][c - FT_W] - sm.mC[cur][c - FT_W])) * vKp1;
        }
      }
      double gu = 0., gv = 0., guD = 0., gvD = 0.;
      if (p.momAdvection) {
        const double fzU1 = 0.25 * (SM(uT, 0, 0) + SM(uT, 1, 0)) * (u00 + SM(u, 1, 0));
        const double fzU0 = 0.25 * (SM(uT, -1, 0) + SM(uT, 0, 0)) * (SM(u, -1, 0) + u00);
        const double fmU1 = 0.25 * (SM(vT, 0, 1) + SM(vT, -1, 1)) * (SM(u, 0, 1) + u00);
        const double fmU0 = 0.25 * (SM(vT, 0, 0) + SM(vT, -1, 0)) * (u00 + SM(u, 0, -1));
        gu = -rhW * rdrF * r_rAw * ((fzU1 - fzU0) * uDudxFac + (fmU1 - fmU0) * uDudxFac + (ukp - ukm) * p.rkSign * uDudxFac);
        const double fzV1 = 0.25 * (SM(uT, 1, 0) + SM(uT, 1, -1)) * (SM(v, 1, 0) + v00);
        const double fzV0 = 0.25 * (SM(uT, 0, 0) + SM(uT, 0, -1)) * (v00 + SM(v, -1, 0));
        const double fmV1 = 0.25 * (SM(vT, 0, 0) + SM(vT, 0, 1)) * (v00 + SM(v, 0, 1));
        const double fmV0 = 0.25 * (SM(vT, 0, -1) + SM(vT, 0, 0)) * (SM(v, 0, -1) + v00);
        gv = -rhS * rdrF * r_rAs * ((fzV1 - fzV0) * uDudxFac + (fmV1 - fmV0) * uDudxFac + (vkp - vkm) * p.rkSign * uDudxFac);
      }
      if (p.momViscosity) {
        const double hZ00 = SM(hZ, 0, 0), hZ01 = SM(hZ, 0, 1), hZ10 = SM(hZ, 1, 0);
        // U: MOM_U_XVISCFLUX, MOM_U_YVISCFLUX, MOM_U_RVISCFLUX
        const double xv1 = dyF00 * drFk * SM(hC, 0, 0) * (-p.viscAhD * (SM(u, 1, 0) - u00) * cfU + p.viscA4D * (0. - 0.) * cfU) * rdxF00;
        const double xv0 = dyFm0 * drFk * SM(hC, -1, 0) * (-p.viscAhD * (u00 - SM(u, -1, 0)) * cfU + p.viscA4D * (0. - 0.) * cfU) * rdxFm0;
        const double yv1 = dxV01 * drFk * hZ01 * (-p.viscAhZ * (SM(u, 0, 1) - u00) + p.viscA4Z * (0. - 0.)) * rdyU01;
        const double yv0 = dxV00 * drFk * hZ00 * (-p.viscAhZ * (u00 - SM(u, 0, -1)) + p.viscA4Z * (0. - 0.)) * rdyU00;
        double fVrUp = 0., fVrDw = 0.;
        if (!p.implicitViscosity) {
          if (k > 1) fVrUp = -kapUk * rAw * (uK - uKm1) * p.rkSign * g.recip_drC[k - 1] * mWk * mWkm1;
          if (below) fVrDw = -kapUkp1 * rAw * (uKp1 - uK) * p.rkSign * g.recip_drC[k] * mWkp1 * mWk;
        }
        guD = -rhW * rdrF * r_rAw * ((xv1 - xv0) * AhFac + (yv1 - yv0) * AhFac + (fVrDw - fVrUp) * p.rkSign * ArFac);
        // V: MOM_V_XVISCFLUX, MOM_V_YVISCFLUX, MOM_V_RVISCFLUX
        const double xw1 = dyU10 * drFk * hZ10 * (-p.viscAhZ * (SM(v, 1, 0) - v00) * cfV + p.viscA4Z * (0. - 0.) * cfV) * rdxV10;
        const double xw0 = dyU00 * drFk * hZ00 * (-p.viscAhZ * (v00 - SM(v, -1, 0)) * cfV + p.viscA4Z * (0. - 0.) * cfV) * rdxV00;
        const double yw1 = dxF00 * drFk * SM(hC, 0, 0) * (-p.viscAhD * (SM(v, 0, 1) - v00) + p.viscA4D * (0. - 0.)) * rdyF00;
        const double yw0 = dxF0m * drFk * SM(hC, 0, -1) * (-p.viscAhD * (v00 - SM(v, 0, -1)) + p.viscA4D * (0. - 0.)) * rdyF0m;
        double gVrUp = 0., gVrDw = 0.;
        if (!p.implicitViscosity) {
          if (k > 1) gVrUp = -kapVk * rAs * (vK - vKm1) * p.rkSign * g.recip_drC[k - 1] * mSk * mSkm1;
          if (below) gVrDw = -kapVkp1 * rAs * (vKp1 - vK) * p.rkSign * g.recip_drC[k] * mSkp1 * mSk;
        }
        gvD = -rhS * rdrF * r_rAs * ((xw1 - xw0) * AhFac + (yw1 - yw0) * AhFac + (gVrDw - gVrUp) * p.rkSign * ArFac);
        if (p.no_slip_sides) {   // MOM_U_SIDEDRAG / MOM_V_SIDEDRAG
          const double hWc = SM(hW, 0, 0), hSc = SM(hS, 0, 0);
          const double tu = p.viscAhZ * u00 - p.viscA4Z * 0.;
          guD = guD + (-rhW * rdrF * r_rAw * ((hWc - hZ00) * dxV00 * rdyU00 * tu + (hWc - hZ01) * dxV01 * rdyU01 * tu) * drFk * p.sideDragFactor);
          const double tv = p.viscAhZ * v00 * cfV - p.viscA4Z * 0. * cfV;
          gvD = gvD + (-rhS * rdrF * r_rAs * ((hSc - hZ00) * dyU00 * rdxV00 * tv + (hSc - hZ10) * dyU10 * rdxV10 * tv) * drFk * p.sideDragFactor);
        }
        if (p.bottomDragTerms) {  // MOM_{U,V}_BOTDRAG_COEFF with selectBotDragQuadr = -1
          const double viscFac = p.no_slip_bottom ? 2. : 0.;
          const double recDrC = (k == g.Nr) ? rdrF : g.recip_drC[k];
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
          uCf = 0.5 * (fC00 * 0.5 * (v00 + SM(v, 0, 1)) + fCm0 * 0.5 * (SM(v, -1, 0) + SM(v, -1, 1)));
          vCf = -0.5 * (fC00 * 0.5 * (u00 + SM(u, 1, 0)) + fC0m * 0.5 * (SM(u, 0, -1) + SM(u, 1, -1)));
        } else {
          uCf = 0.5 * (fC00 + fCm0) * 0.25 * (v00 + SM(v, 0, 1) + SM(v, -1, 0) + SM(v, -1, 1));
          vCf = -0.5 * (fC00 + fC0m) * 0.25 * (u00 + SM(u, 1, 0) + SM(u, 0, -1) + SM(u, 1, -1));
        }
        gu = gu + p.cfFacMom * uCf;
        gv = gv + p.cfFacMom * vCf;
      }
      gu = gu * mWk; guD = guD * mWk; gv = gv * mSk; gvD = gvD * mSk;   // mom_fluxform.F:1044-1051
      // ---- TIMESTEP (timestep.F:95-385), as in dyn_kernel ----
      gu = gu - 0.; gv = gv - 0.;
      if (p.momViscosity && dissInAB) { gu = gu + guD; gv = gv + gvD; }
      if (momForcing) {
        double ge = 0., he = 0.;
        if (k == 1) {
          if (i >= 1 && i <= g.sNx + 1) ge = 0. + sfu * g.recip_drF[0] * rhW;
          if (j >= 1 && j <= g.sNy + 1) he = 0. + sfv * g.recip_drF[0] * rhS;
        }
        gu = gu + ge; gv = gv + he;
      }
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
    }
    uKm1 = uK; uK = uKp1; vKm1 = vK; vK = vKp1; mWkm1 = mWk; mWk = mWkp1; mSkm1 = mSk; mSk = mSkp1;
    kapUk = kapUkp1; kapVk = kapVkp1;
    cur ^= 1;
    __syncthreads();
  }
#undef SM
}

// ---- pipelined variant: the next level is fetched with cp.async while the current one is computed ----
enum { RU = 0, RV, RHW, RHS, RHC, RW, RMC, NRAW };
struct DynPipeSmem {
  double raw[2][NRAW][FT_N];     // ring of raw levels: u, v, hFacW, hFacS, hFacC of level k; w, maskC of level k+1
  double uT[FT_N], vT[FT_N], hZ[FT_N], wA[FT_N], mCk[FT_N];
  double dyG[FT_N], dxG[FT_N], rA[FT_N];
};

__device__ __forceinline__ void dyn_pipe_prefetch(DynPipeSmem &sm, int slot, int e, size_t sg, const TileGrid &g,
                                                  const MomState &st, size_t slab, int k, bool below) {
  const size_t q = sg + slab * (size_t)(k - 1);
  __pipeline_memcpy_async(&sm.raw[slot][RU][e], st.u + q, 8);
  __pipeline_memcpy_async(&sm.raw[slot][RV][e], st.v + q, 8);
  __pipeline_memcpy_async(&sm.raw[slot][RHW][e], g.hFacW + q, 8);
  __pipeline_memcpy_async(&sm.raw[slot][RHS][e], g.hFacS + q, 8);
  __pipeline_memcpy_async(&sm.raw[slot][RHC][e], g.hFacC + q, 8);
  if (below) {
    __pipeline_memcpy_async(&sm.raw[slot][RW][e], st.w + q + slab, 8);
    __pipeline_memcpy_async(&sm.raw[slot][RMC][e], g.maskC + q + slab, 8);
  }
}

#ifndef DYNP_MINB
#define DYNP_MINB 2
#endif
__global__ void __launch_bounds__(FT_X *FT_Y, DYNP_MINB)
    dyn_pipe_kernel(TileGrid g, MomState st, MomPar p, const double *__restrict__ sfU, const double *__restrict__ sfV,
                    double *__restrict__ gU, double *__restrict__ gV, double *__restrict__ guNm1,
                    double *__restrict__ gvNm1, double deltaTMom, double abFac, int momForcing, int dissInAB,
                    const double *__restrict__ phiHyd) {
  extern __shared__ __align__(16) unsigned char dyn_pipe_smem[];
  DynPipeSmem &sm = *reinterpret_cast<DynPipeSmem *>(dyn_pipe_smem);
  __shared__ VertSmem vs;
  const int tx = threadIdx.x, ty = threadIdx.y, t = ty * FT_X + tx;
  stage_vert(vs, g, t, FT_X * FT_Y);
  const int i0 = blockIdx.x * FT_X, j0 = blockIdx.y * FT_Y;     // output range 0..sN+1 (dynamics.F:191-192)
  const int i = i0 + tx, j = j0 + ty;
  const bool active = i <= g.sNx + 1 && j <= g.sNy + 1;
  const int c = (ty + 1) * FT_W + (tx + 1);                     // my cell in the staged tile
#define SM(a, di, dj) SMX_##a(c + (dj)*FT_W + (di))
#define SMX_u(x) sm.raw[rb][RU][x]
#define SMX_v(x) sm.raw[rb][RV][x]
#define SMX_hW(x) sm.raw[rb][RHW][x]
#define SMX_hS(x) sm.raw[rb][RHS][x]
#define SMX_hC(x) sm.raw[rb][RHC][x]
#define SMX_uT(x) sm.uT[x]
#define SMX_vT(x) sm.vT[x]
#define SMX_hZ(x) sm.hZ[x]
#define SMX_wA(x) sm.wA[x]
  // staging map: entries e = t and t + 256 of the 34 x 10 tile, clamped into the halo'd slab
  int se[2];
  size_t sg[2];
  bool sv_[2];
#pragma unroll
  for (int r = 0; r < 2; r++) {
    int e = t + r * FT_X * FT_Y;
    sv_[r] = e < FT_N;
    int li = sv_[r] ? e % FT_W : 0, lj = sv_[r] ? e / FT_W : 0;
    int gi = min(i0 - 1 + li, g.sNx + g.OLx), gj = min(j0 - 1 + lj, g.sNy + g.OLy);
    se[r] = e;
    sg[r] = g.s(gi, gj);
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
  const double sfu = sfU[s], sfv = sfV[s];
  const double uDudxFac = p.afFacMom, AhFac = p.vfFacMom, ArFac = p.implicitViscosity ? 0. : p.vfFacMom;
  // CALC_GRAD_PHI_HYD (calc_grad_phi_hyd.F:150-165) is defined on i >= iMin+1, j >= jMin+1
  const double gpx = (phiHyd && i >= 1) ? g.recip_dxC[s] : 0., gpy = (phiHyd && j >= 1) ? g.recip_dyC[s] : 0.;

  // own-column values carried between levels
  const size_t slab = g.slab;
  double uKm1 = 0., vKm1 = 0., mWkm1 = 0., mSkm1 = 0.;
  double uK = st.u[s], vK = st.v[s], mWk = g.maskW[s], mSk = g.maskS[s];
  double kapUk = st.kapU[s], kapVk = st.kapV[s];
  // k-invariant metrics of the staged cells, the surface interface (w*rA, maskC of level 1:
  // mom_fluxform.F:384-417), and the asynchronous prefetch of level 1 into ring slot 1
  int rb = 1;
#pragma unroll
  for (int r = 0; r < 2; r++)
    if (sv_[r]) {
      sm.dyG[se[r]] = g.dyG[sg[r]]; sm.dxG[se[r]] = g.dxG[sg[r]]; sm.rA[se[r]] = g.rA[sg[r]];
      sm.wA[se[r]] = st.w[sg[r]] * g.rA[sg[r]];
      sm.raw[0][RMC][se[r]] = g.maskC[sg[r]];
      dyn_pipe_prefetch(sm, 1, se[r], sg[r], g, st, slab, 1, 1 + 1 <= g.Nr);
    }
  __pipeline_commit();
  __syncthreads();
  double ukm = 0., vkm = 0.;
  if (p.momAdvection && !p.rigidLid) {
    ukm = (0.5 * (SM(wA, -1, 0) + SM(wA, 0, 0))) * uK;
    vkm = (0.5 * (SM(wA, 0, -1) + SM(wA, 0, 0))) * vK;
  }
  __syncthreads();
  for (int k = 1; k <= g.Nr; k++) {
    const size_t ko = slab * (size_t)(k - 1);
    const double drFk = vs.drF[k - 1], rdrF = vs.rdrF[k - 1];
    const bool below = k + 1 <= g.Nr;
    // own column: issue every global load of this level up front, so one memory latency is exposed
    // per level (the compiler cannot hoist loads over the barriers below)
    const size_t s3 = s + ko;
    double uKp1 = 0., vKp1 = 0., mWkp1 = 0., mSkp1 = 0.;
    const double kapUkp1 = st.kapU[s3 + slab], kapVkp1 = st.kapV[s3 + slab];
    if (below) { uKp1 = st.u[s3 + slab]; vKp1 = st.v[s3 + slab]; mWkp1 = g.maskW[s3 + slab]; mSkp1 = g.maskS[s3 + slab]; }
    const double rhW = g.recip_hFacW[s3], rhS = g.recip_hFacS[s3];
    const double guOld = guNm1[s3], gvOld = gvNm1[s3];
    double dpx = 0., dpy = 0.;
    if (phiHyd) {
      const double ph = phiHyd[s3];
      dpx = gpx * 1. * (ph - phiHyd[s3 - 1]) * 1.;
      dpy = gpy * 1. * (ph - phiHyd[s3 - PX]) * 1.;
    }
    // ---- level k has been prefetched into ring slot rb (cp.async); derive what the fluxes share ----
    __pipeline_wait_prior(0);
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 2; r++)
      if (sv_[r]) {
        const int e = se[r], li = e % FT_W, lj = e / FT_W;
        const double hW = sm.raw[rb][RHW][e], hS = sm.raw[rb][RHS][e];
        sm.uT[e] = sm.raw[rb][RU][e] * (sm.dyG[e] * drFk * hW);
        sm.vT[e] = sm.raw[rb][RV][e] * (sm.dxG[e] * drFk * hS);
        if (below) sm.wA[e] = sm.raw[rb][RW][e] * sm.rA[e];
        sm.mCk[e] = sm.raw[rb ^ 1][RMC][e];          // maskC(k), fetched with level k-1
        if (li >= 1 && lj >= 1) {                    // MOM_CALC_HFACZ at the south-west corner
          double h = fmin(hW, sm.raw[rb][RHW][e - FT_W]);
          h = fmin(hS, h);
          h = fmin(sm.raw[rb][RHS][e - 1], h);
          sm.hZ[e] = h;
        }
      }
    __syncthreads();
    if (below) {      // prefetch level k+1 into the other slot while this level is computed
#pragma unroll
      for (int r = 0; r < 2; r++)
        if (sv_[r]) dyn_pipe_prefetch(sm, rb ^ 1, se[r], sg[r], g, st, slab, k + 1, k + 2 <= g.Nr);
    }
    __pipeline_commit();
    if (active) {
      const double u00 = SM(u, 0, 0), v00 = SM(v, 0, 0);
      // vertical advective fluxes at interface k+1 (MOM_U_ADV_WU / MOM_V_ADV_WV)
      double ukp = 0., vkp = 0.;
      if (p.momAdvection && below) {
        const double rTU = 0.5 * (SM(wA, -1, 0) + SM(wA, 0, 0)), rTV = 0.5 * (SM(wA, 0, -1) + SM(wA, 0, 0));
        ukp = rTU * 0.5 * (uKp1 + uK);
        vkp = rTV * 0.5 * (vKp1 + vK);
        if (!p.rigidLid) {
          const double d00 = sm.raw[rb][RMC][c] - sm.mCk[c];
          ukp = ukp + 0.25 * (SM(wA, 0, 0) * d00 + SM(wA, -1, 0) * (sm.raw[rb][RMC][c - 1] - sm.mCk[c - 1])) * uKp1;
          vkp = vkp + 0.25 * (SM(wA, 0, 0) * d00 + SM(wA, 0, -1) * (sm.raw[rb][RMC][c - FT_W] - sm.mCk[c - FT_W])) * vKp1;
        }
      }
      double gu = 0., gv = 0., guD = 0., gvD = 0.;
      if (p.momAdvection) {
        const double fzU1 = 0.25 * (SM(uT, 0, 0) + SM(uT, 1, 0)) * (u00 + SM(u, 1, 0));
        const double fzU0 = 0.25 * (SM(uT, -1, 0) + SM(uT, 0, 0)) * (SM(u, -1, 0) + u00);
        const double fmU1 = 0.25 * (SM(vT, 0, 1) + SM(vT, -1, 1)) * (SM(u, 0, 1) + u00);
        const double fmU0 = 0.25 * (SM(vT, 0, 0) + SM(vT, -1, 0)) * (u00 + SM(u, 0, -1));
        gu = -rhW * rdrF * r_rAw * ((fzU1 - fzU0) * uDudxFac + (fmU1 - fmU0) * uDudxFac + (ukp - ukm) * p.rkSign * uDudxFac);
        const double fzV1 = 0.25 * (SM(uT, 1, 0) + SM(uT, 1, -1)) * (SM(v, 1, 0) + v00);
        const double fzV0 = 0.25 * (SM(uT, 0, 0) + SM(uT, 0, -1)) * (v00 + SM(v, -1, 0));
        const double fmV1 = 0.25 * (SM(vT, 0, 0) + SM(vT, 0, 1)) * (v00 + SM(v, 0, 1));
        const double fmV0 = 0.25 * (SM(vT, 0, -1) + SM(vT, 0, 0)) * (SM(v, 0, -1) + v00);
        gv = -rhS * rdrF * r_rAs * ((fzV1 - fzV0) * uDudxFac + (fmV1 - fmV0) * uDudxFac + (vkp - vkm) * p.rkSign * uDudxFac);
      }
      if (p.momViscosity) {
        const double hZ00 = SM(hZ, 0, 0), hZ01 = SM(hZ, 0, 1), hZ10 = SM(hZ, 1, 0);
        // U: MOM_U_XVISCFLUX, MOM_U_YVISCFLUX, MOM_U_RVISCFLUX
        const double xv1 = dyF00 * drFk * SM(hC, 0, 0) * (-p.viscAhD * (SM(u, 1, 0) - u00) * cfU + p.viscA4D * (0. - 0.) * cfU) * rdxF00;
        const double xv0 = dyFm0 * drFk * SM(hC, -1, 0) * (-p.viscAhD * (u00 - SM(u, -1, 0)) * cfU + p.viscA4D * (0. - 0.) * cfU) * rdxFm0;
        const double yv1 = dxV01 * drFk * hZ01 * (-p.viscAhZ * (SM(u, 0, 1) - u00) + p.viscA4Z * (0. - 0.)) * rdyU01;
        const double yv0 = dxV00 * drFk * hZ00 * (-p.viscAhZ * (u00 - SM(u, 0, -1)) + p.viscA4Z * (0. - 0.)) * rdyU00;
        double fVrUp = 0., fVrDw = 0.;
        if (!p.implicitViscosity) {
          if (k > 1) fVrUp = -kapUk * rAw * (uK - uKm1) * p.rkSign * vs.rdrC[k - 1] * mWk * mWkm1;
          if (below) fVrDw = -kapUkp1 * rAw * (uKp1 - uK) * p.rkSign * vs.rdrC[k] * mWkp1 * mWk;
        }
        guD = -rhW * rdrF * r_rAw * ((xv1 - xv0) * AhFac + (yv1 - yv0) * AhFac + (fVrDw - fVrUp) * p.rkSign * ArFac);
        // V: MOM_V_XVISCFLUX, MOM_V_YVISCFLUX, MOM_V_RVISCFLUX
        const double xw1 = dyU10 * drFk * hZ10 * (-p.viscAhZ * (SM(v, 1, 0) - v00) * cfV + p.viscA4Z * (0. - 0.) * cfV) * rdxV10;
        const double xw0 = dyU00 * drFk * hZ00 * (-p.viscAhZ * (v00 - SM(v, -1, 0)) * cfV + p.viscA4Z * (0. - 0.) * cfV) * rdxV00;
        const double yw1 = dxF00 * drFk * SM(hC, 0, 0) * (-p.viscAhD * (SM(v, 0, 1) - v00) + p.viscA4D * (0. - 0.)) * rdyF00;
        const double yw0 = dxF0m * drFk * SM(hC, 0, -1) * (-p.viscAhD * (v00 - SM(v, 0, -1)) + p.viscA4D * (0. - 0.)) * rdyF0m;
        double gVrUp = 0., gVrDw = 0.;
        if (!p.implicitViscosity) {
          if (k > 1) gVrUp = -kapVk * rAs * (vK - vKm1) * p.rkSign * vs.rdrC[k - 1] * mSk * mSkm1;
          if (below) gVrDw = -kapVkp1 * rAs * (vKp1 - vK) * p.rkSign * vs.rdrC[k] * mSkp1 * mSk;
        }
        gvD = -rhS * rdrF * r_rAs * ((xw1 - xw0) * AhFac + (yw1 - yw0) * AhFac + (gVrDw - gVrUp) * p.rkSign * ArFac);
        if (p.no_slip_sides) {   // MOM_U_SIDEDRAG / MOM_V_SIDEDRAG
          const double hWc = SM(hW, 0, 0), hSc = SM(hS, 0, 0);
          const double tu = p.viscAhZ * u00 - p.viscA4Z * 0.;
          guD = guD + (-rhW * rdrF * r_rAw * ((hWc - hZ00) * dxV00 * rdyU00 * tu + (hWc - hZ01) * dxV01 * rdyU01 * tu) * drFk * p.sideDragFactor);
          const double tv = p.viscAhZ * v00 * cfV - p.viscA4Z * 0. * cfV;
          gvD = gvD + (-rhS * rdrF * r_rAs * ((hSc - hZ00) * dyU00 * rdxV00 * tv + (hSc - hZ10) * dyU10 * rdxV10 * tv) * drFk * p.sideDragFactor);
        }
        if (p.bottomDragTerms) {  // MOM_{U,V}_BOTDRAG_COEFF with selectBotDragQuadr = -1
          const double viscFac = p.no_slip_bottom ? 2. : 0.;
          const double recDrC = (k == g.Nr) ? rdrF : vs.rdrC[k];
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
          uCf = 0.5 * (fC00 * 0.5 * (v00 + SM(v, 0, 1)) + fCm0 * 0.5 * (SM(v, -1, 0) + SM(v, -1, 1)));
          vCf = -0.5 * (fC00 * 0.5 * (u00 + SM(u, 1, 0)) + fC0m * 0.5 * (SM(u, 0, -1) + SM(u, 1, -1)));
        } else {
          uCf = 0.5 * (fC00 + fCm0) * 0.25 * (v00 + SM(v, 0, 1) + SM(v, -1, 0) + SM(v, -1, 1));
          vCf = -0.5 * (fC00 + fC0m) * 0.25 * (u00 + SM(u, 1, 0) + SM(u, 0, -1) + SM(u, 1, -1));
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
          if (i >= 1 && i <= g.sNx + 1) ge = 0. + sfu * vs.rdrF[0] * rhW;
          if (j >= 1 && j <= g.sNy + 1) he = 0. + sfv * vs.rdrF[0] * rhS;
        }
        gu = gu + ge; gv = gv + he;
      }
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
    }
    uKm1 = uK; uK = uKp1; vKm1 = vK; vK = vKp1; mWkm1 = mWk; mWk = mWkp1; mSkm1 = mSk; mSk = mSkp1;
    kapUk = kapUkp1; kapVk = kapVkp1;
    rb ^= 1;
  }
#undef SM
#undef SMX_u
#undef SMX_v
#undef SMX_hW
#undef SMX_hS
#undef SMX_hC
#undef SMX_uT
#undef SMX_vT
#undef SMX_hZ
#undef SMX_wA
}

struct ThermoSmem {
  double T[FT_N], xA[FT_N], yA[FT_N], uT[FT_N], vT[FT_N];
};

// TEMP_INTEGRATE for centred 2nd-order advection + Laplacian diffusion, k = Nr..1 (see thermo_kernel).
#ifndef THF_MINB
#define THF_MINB 3
#endif
__global__ void __launch_bounds__(FT_X *FT_Y, THF_MINB)
    thermo_fast_kernel(TileGrid g, const double *__restrict__ u, const double *__restrict__ v, const double *__restrict__ w,
                       const double *__restrict__ theta, const double *__restrict__ kapT, double *__restrict__ thetaNew,
                       double *__restrict__ gtNm1, GadPar p, double abFac, const double *__restrict__ sfT) {
  __shared__ ThermoSmem sm;
  __shared__ VertSmem vs;
  const int tx = threadIdx.x, ty = threadIdx.y, t = ty * FT_X + tx;
  stage_vert(vs, g, t, FT_X * FT_Y);
  __syncthreads();
  const int i0 = 1 + blockIdx.x * FT_X, j0 = 1 + blockIdx.y * FT_Y;
  const int i = i0 + tx, j = j0 + ty;
  const bool active = i <= g.sNx && j <= g.sNy;
  const int c = (ty + 1) * FT_W + (tx + 1);
#define SM(a, di, dj) sm.a[c + (dj)*FT_W + (di)]
  int se[2];
  size_t sg[2];
  bool sv_[2];
#pragma unroll
  for (int r = 0; r < 2; r++) {
    int e = t + r * FT_X * FT_Y;
    sv_[r] = e < FT_N;
    int li = sv_[r] ? e % FT_W : 0, lj = sv_[r] ? e / FT_W : 0;
    int gi = min(i0 - 1 + li, g.sNx + g.OLx), gj = min(j0 - 1 + lj, g.sNy + g.OLy);
    se[r] = e;
    sg[r] = g.s(gi, gj);
  }
  const size_t s = active ? g.s(i, j) : g.s(1, 1);
  const size_t slab = g.slab;
  const int PX = g.PX;
  const double rdxC0 = g.recip_dxC[s], rdxC1 = g.recip_dxC[s + 1], rdyC0 = g.recip_dyC[s], rdyC1 = g.recip_dyC[s + PX];
  const double rA = g.rA[s], r_rA = g.recip_rA[s], cfU = g.cosFacU[active ? j + g.OLy - 1 : 0];
  const double advFac = p.calcAdvection ? 1. : 0.;
  double rAdvFac = p.rkSign * advFac;
  if (p.implicitAdvection) rAdvFac = p.rkSign;
  double fVdn = 0., rTransKp1 = 0.;
  for (int k = g.Nr; k >= 1; k--) {
    const size_t ko = slab * (size_t)(k - 1);
    const double drFk = vs.drF[k - 1];
    // own-column loads first: one exposed memory latency per level
    const size_t s3 = s + ko;
    const double Tkm1 = k >= 2 ? theta[s3 - slab] : 0., mC = g.maskC[s3], mCm1 = k >= 2 ? g.maskC[s3 - slab] : 0.;
    const double wK = w[s3], kapK = kapT[s3], rhC = g.recip_hFacC[s3], gtOld = gtNm1[s3];
#pragma unroll
    for (int r = 0; r < 2; r++)
      if (sv_[r]) {
        const size_t q = sg[r] + ko;
        const double xA = g.dyG[sg[r]] * drFk * g.hFacW[q], yA = g.dxG[sg[r]] * drFk * g.hFacS[q];
        sm.T[se[r]] = theta[q]; sm.xA[se[r]] = xA; sm.yA[se[r]] = yA;
        sm.uT[se[r]] = u[q] * xA; sm.vT[se[r]] = v[q] * yA;
      }
    __syncthreads();
    if (active) {
      const double T00 = SM(T, 0, 0);
      // X / Y fluxes (GAD_C2_ADV_X/Y, GAD_DIFF_X/Y), west+east and south+north faces
      double fz0 = 0., fz1 = 0., fm0 = 0., fm1 = 0.;
      if (p.calcAdvection) {
        fz0 = fz0 + SM(uT, 0, 0) * (T00 + SM(T, -1, 0)) * 0.5;
        fz1 = fz1 + SM(uT, 1, 0) * (SM(T, 1, 0) + T00) * 0.5;
        fm0 = fm0 + SM(vT, 0, 0) * (T00 + SM(T, 0, -1)) * 0.5;
        fm1 = fm1 + SM(vT, 0, 1) * (SM(T, 0, 1) + T00) * 0.5;
      }
      double dz0 = 0., dz1 = 0., dm0 = 0., dm1 = 0.;
      if (p.diffKh != 0.) {
        dz0 = -p.diffKh * SM(xA, 0, 0) * rdxC0 * (T00 - SM(T, -1, 0)) * cfU;
        dz1 = -p.diffKh * SM(xA, 1, 0) * rdxC1 * (SM(T, 1, 0) - T00) * cfU;
        dm0 = -p.diffKh * SM(yA, 0, 0) * rdyC0 * (T00 - SM(T, 0, -1));
        dm1 = -p.diffKh * SM(yA, 0, 1) * rdyC1 * (SM(T, 0, 1) - T00);
      }
      fz0 = fz0 + dz0; fz1 = fz1 + dz1; fm0 = fm0 + dm0; fm1 = fm1 + dm1;
      // vertical flux at the upper interface of level k (GAD_C2_ADV_R, GAD_DIFF_R)
      double fvu = 0., rTrans = 0.;
      if (k >= 2) {
        const double maskUp = mCm1 * mC;
        rTrans = wK * rA * maskUp;
        if (p.calcAdvection && !p.implicitAdvection) fvu = fvu + mCm1 * rTrans * (T00 + Tkm1) * 0.5;
        double df = 0.;
        if (!p.implicitDiffusion) df = -kapK * maskUp * rA * vs.rdrC[k - 1] * (T00 - Tkm1) * p.rkSign;
        fvu = fvu + df;
      }
      double gT = 0. - rhC * vs.rdrF[k - 1] * r_rA *
                           ((fz1 - fz0) + (fm1 - fm0) + (fVdn - fvu) * p.rkSign -
                            T00 * ((SM(uT, 1, 0) - SM(uT, 0, 0)) * advFac + (SM(vT, 0, 1) - SM(vT, 0, 0)) * advFac +
                                   (rTransKp1 - rTrans) * rAdvFac));
      if (sfT) {   // APPLY_FORCING_T at k = kSurface, inside Adams-Bashforth (tracForcingOutAB = 0)
        double gtForc = 0.;
        if (k == 1) gtForc = gtForc + sfT[g.s(i, j)] * vs.rdrF[0] * rhC;
        gT = gT + gtForc;
      }
      const double ab = abFac * (gT - gtOld);
      gtNm1[s3] = gT;
      gT = gT + ab;
      thetaNew[s3] = T00 + p.deltaT * gT;
      fVdn = fvu;
      rTransKp1 = rTrans;
    }
    __syncthreads();
  }
#undef SM
}

// ---- pipelined variant: level k-1 (theta, u, v, hFacW, hFacS of the patch) is fetched with cp.async into a two-slot
// ring while level k is computed; same arithmetic as thermo_fast_kernel (bit-identical results).
enum { TR_T = 0, TR_U, TR_V, TR_HW, TR_HS, TR_N };
template <bool CG>
struct ThermoPipeSmemT {
  double raw[2][CG ? TR_HW : TR_N][FT_N];      // CG: no ring slots for hFacW / hFacS ...
  double T[FT_N], xA[FT_N], yA[FT_N], uT[FT_N], vT[FT_N], dyG[FT_N], dxG[FT_N];
  // ... but the column geometry of the patch, from which they are rebuilt per level
  double hLW[CG ? FT_N : 1], hLS[CG ? FT_N : 1];
  int kLW[CG ? FT_N : 1], kLS[CG ? FT_N : 1];
};
typedef ThermoPipeSmemT<false> ThermoPipeSmem;
template <bool CG>
__device__ __forceinline__ void thermo_pipe_prefetch(ThermoPipeSmemT<CG> &sm, int slot, int e, size_t q, const TileGrid &g,
                                                     const double *u, const double *v, const double *theta) {
  __pipeline_memcpy_async(&sm.raw[slot][TR_T][e], theta + q, 8);
  __pipeline_memcpy_async(&sm.raw[slot][TR_U][e], u + q, 8);
  __pipeline_memcpy_async(&sm.raw[slot][TR_V][e], v + q, 8);
  if (!CG) {
    __pipeline_memcpy_async(&sm.raw[slot][CG ? 0 : TR_HW][e], g.hFacW + q, 8);
    __pipeline_memcpy_async(&sm.raw[slot][CG ? 0 : TR_HS][e], g.hFacS + q, 8);
  }
}
#ifndef THP_MINB
#define THP_MINB 3
#endif
template <bool CG>      // CG: hFacW/S, maskC, recip_hFacC from the column geometry (colgeom.cu), not from the 3-D arrays
__global__ void __launch_bounds__(FT_X *FT_Y, THP_MINB)
    thermo_pipe_kernel(TileGrid g, const double *__restrict__ u, const double *__restrict__ v, const double *__restrict__ w,
                       const double *__restrict__ theta, const double *__restrict__ kapT, double *__restrict__ thetaNew,
                       double *__restrict__ gtNm1, GadPar p, double abFac, const double *__restrict__ sfT,
                       const double *__restrict__ gT0, int doAB) {
  // gT0: tendency to start from (GAD_ADVECTION's, with calcAdvection = F); doAB: AdamsBashforthGt
  extern __shared__ __align__(16) unsigned char thermo_pipe_smem[];
  ThermoPipeSmemT<CG> &sm = *reinterpret_cast<ThermoPipeSmemT<CG> *>(thermo_pipe_smem);
  __shared__ VertSmem vs;
  const int tx = threadIdx.x, ty = threadIdx.y, t = ty * FT_X + tx;
  stage_vert(vs, g, t, FT_X * FT_Y);
  __syncthreads();
  const int i0 = 1 + blockIdx.x * FT_X, j0 = 1 + blockIdx.y * FT_Y;
  const int i = i0 + tx, j = j0 + ty;
  const bool active = i <= g.sNx && j <= g.sNy;
  const int c = (ty + 1) * FT_W + (tx + 1);
#define SM(a, di, dj) sm.a[c + (dj)*FT_W + (di)]
  int se[2];
  size_t sg[2];
  bool sv_[2];
#pragma unroll
  for (int r = 0; r < 2; r++) {
    int e = t + r * FT_X * FT_Y;
    sv_[r] = e < FT_N;
    int li = sv_[r] ? e % FT_W : 0, lj = sv_[r] ? e / FT_W : 0;
    int gi = min(i0 - 1 + li, g.sNx + g.OLx), gj = min(j0 - 1 + lj, g.sNy + g.OLy);
    se[r] = e;
    sg[r] = g.s(gi, gj);
  }
  const size_t s = active ? g.s(i, j) : g.s(1, 1);
  const size_t slab = g.slab;
  const int PX = g.PX;
  const double rdxC0 = g.recip_dxC[s], rdxC1 = g.recip_dxC[s + 1], rdyC0 = g.recip_dyC[s], rdyC1 = g.recip_dyC[s + PX];
  const double rA = g.rA[s], r_rA = g.recip_rA[s], cfU = g.cosFacU[active ? j + g.OLy - 1 : 0];
  const double advFac = p.calcAdvection ? 1. : 0.;
  double rAdvFac = p.rkSign * advFac;
  if (p.implicitAdvection) rAdvFac = p.rkSign;
  double fVdn = 0., rTransKp1 = 0.;
  // k-invariant patch metrics and the prefetch of the bottom level into ring slot 0
  int rb = 0;
#pragma unroll
  for (int r = 0; r < 2; r++)
    if (sv_[r]) {
      sm.dyG[se[r]] = g.dyG[sg[r]]; sm.dxG[se[r]] = g.dxG[sg[r]];
      if (CG) {
        sm.kLW[CG ? se[r] : 0] = g.kLowW[sg[r]]; sm.hLW[CG ? se[r] : 0] = g.hLowW[sg[r]];
        sm.kLS[CG ? se[r] : 0] = g.kLowS[sg[r]]; sm.hLS[CG ? se[r] : 0] = g.hLowS[sg[r]];
      }
      thermo_pipe_prefetch<CG>(sm, 0, se[r], sg[r] + slab * (size_t)(g.Nr - 1), g, u, v, theta);
    }
  __pipeline_commit();
  int kLC = 0;
  double rhLC = 0.;
  if (CG) { kLC = g.kLowC[s]; rhLC = g.rhLowC[s]; }
  for (int k = g.Nr; k >= 1; k--) {
    const size_t ko = slab * (size_t)(k - 1);
    const double drFk = vs.drF[k - 1];
    // own-column loads first: one exposed memory latency per level
    const size_t s3 = s + ko;
    const double Tkm1 = k >= 2 ? theta[s3 - slab] : 0.;
    const double mC = CG ? cg_mask(k, kLC) : g.maskC[s3], mCm1 = k >= 2 ? (CG ? cg_mask(k - 1, kLC) : g.maskC[s3 - slab]) : 0.;
    const double wK = w[s3], kapK = kapT[s3], rhC = CG ? cg_hfac(k, kLC, rhLC) : g.recip_hFacC[s3], gtOld = doAB ? gtNm1[s3] : 0.;
    const double gTin = gT0 ? gT0[s3] : 0.;
    __pipeline_wait_prior(0);
    __syncthreads();             // level k has landed in slot rb; everybody is done with the previous level's tiles
#pragma unroll
    for (int r = 0; r < 2; r++)
      if (sv_[r]) {
        const int e = se[r];
        const double hW = CG ? cg_hfac(k, sm.kLW[CG ? e : 0], sm.hLW[CG ? e : 0]) : sm.raw[rb][CG ? 0 : TR_HW][e];
        const double hS = CG ? cg_hfac(k, sm.kLS[CG ? e : 0], sm.hLS[CG ? e : 0]) : sm.raw[rb][CG ? 0 : TR_HS][e];
        const double xA = sm.dyG[e] * drFk * hW, yA = sm.dxG[e] * drFk * hS;
        sm.T[e] = sm.raw[rb][TR_T][e]; sm.xA[e] = xA; sm.yA[e] = yA;
        sm.uT[e] = sm.raw[rb][TR_U][e] * xA; sm.vT[e] = sm.raw[rb][TR_V][e] * yA;
      }
    __syncthreads();
    if (k >= 2) {                // fetch level k-1 into the other slot while this level is computed
#pragma unroll
      for (int r = 0; r < 2; r++)
        if (sv_[r]) thermo_pipe_prefetch<CG>(sm, rb ^ 1, se[r], sg[r] + ko - slab, g, u, v, theta);
    }
    __pipeline_commit();
    if (active) {
      const double T00 = SM(T, 0, 0);
      // X / Y fluxes (GAD_C2_ADV_X/Y, GAD_DIFF_X/Y), west+east and south+north faces
      double fz0 = 0., fz1 = 0., fm0 = 0., fm1 = 0.;
      if (p.calcAdvection) {
        fz0 = fz0 + SM(uT, 0, 0) * (T00 + SM(T, -1, 0)) * 0.5;
        fz1 = fz1 + SM(uT, 1, 0) * (SM(T, 1, 0) + T00) * 0.5;
        fm0 = fm0 + SM(vT, 0, 0) * (T00 + SM(T, 0, -1)) * 0.5;
        fm1 = fm1 + SM(vT, 0, 1) * (SM(T, 0, 1) + T00) * 0.5;
      }
      double dz0 = 0., dz1 = 0., dm0 = 0., dm1 = 0.;
      if (p.diffKh != 0.) {
        dz0 = -p.diffKh * SM(xA, 0, 0) * rdxC0 * (T00 - SM(T, -1, 0)) * cfU;
        dz1 = -p.diffKh * SM(xA, 1, 0) * rdxC1 * (SM(T, 1, 0) - T00) * cfU;
        dm0 = -p.diffKh * SM(yA, 0, 0) * rdyC0 * (T00 - SM(T, 0, -1));
        dm1 = -p.diffKh * SM(yA, 0, 1) * rdyC1 * (SM(T, 0, 1) - T00);
      }
      fz0 = fz0 + dz0; fz1 = fz1 + dz1; fm0 = fm0 + dm0; fm1 = fm1 + dm1;
      // vertical flux at the upper interface of level k (GAD_C2_ADV_R, GAD_DIFF_R)
      double fvu = 0., rTrans = 0.;
      if (k >= 2) {
        const double maskUp = mCm1 * mC;
        rTrans = wK * rA * maskUp;
        if (p.calcAdvection && !p.implicitAdvection) fvu = fvu + mCm1 * rTrans * (T00 + Tkm1) * 0.5;
        double df = 0.;
        if (!p.implicitDiffusion) df = -kapK * maskUp * rA * vs.rdrC[k - 1] * (T00 - Tkm1) * p.rkSign;
        fvu = fvu + df;
      }
      double gT = gTin - rhC * vs.rdrF[k - 1] * r_rA *
                           ((fz1 - fz0) + (fm1 - fm0) + (fVdn - fvu) * p.rkSign -
                            T00 * ((SM(uT, 1, 0) - SM(uT, 0, 0)) * advFac + (SM(vT, 0, 1) - SM(vT, 0, 0)) * advFac +
                                   (rTransKp1 - rTrans) * rAdvFac));
      if (sfT) {   // APPLY_FORCING_T at k = kSurface, inside Adams-Bashforth (tracForcingOutAB = 0)
        double gtForc = 0.;
        if (k == 1) gtForc = gtForc + sfT[g.s(i, j)] * vs.rdrF[0] * rhC;
        gT = gT + gtForc;
      }
      if (doAB) {
        const double ab = abFac * (gT - gtOld);
        gtNm1[s3] = gT;
        gT = gT + ab;
      }
      thetaNew[s3] = T00 + p.deltaT * gT;
      fVdn = fvu;
      rTransKp1 = rTrans;
    }
    rb ^= 1;
  }
#undef SM
}


inline bool thermo_fast_ok(const Geom &g, const GadPar &p) {
  // with calcAdvection = F (multi-dimensional advection done by GAD_ADVECTION) the scheme does not matter
  return ((p.advScheme == ADV_CENTERED_2ND && p.vertAdvScheme == ADV_CENTERED_2ND) || !p.calcAdvection) && p.diffK4 == 0. && !p.useDiffKr4 &&
         g.OLx >= 2 && g.OLy >= 2 && g.Nr < FT_NRMAX && !getenv("MITGCM_B200_GENERIC_STEP");
}

inline bool dyn_fast_ok(const Geom &g, const MomPar &p) {
  return !p.useBiharmonicVisc && p.selectBotDragQuadr == -1 && (p.selectCoriScheme == 0 || p.selectCoriScheme == 2) &&
         !(p.usingSphericalPolarGrid && p.metricTerms) && g.OLx >= 2 && g.OLy >= 2 && g.Nr < FT_NRMAX &&
         !getenv("MITGCM_B200_GENERIC_STEP");
}

}  // namespace mg
