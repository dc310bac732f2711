// vecinv_fast.cuh -- shared-memory staged, cp.async pipelined DYNAMICS with MOM_VECINV for the resident step:
// the structure of dyn_pipe_kernel (step_fast.cuh) with the vector-invariant tendencies of vecinv.cuh.
// A CTA of 32 x 8 columns marches in k.  Per level it receives u, v, hFacW, hFacS, recip_hFacC (level k) and
// w, maskC (level k+1) for its 34 x 10 patch by cp.async into a two-slot ring while the previous level is being
// computed, derives hFacZ, the masked relative vorticity (with the cube's three-cell facet corners), the
// absolute vorticity, KE and the horizontal divergence ONCE per patch cell into shared memory, and every thread
// then forms its tendencies from shared memory through the same templated leaf functions the per-level entry
// point uses (vi_coriolis, vi_{u,v}_coriolis: identical expression order, so results are bit-identical to
// dyn_kernel<1>).  k-invariant metrics sit in registers / shared memory, own-column values at k-1 / k+1 and the
// vertical-shear transports are carried between levels.
//
// Fast path conditions (host: vi_fast_ok; otherwise dyn_kernel<2> runs): no biharmonic viscosity,
// selectBotDragQuadr = -1, selectVortScheme in {0, 1, 2}, selectKEscheme in {-1, 0, 2}, OLx, OLy >= 2.
#pragma once
#include "step_fast.cuh"
#include "vecinv.cuh"

namespace mg {

enum { VR_U = 0, VR_V, VR_HW, VR_HS, VR_RHC, VR_W, VR_MC, VR_N };
struct ViPipeSmem {
  double raw[2][VR_N][FT_N];      // ring: u, v, hFacW, hFacS, recip_hFacC of level k; w, maskC of level k+1
  double hZ[FT_N], z3[FT_N], om[FT_N], KE[FT_N], hD[FT_N], wA[FT_N], mCk[FT_N];
  double dxG[FT_N], dyG[FT_N], dxC[FT_N], dyC[FT_N], rAzI[FT_N], rA[FT_N], rrA[FT_N], fG[FT_N];
  // column geometry of the patch (CG variant, colgeom.cu): the four geometry slots of the ring are then FILLED by
  // compare + select from these instead of being fetched, so everything downstream reads the same shared memory
  double hLW[FT_N], hLS[FT_N], rhLC[FT_N];
  int kLW[FT_N], kLS[FT_N], kLC[FT_N];
};

struct ViRingSrc {
  const ViPipeSmem &sm; int rb, i0, j0;
  __device__ int e(int i, int j) const { return (j - j0 + 1) * FT_W + (i - i0 + 1); }
  __device__ double u(int i, int j) const { return sm.raw[rb][VR_U][e(i, j)]; }
  __device__ double v(int i, int j) const { return sm.raw[rb][VR_V][e(i, j)]; }
  __device__ double hW(int i, int j) const { return sm.raw[rb][VR_HW][e(i, j)]; }
  __device__ double hS(int i, int j) const { return sm.raw[rb][VR_HS][e(i, j)]; }
  __device__ double dxG(int i, int j) const { return sm.dxG[e(i, j)]; }
  __device__ double dyG(int i, int j) const { return sm.dyG[e(i, j)]; }
  __device__ double dxC(int i, int j) const { return sm.dxC[e(i, j)]; }
  __device__ double dyC(int i, int j) const { return sm.dyC[e(i, j)]; }
  __device__ double rAzI(int i, int j) const { return sm.rAzI[e(i, j)]; }
};
struct ViRingAcc {
  const ViPipeSmem &sm; int i0, j0;
  __device__ int e(int i, int j) const { return (j - j0 + 1) * FT_W + (i - i0 + 1); }
  __device__ double hFacZ(int i, int j) const { return sm.hZ[e(i, j)]; }
  __device__ double vort3(int i, int j) const { return sm.z3[e(i, j)]; }
  __device__ double omega3(int i, int j) const { return sm.om[e(i, j)]; }
  __device__ double KE(int i, int j) const { return sm.KE[e(i, j)]; }
  __device__ double hDiv(int i, int j) const { return sm.hD[e(i, j)]; }
};

template <bool CG>
__device__ __forceinline__ void vi_pipe_prefetch(ViPipeSmem &sm, int slot, int e, size_t sg, const TileGrid &g,
                                                 const MomState &st, size_t slab, int k, bool below) {
  const size_t q = sg + slab * (size_t)(k - 1);
  __pipeline_memcpy_async(&sm.raw[slot][VR_U][e], st.u + q, 8);
  __pipeline_memcpy_async(&sm.raw[slot][VR_V][e], st.v + q, 8);
  if (CG) {      // plain stores into the slot nobody reads before the next barrier
    sm.raw[slot][VR_HW][e] = cg_hfac(k, sm.kLW[e], sm.hLW[e]);
    sm.raw[slot][VR_HS][e] = cg_hfac(k, sm.kLS[e], sm.hLS[e]);
    sm.raw[slot][VR_RHC][e] = cg_hfac(k, sm.kLC[e], sm.rhLC[e]);
  } else {
    __pipeline_memcpy_async(&sm.raw[slot][VR_HW][e], g.hFacW + q, 8);
    __pipeline_memcpy_async(&sm.raw[slot][VR_HS][e], g.hFacS + q, 8);
    __pipeline_memcpy_async(&sm.raw[slot][VR_RHC][e], g.recip_hFacC + q, 8);
  }
  if (below) {
    __pipeline_memcpy_async(&sm.raw[slot][VR_W][e], st.w + q + slab, 8);
    if (CG) sm.raw[slot][VR_MC][e] = cg_mask(k + 1, sm.kLC[e]);
    else __pipeline_memcpy_async(&sm.raw[slot][VR_MC][e], g.maskC + q + slab, 8);
  }
}

#ifndef VIP_MINB
#define VIP_MINB 2
#endif
template <bool CG>
__global__ void __launch_bounds__(FT_X *FT_Y, VIP_MINB)
    vi_pipe_kernel(TileGrid g, MomState st, ViPar vp, const double *__restrict__ sfU, const double *__restrict__ sfV,
                   double *__restrict__ gU, double *__restrict__ gV, double *__restrict__ guNm1,
                   double *__restrict__ gvNm1, double deltaTMom, double abFac, int momForcing, int dissInAB,
                   const double *__restrict__ phiHyd) {
  extern __shared__ __align__(16) unsigned char vi_pipe_smem[];
  ViPipeSmem &sm = *reinterpret_cast<ViPipeSmem *>(vi_pipe_smem);
  __shared__ VertSmem vs;
  const MomPar &p = vp.m;
  const int tx = threadIdx.x, ty = threadIdx.y, t = ty * FT_X + tx;
  stage_vert(vs, g, t, FT_X * FT_Y);
  const int i0 = blockIdx.x * FT_X, j0 = blockIdx.y * FT_Y;     // output range 0..sN+1 (dynamics.F:191-192)
  const int i = i0 + tx, j = j0 + ty;
  const bool active = i <= g.sNx + 1 && j <= g.sNy + 1;
  const int c = (ty + 1) * FT_W + (tx + 1);                     // my cell in the staged patch
  // staging map: entries e = t and t + 256 of the 34 x 10 patch, clamped into the halo'd slab
  int se[2], sgi[2], sgj[2];
  size_t sg[2];
  bool sv_[2];
#pragma unroll
  for (int r = 0; r < 2; r++) {
    int e = t + r * FT_X * FT_Y;
    sv_[r] = e < FT_N;
    int li = sv_[r] ? e % FT_W : 0, lj = sv_[r] ? e / FT_W : 0;
    sgi[r] = min(i0 - 1 + li, g.sNx + g.OLx); sgj[r] = min(j0 - 1 + lj, g.sNy + g.OLy);
    se[r] = e;
    sg[r] = g.s(sgi[r], sgj[r]);
  }
  // k-invariant metrics of my column
  const size_t s = active ? g.s(i, j) : g.s(0, 0);
  const int PX = g.PX;
  const double r_rAw = g.recip_rAw[s], r_rAs = g.recip_rAs[s], rAw = g.rAw[s], rAs = g.rAs[s];
  const double rdxC = g.recip_dxC[s], rdyC = g.recip_dyC[s], rdxG = g.recip_dxG[s], rdyG = g.recip_dyG[s];
  const double cfU = g.cosFacU[j + g.OLy - 1 < g.PY ? j + g.OLy - 1 : 0], cfV = g.cosFacV[j + g.OLy - 1 < g.PY ? j + g.OLy - 1 : 0];
  const double sfu = sfU[s], sfv = sfV[s];
  const double gpx = (phiHyd && i >= 1) ? rdxC : 0., gpy = (phiHyd && j >= 1) ? rdyC : 0.;
  double dxV00 = 0., dxV01 = 0., rdyU00 = 0., rdyU01 = 0., dyU00 = 0., dyU10 = 0., rdxV00 = 0., rdxV10 = 0.;
  if (p.momViscosity && p.no_slip_sides) {
    dxV00 = g.dxV[s]; dxV01 = g.dxV[s + PX]; rdyU00 = g.recip_dyU[s]; rdyU01 = g.recip_dyU[s + PX];
    dyU00 = g.dyU[s]; dyU10 = g.dyU[s + 1]; rdxV00 = g.recip_dxV[s]; rdxV10 = g.recip_dxV[s + 1];
  }
  const bool rAdvAreaWeight = true;              // selectKEscheme 1, 3 are not on this path
  (void)rAdvAreaWeight;

  // own-column values carried between levels
  const size_t slab = g.slab;
  double uKm1 = 0., vKm1 = 0.;
  int kLoW = 0, kLoS = 0;
  double rhLoW = 0., rhLoS = 0.;
  if (CG) { kLoW = g.kLowW[s]; kLoS = g.kLowS[s]; rhLoW = g.rhLowW[s]; rhLoS = g.rhLowS[s]; }
  double uK = st.u[s], vK = st.v[s], mWk = CG ? cg_mask(1, kLoW) : g.maskW[s], mSk = CG ? cg_mask(1, kLoS) : g.maskS[s];
  // k-invariant patch metrics, the interface of level 1 (w*rA and maskC of level 1: MOM_VI_{U,V}_VERTSHEAR at
  // k = 1 reads them with mask_Km1 = 0), and the asynchronous prefetch of level 1 into ring slot 1
  int rb = 1;
#pragma unroll
  for (int r = 0; r < 2; r++)
    if (sv_[r]) {
      const int e = se[r];
      sm.dxG[e] = g.dxG[sg[r]]; sm.dyG[e] = g.dyG[sg[r]]; sm.dxC[e] = g.dxC[sg[r]]; sm.dyC[e] = g.dyC[sg[r]];
      sm.rAzI[e] = g.recip_rAz[sg[r]]; sm.rA[e] = g.rA[sg[r]]; sm.rrA[e] = g.recip_rA[sg[r]]; sm.fG[e] = g.fCoriG[sg[r]];
      sm.wA[e] = st.w[sg[r]] * g.rA[sg[r]];
      if (CG) {
        sm.kLW[e] = g.kLowW[sg[r]]; sm.hLW[e] = g.hLowW[sg[r]]; sm.kLS[e] = g.kLowS[sg[r]]; sm.hLS[e] = g.hLowS[sg[r]];
        sm.kLC[e] = g.kLowC[sg[r]]; sm.rhLC[e] = g.rhLowC[sg[r]];
        sm.raw[0][VR_MC][e] = cg_mask(1, sm.kLC[e]);
      } else
        sm.raw[0][VR_MC][e] = g.maskC[sg[r]];
      vi_pipe_prefetch<CG>(sm, 1, e, sg[r], g, st, slab, 1, 1 + 1 <= g.Nr);
    }
  __pipeline_commit();
  __syncthreads();
  // vertical-shear transport at the upper interface of level 1 (multiplied by mask_Km1 = 0, kept for the sign of zero)
  double wBmU = 0.5 * (sm.wA[c] * sm.raw[0][VR_MC][c] + sm.wA[c - 1] * sm.raw[0][VR_MC][c - 1]) * 0. * r_rAw;
  double wBmV = 0.5 * (sm.wA[c] * sm.raw[0][VR_MC][c] + sm.wA[c - FT_W] * sm.raw[0][VR_MC][c - FT_W]) * 0. * r_rAs;
  double ukm = 0., vkm = 0.;                     // viscous vertical fluxes fVerU/V at the upper interface
  __syncthreads();
  for (int k = 1; k <= g.Nr; k++) {
    const size_t ko = slab * (size_t)(k - 1);
    const double rdrF = vs.rdrF[k - 1];
    const bool below = k + 1 <= g.Nr;
    // own column: issue every global load of this level up front
    const size_t s3 = s + ko;
    double uKp1 = 0., vKp1 = 0., mWkp1 = 0., mSkp1 = 0.;
    const double kapUkp1 = st.kapU[s3 + slab], kapVkp1 = st.kapV[s3 + slab];
    if (below) {
      uKp1 = st.u[s3 + slab]; vKp1 = st.v[s3 + slab];
      mWkp1 = CG ? cg_mask(k + 1, kLoW) : g.maskW[s3 + slab]; mSkp1 = CG ? cg_mask(k + 1, kLoS) : g.maskS[s3 + slab];
    }
    const double rhW = CG ? cg_hfac(k, kLoW, rhLoW) : g.recip_hFacW[s3], rhS = CG ? cg_hfac(k, kLoS, rhLoS) : g.recip_hFacS[s3];
    const double guOld = guNm1[s3], gvOld = gvNm1[s3];
    double dpx = 0., dpy = 0.;
    if (phiHyd) {
      const double ph = phiHyd[s3];
      dpx = gpx * 1. * (ph - phiHyd[s3 - 1]) * 1.;
      dpy = gpy * 1. * (ph - phiHyd[s3 - PX]) * 1.;
    }
    // ---- level k has been prefetched into ring slot rb (cp.async); derive what the tendencies share ----
    __pipeline_wait_prior(0);
    __syncthreads();
    const ViRingSrc f{sm, rb, i0, j0};
#pragma unroll
    for (int r = 0; r < 2; r++)
      if (sv_[r]) {
        const int e = se[r], li = e % FT_W, lj = e / FT_W, gi = sgi[r], gj = sgj[r];
        sm.mCk[e] = sm.raw[rb ^ 1][VR_MC][e];        // maskC(k), fetched with level k-1
        double hz = 0., z = 0.;
        // vorticity-point quantities need the west / south neighbours: patch cells li, lj >= 1 (all that are read)
        if (li >= 1 && lj >= 1 && gi == i0 - 1 + li && gj == j0 - 1 + lj) {
          hz = vi_hfacz(g, f, gi, gj);
          if (hz != 0.) z = vi_relvort3(g, f, vp.csCorners, vp.myFace, gi, gj);
        }
        sm.hZ[e] = hz; sm.z3[e] = z;
        sm.om[e] = sm.fG[e] * (vp.useCoriolis ? 1. : 0.) + z * (p.momAdvection ? 1. : 0.);
        // cell-centred quantities need the east / north neighbours: li <= 32, lj <= 8 (all that are read)
        double ke = 0., hd = 0.;
        if (li <= FT_W - 2 && lj <= FT_H - 2 && gi == i0 - 1 + li && gj == j0 - 1 + lj) {
          const double u0 = f.u(gi, gj), u1 = f.u(gi + 1, gj), v0 = f.v(gi, gj), v1 = f.v(gi, gj + 1);
          if (gi <= g.sNx + g.OLx - 1 && gj <= g.sNy + g.OLy - 1) {
            if (vp.selectKEscheme == -1) ke = 0.125 * ((u0 + u1) * (u0 + u1) + (v0 + v1) * (v0 + v1));
            else if (vp.selectKEscheme == 0) ke = 0.25 * ((u0 * u0 + u1 * u1) + (v0 * v0 + v1 * v1));
            else ke = 0.25 * ((u0 * u0 * f.hW(gi, gj) + u1 * u1 * f.hW(gi + 1, gj)) + (v0 * v0 * f.hS(gi, gj) + v1 * v1 * f.hS(gi, gj + 1))) *
                      sm.raw[rb][VR_RHC][e];
            if (p.momViscosity)
              hd = ((u1 * f.dyG(gi + 1, gj) * f.hW(gi + 1, gj) - u0 * f.dyG(gi, gj) * f.hW(gi, gj)) +
                    (v1 * f.dxG(gi, gj + 1) * f.hS(gi, gj + 1) - v0 * f.dxG(gi, gj) * f.hS(gi, gj))) *
                   sm.rrA[e] * sm.raw[rb][VR_RHC][e];
          }
        }
        sm.KE[e] = ke; sm.hD[e] = hd;
        if (below) sm.wA[e] = sm.raw[rb][VR_W][e] * sm.rA[e];     // w(k+1)*rA; at k = Nr it keeps w(Nr)*rA
      }
    __syncthreads();
    double wBpU = 0., wBpV = 0., wBmUn = 0., wBmVn = 0.;
    if (below) {      // prefetch level k+1 into the other slot while this level is computed
#pragma unroll
      for (int r = 0; r < 2; r++)
        if (sv_[r]) vi_pipe_prefetch<CG>(sm, rb ^ 1, se[r], sg[r], g, st, slab, k + 1, k + 2 <= g.Nr);
    }
    __pipeline_commit();
    if (active) {
      const ViRingAcc a{sm, i0, j0};
      const double mask_Kp1 = (k == g.Nr) ? 0. : 1., mask_Km1 = (k == 1) ? 0. : 1.;
      // MOM_VI_{U,V}_VERTSHEAR transports (mom_vi_u_vertshear.F:63-77): interface k+1 for this level, and the
      // same interface seen from level k+1 (with maskC(k)) carried to the next iteration.  At k = Nr the
      // reference reads interface Nr with mask_Kp1 = 0: wA still holds w(Nr)*rA then.
      wBpU = 0.5 * (sm.wA[c] + sm.wA[c - 1]) * mask_Kp1 * r_rAw;
      wBpV = 0.5 * (sm.wA[c] + sm.wA[c - FT_W]) * mask_Kp1 * r_rAs;
      if (below) {
        wBmUn = 0.5 * (sm.wA[c] * sm.mCk[c] + sm.wA[c - 1] * sm.mCk[c - 1]) * 1. * r_rAw;
        wBmVn = 0.5 * (sm.wA[c] * sm.mCk[c] + sm.wA[c - FT_W] * sm.mCk[c - FT_W]) * 1. * r_rAs;
      }
      double uD = 0., vD = 0., ukp = 0., vkp = 0.;
      if (p.momViscosity) {
        if (vp.harmonic) {      // MOM_VI_HDISSIP (mom_vi_hdissip.F:60-123), constant coefficients
          const double Dim = a.hDiv(i, j - 1), Dij = a.hDiv(i, j), Dmj = a.hDiv(i - 1, j);
          const double Zip = a.hFacZ(i, j + 1) * a.vort3(i, j + 1), Zij = a.hFacZ(i, j) * a.vort3(i, j),
                       Zpj = a.hFacZ(i + 1, j) * a.vort3(i + 1, j);
          const double uD2 = p.viscAhD * cfU * (Dij - Dmj) * rdxC - p.viscAhZ * rhW * (Zip - Zij) * rdyG;
          const double vD2 = p.viscAhZ * rhS * cfV * (Zpj - Zij) * rdxG + p.viscAhD * (Dij - Dim) * rdyC;
          uD = uD2 * mWk;
          vD = vD2 * mSk;
        }
        if (!p.implicitViscosity) {     // MOM_U_RVISCFLUX at interface k+1 (mom_u_rviscflux.F), mom_vecinv.F:432-450
          double rv = 0.;
          if (below) rv = -kapUkp1 * rAw * (uKp1 - uK) * p.rkSign * vs.rdrC[k] * mWkp1 * mWk;
          ukp = p.vfFacMom * 1. * rv;
          uD = uD - rhW * rdrF * r_rAw * (ukp - ukm) * p.rkSign;
        }
        if (p.no_slip_sides) {          // MOM_U_SIDEDRAG
          const double hWc = f.hW(i, j), t_ = p.viscAhZ * uK - p.viscA4Z * 0.;
          uD = uD + (-rhW * rdrF * r_rAw * ((hWc - a.hFacZ(i, j)) * dxV00 * rdyU00 * t_ + (hWc - a.hFacZ(i, j + 1)) * dxV01 * rdyU01 * t_) *
                     vs.drF[k - 1] * p.sideDragFactor);
        }
        const double viscFac = p.no_slip_bottom ? 2. : 0.;
        const double recDrC = (k == g.Nr) ? rdrF : vs.rdrC[k];
        if (p.bottomDragTerms) {        // MOM_U_BOTDRAG_COEFF with selectBotDragQuadr = -1
          double cu = p.bottomDragLinear * 1.;
          if (p.no_slip_bottom && p.bottomVisc_pCell) cu = cu + kapUkp1 * recDrC * viscFac * rhW;
          else if (p.no_slip_bottom) cu = cu + kapUkp1 * recDrC * viscFac;
          if (k == g.Nr) cu = cu * mWk;
          else cu = cu * mWk * (1. - mWkp1);
          uD = uD + (-cu * uK * rhW * rdrF);
        }
        if (!p.implicitViscosity) {
          double rv = 0.;
          if (below) rv = -kapVkp1 * rAs * (vKp1 - vK) * p.rkSign * vs.rdrC[k] * mSkp1 * mSk;
          vkp = p.vfFacMom * 1. * rv;
          vD = vD - rhS * rdrF * r_rAs * (vkp - vkm) * p.rkSign;
        }
        if (p.no_slip_sides) {          // MOM_V_SIDEDRAG
          const double hSc = f.hS(i, j), t_ = p.viscAhZ * vK * cfV - p.viscA4Z * 0. * cfV;
          vD = vD + (-rhS * rdrF * r_rAs * ((hSc - a.hFacZ(i, j)) * dyU00 * rdxV00 * t_ + (hSc - a.hFacZ(i + 1, j)) * dyU10 * rdxV10 * t_) *
                     vs.drF[k - 1] * p.sideDragFactor);
        }
        if (p.bottomDragTerms) {
          double cv = p.bottomDragLinear * 1.;
          if (p.no_slip_bottom && p.bottomVisc_pCell) cv = cv + kapVkp1 * recDrC * viscFac * rhS;
          else if (p.no_slip_bottom) cv = cv + kapVkp1 * recDrC * viscFac;
          if (k == g.Nr) cv = cv * mSk;
          else cv = cv * mSk * (1. - mSkp1);
          vD = vD + (-cv * vK * rhS * rdrF);
        }
      }
      // ---- Coriolis and advection (mom_vecinv.F:672-884), order of vi_cell
      double tU = 0., tV = 0.;
      if (vp.useCoriolis && !(p.useCDscheme || (vp.useAbsVorticity && p.momAdvection))) {
        if (vp.useAbsVorticity) {
          tU = vi_u_coriolis(g, f, vp, a, true, k, i, j);
          tV = vi_v_coriolis(g, f, vp, a, true, k, i, j);
        } else {
          vi_coriolis(g, f, p.selectCoriScheme, k, i, j, tU, tV);
        }
      }
      if (p.momAdvection) {
        tU = tU + vi_u_coriolis(g, f, vp, a, vp.useAbsVorticity != 0, k, i, j);
        tV = tV + vi_v_coriolis(g, f, vp, a, vp.useAbsVorticity != 0, k, i, j);
        {   // MOM_VI_U_VERTSHEAR / MOM_VI_V_VERTSHEAR (mom_vi_{u,v}_vertshear.F:90-131)
          const double uZm = (uK - mask_Km1 * (k == 1 ? uK : uKm1)) * p.rkSign;
          const double uZp = (mask_Kp1 * (below ? uKp1 : uK) - uK) * p.rkSign;
          const double vZm = (vK - mask_Km1 * (k == 1 ? vK : vKm1)) * p.rkSign;
          const double vZp = (mask_Kp1 * (below ? vKp1 : vK) - vK) * p.rkSign;
          double shU, shV;
          if (vp.upwindShear) {
            shU = -0.5 * ((wBpU * uZp + wBmU * uZm) + (fabs(wBpU) * uZp - fabs(wBmU) * uZm)) * rhW * rdrF;
            shV = -0.5 * ((wBpV * vZp + wBmV * vZm) + (fabs(wBpV) * vZp - fabs(wBmV) * vZm)) * rhS * rdrF;
          } else {
            shU = -0.5 * (wBpU * uZp + wBmU * uZm) * rhW * rdrF;
            shV = -0.5 * (wBpV * vZp + wBmV * vZm) * rhS * rdrF;
          }
          tU = tU + shU;
          tV = tV + shV;
        }
        const double ke = a.KE(i, j);
        tU = tU + (-rdxC * (ke - a.KE(i - 1, j)) * mWk);
        tV = tV + (-rdyC * (ke - a.KE(i, j - 1)) * mSk);
      }
      double gu = tU * mWk, gv = tV * mSk;
      // ---- TIMESTEP (timestep.F:95-385), as in dyn_kernel ----
      gu = gu - 1. * dpx; gv = gv - 1. * dpy;
      if (p.momViscosity && dissInAB) { gu = gu + uD; gv = gv + vD; }
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
      if (p.momViscosity && !dissInAB) { gu = gu + uD; gv = gv + vD; }
      gU[s3] = uK + deltaTMom * (gu + 0.) * mWk;
      gV[s3] = vK + deltaTMom * (gv + 0.) * mSk;
      ukm = ukp; vkm = vkp;
      wBmU = wBmUn; wBmV = wBmVn;
    }
    uKm1 = uK; uK = uKp1; vKm1 = vK; vK = vKp1; mWk = mWkp1; mSk = mSkp1;
    rb ^= 1;
  }
}

inline bool vi_fast_ok(const Geom &g, const ViPar &p) {
  return !p.m.useBiharmonicVisc && p.m.selectBotDragQuadr == -1 && p.selectVortScheme >= 0 && p.selectVortScheme <= 2 &&
         !p.highOrderVorticity && !p.upwindVorticity &&
         (p.selectKEscheme == -1 || p.selectKEscheme == 0 || p.selectKEscheme == 2) && g.OLx >= 2 && g.OLy >= 2 &&
         g.Nr < FT_NRMAX && !getenv("MITGCM_B200_GENERIC_STEP") && !getenv("MITGCM_B200_VI_NOPIPE");
}

}  // namespace mg
