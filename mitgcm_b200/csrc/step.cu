// step.cu -- resident model step: the glue either side of the hot path fused into the kernels
// (SURVEY.md section 8(f) rank 1), in the order of model/src/forward_step.F (non-staggered).
//   thermo_kernel : TEMP_INTEGRATE for k = Nr..1 per column: CALC_ADV_FLOW (calc_adv_flow.F) +
//                   GAD_CALC_RHS (gad.cuh) + ADAMS_BASHFORTH2 (adams_bashforth2.F:84-86) +
//                   TIMESTEP_TRACER (timestep_tracer.F) + CYCLE_TRACER, vertical flux kept in a register
//   dyn_kernel    : DYNAMICS for k = 1..Nr per column on 0..sN+1: MOM_FLUXFORM (mom.cuh) + TIMESTEP
//                   (timestep.F:95-385: dissipation and surface stress inside AB, AB2, u* = u + dt G)
//   rhs_kernel    : SOLVE_FOR_PRESSURE right-hand side (solve_for_pressure.F:120-238, calc_div_ghat.F:64-167)
//   corr_kernel   : MOMENTUM_CORRECTION_STEP (calc_grad_phi_surf.F, correction_step.F:152-231) +
//                   INTEGRATE_FOR_W (integrate_for_w.F), k = Nr..1 per column
//   exch_kernel   : EXCH_XYZ_RL with corners on one periodic process (exch1_rx.template:170-201)
// Assumptions of this driver (checked): linear free surface, no CD scheme, z coordinates, linear EOS,
// surface stress + SST-relaxation forcing; implicSurfPress / implicDiv2DFlow < 1 go through the generic
// dynamics kernel; tile graphs (pkg/exch2) on one GPU.
#include <cstdio>
#include <cstdlib>
#include "step_fast.cuh"
#include "dyn_tma.cuh"
#include "vecinv.cuh"
#include "vecinv_fast.cuh"
#include "phys.cuh"

#define MG_DO_PRAGMA(x) _Pragma(#x)
#define UNROLL_N(n) MG_DO_PRAGMA(unroll n)

namespace mg {

bool cg2d_run(bool sr, double *cg2d_b, double *cg2d_x, double *firstResidual, double *minResidualSq,
              double *lastResidual, int *numIters, int *nIterMin);
bool make_mom_par(MomPar &p);

// ---- halo exchange ---------------------------------------------------------------------------
// One pass: every halo cell (corners included) reads the interior cell it mirrors under the
// periodic nSx x nSy tiling.  Equivalent to the X-then-Y sequence of EXCH1_RX because all sources
// are interior cells.
__global__ void exch_kernel(double *f, int nz, int sNx, int sNy, int OLx, int OLy, int nSx, int nSy) {
  const int PX = sNx + 2 * OLx, PY = sNy + 2 * OLy;
  const size_t slab = (size_t)PX * PY;
  const int nHalo = PX * PY - sNx * sNy;
  const size_t total = (size_t)nHalo * nz * nSx * nSy;
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    int h = (int)(t % nHalo);
    size_t r = t / nHalo;
    int k = (int)(r % nz);
    int tile = (int)(r / nz);
    // enumerate halo cells: south rows, north rows, then west/east columns of the middle rows
    int ii, jj;
    if (h < OLy * PX) { jj = h / PX; ii = h % PX; }
    else if (h < 2 * OLy * PX) { int q = h - OLy * PX; jj = OLy + sNy + q / PX; ii = q % PX; }
    else {
      int q = h - 2 * OLy * PX;
      jj = OLy + q / (2 * OLx);
      int c = q % (2 * OLx);
      ii = c < OLx ? c : sNx + c;
    }
    int bi = tile % nSx, bj = tile / nSx;
    // global interior coordinates of the mirrored cell
    int gi = bi * sNx + (ii - OLx), gj = bj * sNy + (jj - OLy);
    const int Nx = sNx * nSx, Ny = sNy * nSy;
    gi = ((gi % Nx) + Nx) % Nx;
    gj = ((gj % Ny) + Ny) % Ny;
    int sbi = gi / sNx, sbj = gj / sNy;
    int si = gi - sbi * sNx + OLx, sj = gj - sbj * sNy + OLy;
    size_t dst = (size_t)ii + (size_t)PX * jj + slab * ((size_t)k + (size_t)nz * tile);
    size_t src = (size_t)si + (size_t)PX * sj + slab * ((size_t)k + (size_t)nz * (sbi + nSx * sbj));
    f[dst] = f[src];
  }
}

bool exch_field(double *f, int nz) {
  Ctx &c = ctx();
  const Geom &g = c.g;
  if (exch2_active()) return exch2_field(f, nz);
  if (g.nPx != 1 || g.nPy != 1) {      // across ranks: peer pushes (the mirror must live in the peer arena)
    if (halo_connected())
      for (auto &kv : c.fields)
        if (kv.second == f) { const int id = kv.first; return halo_exchange(&id, 1); }
    return fail(60, "exch: multi-process exchange needs mitgcm_b200_comm_connect_ and a mirror in the peer arena");
  }
  const size_t total = (size_t)(g.PX * g.PY - g.sNx * g.sNy) * nz * g.nTiles;
  int blocks = (int)std::min<size_t>((total + 255) / 256, (size_t)c.numSMs * 16);
  c.launches++;
  exch_kernel<<<std::max(blocks, 1), 256, 0, c.stream>>>(f, nz, g.sNx, g.sNy, g.OLx, g.OLy, g.nSx, g.nSy);
  MG_CUDA(cudaGetLastError());
  return true;
}

// ---- thermodynamics ---------------------------------------------------------------------------
// CALC_ADV_FLOW fused: level-k transports derived from the resident 3-D state.
struct FusedAcc {
  TileGrid g;
  const double *u, *v, *w, *T_, *kapT;
  int k;
  __device__ __forceinline__ double T(int i, int j, int kk) const { return T_[g.s3(i, j, kk)]; }
  __device__ __forceinline__ double TA(int i, int j, int kk) const { return T_[g.s3(i, j, kk)]; }
  __device__ __forceinline__ double xA(int i, int j) const { return g.dyG[g.s(i, j)] * g.drF[k - 1] * g.hFacW[g.s3(i, j, k)]; }
  __device__ __forceinline__ double yA(int i, int j) const { return g.dxG[g.s(i, j)] * g.drF[k - 1] * g.hFacS[g.s3(i, j, k)]; }
  __device__ __forceinline__ double uFld(int i, int j) const { return u[g.s3(i, j, k)]; }
  __device__ __forceinline__ double vFld(int i, int j) const { return v[g.s3(i, j, k)]; }
  __device__ __forceinline__ double wFld(int i, int j) const { return w[g.s3(i, j, k)]; }
  __device__ __forceinline__ double uTrans(int i, int j) const { return u[g.s3(i, j, k)] * xA(i, j); }
  __device__ __forceinline__ double vTrans(int i, int j) const { return v[g.s3(i, j, k)] * yA(i, j); }
  __device__ __forceinline__ double maskUpAt(int i, int j, int kk) const {
    return kk == 1 ? 0. : g.maskC[g.s3(i, j, kk - 1)] * g.maskC[g.s3(i, j, kk)];
  }
  __device__ __forceinline__ double rTransAt(int i, int j, int kk) const {
    return kk == 1 ? 0. : w[g.s3(i, j, kk)] * g.rA[g.s(i, j)] * maskUpAt(i, j, kk);
  }
  __device__ __forceinline__ double maskUp(int i, int j) const { return maskUpAt(i, j, k); }
  __device__ __forceinline__ double rTrans(int i, int j) const { return rTransAt(i, j, k); }
  __device__ __forceinline__ double rTransKp1(int i, int j) const { return k == g.Nr ? 0. : rTransAt(i, j, k + 1); }
  __device__ __forceinline__ double KappaR(int i, int j) const { return kapT[g.s3(i, j, k)]; }
};

#ifndef THERMO_MINB
#define THERMO_MINB 4
#endif
__global__ void __launch_bounds__(128, THERMO_MINB) thermo_kernel(TileGrid g, const double *u, const double *v, const double *w,
                                                     const double *theta, const double *kapT, double *thetaNew,
                                                     double *gtNm1, GadPar p0, double abFac, const double *sfT,
                                                     const double *gT0, int doAB) {
  // gT0: tendency to start from (the multi-dimensional advection of GAD_ADVECTION, temp_integrate.F:276-290);
  // doAB: AdamsBashforthGt (gad_init_fixed.F: only the centred / 3rd-order upwind schemes 2, 3, 4)
  const int i = 1 + blockIdx.x * 32 + threadIdx.x;
  const int j = 1 + blockIdx.y * 4 + threadIdx.y;
  if (i > g.sNx || j > g.sNy) return;
  FusedAcc a{g, u, v, w, theta, kapT, 0};
  GadPar p = p0;
  double fVdn = 0.;   // fVer(kDown): flux left by level k+1 (zero below the bottom level, temp_integrate.F:211-215)
  for (int k = g.Nr; k >= 1; k--) {
    a.k = k;
    p.k = k;
    const double fz0 = gad_fzon(g, a, p, i, j), fz1 = gad_fzon(g, a, p, i + 1, j);
    const double fm0 = gad_fmer(g, a, p, i, j), fm1 = gad_fmer(g, a, p, i, j + 1);
    const double fvu = gad_fver(g, a, p, i, j);
    const size_t s3 = g.s3(i, j, k);
    double gT = gad_tendency(g, a, p, i, j, gT0 ? gT0[s3] : 0., fz0, fz1, fm0, fm1, fvu, fVdn);
    if (sfT) {    // APPLY_FORCING_T (apply_forcing.F, k = kSurface), tracForcingOutAB = 0 (temp_integrate.F:392-398)
      double gtForc = 0.;
      if (k == 1) gtForc = gtForc + sfT[g.s(i, j)] * g.recip_drF[0] * g.recip_hFacC[s3];
      gT = gT + gtForc;
    }
    if (doAB) {
      const double ab = abFac * (gT - gtNm1[s3]);          // ADAMS_BASHFORTH2
      gtNm1[s3] = gT;
      gT = gT + ab;
    }
    thetaNew[s3] = theta[s3] + p.deltaT * gT;              // TIMESTEP_TRACER + CYCLE_TRACER
    fVdn = fvu;
  }
}

// ---- dynamics ---------------------------------------------------------------------------------
#ifndef DYN_MINB
#define DYN_MINB 4
#endif
// VI = 1: MOM_VECINV (vecinv.cuh) instead of MOM_FLUXFORM, vorticity / KE / divergence re-evaluated where
// they are read; VI = 2: the CTA (32 x 8) evaluates them once per level for its patch into shared memory.
// The values carried down the column are then the viscous vertical fluxes fVerU/V (dynamics.F:527-533).
template <int VI>
__global__ void __launch_bounds__(VI == 2 ? 256 : 128, VI == 2 ? 2 : DYN_MINB) dyn_kernel(TileGrid g, MomState st, MomPar p, ViPar vp, const double *sfU, const double *sfV,
                                                  double *gU, double *gV, double *guNm1, double *gvNm1,
                                                  double deltaTMom, double abFac, int momForcing, int dissInAB,
                                                  const double *phiHyd, const double *etaN, const double *Bo_surf,
                                                  double psFacTS) {
  __shared__ double vtRaw[VI == 2 ? (sizeof(ViTile) + sizeof(ViPatch)) / sizeof(double) : 1];
  ViTile &vt = *reinterpret_cast<ViTile *>(vtRaw);
  ViPatch &vpt = *reinterpret_cast<ViPatch *>(vtRaw + (VI == 2 ? sizeof(ViTile) / sizeof(double) : 0));
  const int i0 = blockIdx.x * 32, j0 = blockIdx.y * blockDim.y;
  int i = i0 + threadIdx.x;        // 0 .. sNx+1 (dynamics.F:191-192)
  int j = j0 + threadIdx.y;
  const bool active = i <= g.sNx + 1 && j <= g.sNy + 1;
  if (VI != 2 && !active) return;
  if (!active) { i = 0; j = 0; }
  if (VI == 2) vi_fill_patch_metrics(vpt, g, i0, j0, threadIdx.y * 32 + threadIdx.x);
  double ukm = 0., vkm = 0.;
  if (!VI && p.momAdvection && !p.rigidLid) { ukm = mom_adv_wu(g, st, p, 1, i, j); vkm = mom_adv_wv(g, st, p, 1, i, j); }
  const size_t s = g.s(i, j);
  // CALC_GRAD_PHI_SURF (calc_grad_phi_surf.F) when implicSurfPress != 1 (dynamics.F:249-255); the explicit
  // part of the surface pressure gradient enters TIMESTEP as gUdPx = -psFac*phiSurfX, psFac = 1 - implicSurfPress
  double gUdPx = 0., gVdPy = 0.;
  if (etaN) {
    const double phiSurfX = g.recip_dxC[s] * (Bo_surf[s] * etaN[s] - Bo_surf[s - 1] * etaN[s - 1]);
    const double phiSurfY = g.recip_dyC[s] * (Bo_surf[s] * etaN[s] - Bo_surf[s - g.PX] * etaN[s - g.PX]);
    gUdPx = -psFacTS * phiSurfX;
    gVdPy = -psFacTS * phiSurfY;
  }
  for (int k = 1; k <= g.Nr; k++) {
    double ukp = 0., vkp = 0.;
    MomOut o;
    if (VI == 2) {
      __syncthreads();
      vi_fill_patch(vpt, g, st, k, i0, j0, threadIdx.y * 32 + threadIdx.x);
      __syncthreads();
      vi_fill_tile(vt, vpt, g, st, vp, k, i0, j0, threadIdx.y * 32 + threadIdx.x);
      __syncthreads();
      if (!active) continue;
    }
    if (VI) {
      const ViGlobalSrc fg{g, st.u + g.slab * (size_t)(k - 1), st.v + g.slab * (size_t)(k - 1), k};
      const ViOut vo = VI == 2 ? vi_cell(g, st, ViPatchSrc{vpt, i0, j0}, vp, ViTileAcc{vt, i0, j0}, k, i, j, true, ukm, vkm)
                               : vi_cell(g, st, fg, vp, ViFusedAcc<ViGlobalSrc>{g, fg, vp, k}, k, i, j, true, ukm, vkm);
      o.gU = vo.gU; o.gV = vo.gV; o.guDiss = vo.guDiss; o.gvDiss = vo.gvDiss;
      ukp = vo.fVerUkp; vkp = vo.fVerVkp;
    } else {
      if (p.momAdvection) { ukp = mom_adv_wu(g, st, p, k + 1, i, j); vkp = mom_adv_wv(g, st, p, k + 1, i, j); }
      o = mom_cell(g, st, p, k, i, j, ukm, ukp, vkm, vkp);
    }
    const size_t s3 = g.s3(i, j, k);
    double gu = o.gU, gv = o.gV;
    // timestep.F:120-121: gU - phFac*dPhiHydX with CALC_GRAD_PHI_HYD (calc_grad_phi_hyd.F:150-165) on
    // i = iMin+1..iMax, j = jMin+1..jMax; zero when the buoyancy is decoupled (phiHyd == nullptr)
    double dpx = 0., dpy = 0.;
    if (phiHyd) {
      if (i >= 1) dpx = g.recip_dxC[s] * 1. * (phiHyd[s3] - phiHyd[s3 - 1]) * 1.;
      if (j >= 1) dpy = g.recip_dyC[s] * 1. * (phiHyd[s3] - phiHyd[s3 - g.PX]) * 1.;
    }
    gu = gu - 1. * dpx; gv = gv - 1. * dpy;
    if (p.momViscosity && dissInAB) { gu = gu + o.guDiss; gv = gv + o.gvDiss; }
    if (momForcing) {       // apply_forcing.F:142-148: surface stress in the top level
      double ge = 0., he = 0.;
      if (k == 1) {
        if (i >= 1 && i <= g.sNx + 1) ge = 0. + sfU[s] * g.recip_drF[0] * g.recip_hFacW[s3];
        if (j >= 1 && j <= g.sNy + 1) he = 0. + sfV[s] * g.recip_drF[0] * g.recip_hFacS[s3];
      }
      gu = gu + ge; gv = gv + he;
    }
    double ab = abFac * (gu - guNm1[s3]);                  // ADAMS_BASHFORTH2
    guNm1[s3] = gu;
    gu = gu + ab;
    ab = abFac * (gv - gvNm1[s3]);
    gvNm1[s3] = gv;
    gv = gv + ab;
    if (p.momViscosity && !dissInAB) { gu = gu + o.guDiss; gv = gv + o.gvDiss; }
    gU[s3] = st.u[s3] + deltaTMom * (gu + gUdPx) * g.maskW[s3];   // timestep.F:375-384
    gV[s3] = st.v[s3] + deltaTMom * (gv + gVdPy) * g.maskS[s3];
    ukm = ukp; vkm = vkp;
  }
}

// ---- surface pressure right-hand side -----------------------------------------------------------
// Tuned on B200 at 2048x2048x50 (build variants, scripts/phase_times.sh): 16 CTAs/SM (32 registers,
// full occupancy) and 10 levels of loads in flight: 2.20 -> 1.71 ms.  The kernel is latency-bound
// (long-scoreboard stalls), so resident warps matter more than registers.
#ifndef RHS_MINB
#define RHS_MINB 16
#endif
#ifndef RHS_UNROLL
#define RHS_UNROLL 10
#endif
// the column-geometry variants have fewer loads in flight per level: tuned separately (2048^2 x 50: rhs 1.33 -> 1.23 ms at
// 12 CTAs/SM with 5 levels unrolled, corr 2.02 -> 1.88 ms at 4 CTAs/SM)
#ifndef RHS_MINB_CG
#define RHS_MINB_CG 12
#endif
#ifndef RHS_UNROLL_CG
#define RHS_UNROLL_CG 5
#endif
#ifndef CORR_MINB_CG
#define CORR_MINB_CG 4
#endif
template <bool CG>      // CG: open-water fractions from the column geometry (colgeom.cu) instead of the 3-D arrays
__global__ void __launch_bounds__(128, CG ? RHS_MINB_CG : RHS_MINB) rhs_kernel(TileGrid g, const double *__restrict__ gU, const double *__restrict__ gV,
                                                  const double *__restrict__ etaN, const double *__restrict__ etaFS,
                                                  const double *__restrict__ Bo_surf,
                                                  double *__restrict__ cg2d_b, double *__restrict__ cg2d_x,
                                                  double deltaTMom, double deltaTFreeSurf, double freeSurfFac,
                                                  double implicDiv2DFlow, const double *__restrict__ uVel,
                                                  const double *__restrict__ vVel) {
  const int i = 1 - g.OLx + blockIdx.x * blockDim.x + threadIdx.x;
  const int j = 1 - g.OLy + blockIdx.y * blockDim.y + threadIdx.y;
  if (i > g.sNx + g.OLx || j > g.sNy + g.OLy) return;
  const size_t s = g.s(i, j);
  cg2d_x[s] = Bo_surf[s] * etaN[s];                       // solve_for_pressure.F:129 (full halo range)
  double b = 0.;
  if (i >= 1 && i <= g.sNx && j >= 1 && j <= g.sNy) {
    const size_t slab = g.slab;
    const int PX = g.PX;
    const double dyG0 = g.dyG[s], dyG1 = g.dyG[s + 1], dxG0 = g.dxG[s], dxG1 = g.dxG[s + PX];
    const double *__restrict__ hW = g.hFacW, *__restrict__ hS = g.hFacS;
    int kW0 = 0, kW1 = 0, kS0 = 0, kS1 = 0;
    double lW0 = 0., lW1 = 0., lS0 = 0., lS1 = 0.;
    if (CG) {
      kW0 = g.kLowW[s]; kW1 = g.kLowW[s + 1]; kS0 = g.kLowS[s]; kS1 = g.kLowS[s + PX];
      lW0 = g.hLowW[s]; lW1 = g.hLowW[s + 1]; lS0 = g.hLowS[s]; lS1 = g.hLowS[s + PX];
    }
UNROLL_N((CG ? RHS_UNROLL_CG : RHS_UNROLL))
    for (int k = g.Nr; k >= 1; k--) {
      const size_t q = s + slab * (size_t)(k - 1);
      const double drFk = g.drF[k - 1];
      const double hW0 = CG ? cg_hfac(k, kW0, lW0) : hW[q], hW1 = CG ? cg_hfac(k, kW1, lW1) : hW[q + 1];
      const double hS0 = CG ? cg_hfac(k, kS0, lS0) : hS[q], hS1 = CG ? cg_hfac(k, kS1, lS1) : hS[q + PX];
      // calc_div_ghat.F:64-167: pf = xA*gU/deltaTMom with xA = dyG*drF*hFacW
      double px1, px0, py1, py0;
      if (implicDiv2DFlow == 1.) {
        px1 = dyG1 * drFk * hW1 * gU[q + 1] / deltaTMom; px0 = dyG0 * drFk * hW0 * gU[q] / deltaTMom;
        py1 = dxG1 * drFk * hS1 * gV[q + PX] / deltaTMom; py0 = dxG0 * drFk * hS0 * gV[q] / deltaTMom;
      } else if (!uVel) {      // exactConserv (calc_div_ghat.F:81-87): the explicit part lives in etaH
        px1 = implicDiv2DFlow * (dyG1 * drFk * hW1) * gU[q + 1] / deltaTMom;
        px0 = implicDiv2DFlow * (dyG0 * drFk * hW0) * gU[q] / deltaTMom;
        py1 = implicDiv2DFlow * (dxG1 * drFk * hS1) * gV[q + PX] / deltaTMom;
        py0 = implicDiv2DFlow * (dxG0 * drFk * hS0) * gV[q] / deltaTMom;
      } else {                 // calc_div_ghat.F:88-95
        const double om = 1. - implicDiv2DFlow;
        px1 = (implicDiv2DFlow * gU[q + 1] + om * uVel[q + 1]) * (dyG1 * drFk * hW1) / deltaTMom;
        px0 = (implicDiv2DFlow * gU[q] + om * uVel[q]) * (dyG0 * drFk * hW0) / deltaTMom;
        py1 = (implicDiv2DFlow * gV[q + PX] + om * vVel[q + PX]) * (dxG1 * drFk * hS1) / deltaTMom;
        py0 = (implicDiv2DFlow * gV[q] + om * vVel[q]) * (dxG0 * drFk * hS0) / deltaTMom;
      }
      b = b + px1 - px0;
      b = b + py1 - py0;
    }
    // free-surface term on etaH (exactConserv, solve_for_pressure.F:213-222) or etaN (:224-233)
    b = b - freeSurfFac * g.rA[s] * 1. / deltaTMom / deltaTFreeSurf * etaFS[s];
  }
  cg2d_b[s] = b;
}

__global__ void eta_kernel(size_t n, const double *recip_Bo, const double *x, double *etaN) {
  size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (t < n) etaN[t] = recip_Bo[t] * x[t];               // solve_for_pressure.F:377-385
}

// CORRECTION_STEP covers the halo'd slab (momentum_correction_step.F:60-63); INTEGR_CONTINUITY's exactConserv
// sum reads u(sNx+1,j), v(i,sNy+1) before the exchange.  Launched only with exactConserv: keeping the two
// predicated stores in corr_kernel's level loop cost 0.6 ms of 3.2 at 2048^2 x 50.
__global__ void corr_edge_kernel(TileGrid g, const double *__restrict__ gU, const double *__restrict__ gV,
                                 const double *__restrict__ etaN, const double *__restrict__ Bo_surf,
                                 double *__restrict__ uVel, double *__restrict__ vVel, double deltaTMom, double implicSurfPress) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;      // 0..sNy-1: east edge, sNy..sNy+sNx-1: north edge
  const int k = 1 + blockIdx.y;
  if (n >= g.sNy + g.sNx) return;
  const double psFac = 1. * implicSurfPress;
  if (n < g.sNy) {
    const int i = g.sNx + 1, j = 1 + n;
    const size_t s = g.s(i, j), q = g.s3(i, j, k);
    const double px = g.recip_dxC[s] * (Bo_surf[s] * etaN[s] - Bo_surf[s - 1] * etaN[s - 1]);
    const double dpx = -psFac * px * g.maskW[q];
    uVel[q] = (gU[q] + deltaTMom * dpx) * g.maskW[q];
  } else {
    const int i = 1 + n - g.sNy, j = g.sNy + 1;
    const size_t s = g.s(i, j), q = g.s3(i, j, k);
    const double py = g.recip_dyC[s] * (Bo_surf[s] * etaN[s] - Bo_surf[s - g.PX] * etaN[s - g.PX]);
    const double dpy = -psFac * py * g.maskS[q];
    vVel[q] = (gV[q] + deltaTMom * dpy) * g.maskS[q];
  }
}

// ---- correction step + continuity ---------------------------------------------------------------
// 8 CTAs/SM (64 registers): 4.05 -> 3.22 ms; more CTAs spill, fewer leave too few loads in flight.
#ifndef CORR_MINB
#define CORR_MINB 8
#endif
#ifndef CORR_UNROLL
#define CORR_UNROLL 5
#endif
template <bool CG>
__global__ void __launch_bounds__(128, CG ? CORR_MINB_CG : CORR_MINB) corr_kernel(TileGrid g, const double *__restrict__ gU, const double *__restrict__ gV,
                                                   const double *__restrict__ etaN, const double *__restrict__ Bo_surf,
                                                   double *__restrict__ uVel, double *__restrict__ vVel, double *__restrict__ wVel,
                                                   double deltaTMom, double implicSurfPress, int rigidLid) {
  const int i = 1 + blockIdx.x * blockDim.x + threadIdx.x;
  const int j = 1 + blockIdx.y * blockDim.y + threadIdx.y;
  if (i > g.sNx || j > g.sNy) return;
  const double psFac = 1. * implicSurfPress;
  auto phiX = [&](int ii) { return g.recip_dxC[g.s(ii, j)] * (Bo_surf[g.s(ii, j)] * etaN[g.s(ii, j)] - Bo_surf[g.s(ii - 1, j)] * etaN[g.s(ii - 1, j)]); };
  auto phiY = [&](int jj) { return g.recip_dyC[g.s(i, jj)] * (Bo_surf[g.s(i, jj)] * etaN[g.s(i, jj)] - Bo_surf[g.s(i, jj - 1)] * etaN[g.s(i, jj - 1)]); };
  const double px0 = phiX(i), px1 = phiX(i + 1), py0 = phiY(j), py1 = phiY(j + 1);
  double wKp1 = 0.;
  int kW0 = 0, kW1 = 0, kS0 = 0, kS1 = 0, kC = 0;
  double lW0 = 0., lW1 = 0., lS0 = 0., lS1 = 0.;
  if (CG) {
    const size_t s = g.s(i, j);
    kW0 = g.kLowW[s]; kW1 = g.kLowW[s + 1]; kS0 = g.kLowS[s]; kS1 = g.kLowS[s + g.PX]; kC = g.kLowC[s];
    lW0 = g.hLowW[s]; lW1 = g.hLowW[s + 1]; lS0 = g.hLowS[s]; lS1 = g.hLowS[s + g.PX];
  }
UNROLL_N(CORR_UNROLL)
  for (int k = g.Nr; k >= 1; k--) {
    auto uNew = [&](int ii, double px, int kW) {
      size_t q = g.s3(ii, j, k);
      const double mW = CG ? cg_mask(k, kW) : g.maskW[q];
      double dpx = -psFac * px * mW;
      return (gU[q] + deltaTMom * dpx) * mW;
    };
    auto vNew = [&](int jj, double py, int kS) {
      size_t q = g.s3(i, jj, k);
      const double mS = CG ? cg_mask(k, kS) : g.maskS[q];
      double dpy = -psFac * py * mS;
      return (gV[q] + deltaTMom * dpy) * mS;
    };
    const double u0 = uNew(i, px0, kW0), u1 = uNew(i + 1, px1, kW1), v0 = vNew(j, py0, kS0), v1 = vNew(j + 1, py1, kS1);
    const size_t s3 = g.s3(i, j, k);
    uVel[s3] = u0;
    vVel[s3] = v0;
    // INTEGRATE_FOR_W
    const double uT0 = u0 * g.dyG[g.s(i, j)] * g.drF[k - 1] * (CG ? cg_hfac(k, kW0, lW0) : g.hFacW[s3]);
    const double uT1 = u1 * g.dyG[g.s(i + 1, j)] * g.drF[k - 1] * (CG ? cg_hfac(k, kW1, lW1) : g.hFacW[g.s3(i + 1, j, k)]);
    const double vT0 = v0 * g.dxG[g.s(i, j)] * g.drF[k - 1] * (CG ? cg_hfac(k, kS0, lS0) : g.hFacS[s3]);
    const double vT1 = v1 * g.dxG[g.s(i, j + 1)] * g.drF[k - 1] * (CG ? cg_hfac(k, kS1, lS1) : g.hFacS[g.s3(i, j + 1, k)]);
    const double conv2d = -(uT1 - uT0 + vT1 - vT0);
    const double mCk = CG ? cg_mask(k, kC) : g.maskC[s3];
    double wv;
    if (rigidLid) {
      const double mCkm1 = k == 1 ? 0. : (CG ? cg_mask(k - 1, kC) : g.maskC[g.s3(i, j, k - 1)]);
      if (k == 1) wv = 0.;
      else if (k == g.Nr) wv = conv2d * g.recip_rA[g.s(i, j)] * mCk * mCkm1;
      else wv = (wKp1 + conv2d * g.recip_rA[g.s(i, j)]) * mCk * mCkm1;
    } else {
      if (k == g.Nr) wv = conv2d * g.recip_rA[g.s(i, j)] * mCk;
      else wv = (wKp1 + conv2d * g.recip_rA[g.s(i, j)]) * mCk;
    }
    wVel[s3] = wv;
    wKp1 = wv;
  }
}

// Tensor maps of the arrays dyn_tma_kernel stages (one per array and tile; the encode is host-only and cached).
static bool dyn_tma_maps(const TileGrid &tg, const MomState &st, const double *guN, const double *gvN, const double *phi,
                         DynTmaMaps &m) {
  struct Key {
    const void *p; int PX, PY, nz, bx;
    bool operator<(const Key &o) const { return std::tie(p, PX, PY, nz, bx) < std::tie(o.p, o.PX, o.PY, o.nz, o.bx); }
  };
  static std::map<Key, CUtensorMap> cache;
  auto get = [&](CUtensorMap &out, const double *base, int nz, bool patch) -> bool {
    Key k{base, tg.PX, tg.PY, nz, patch ? FT_H : FT_Y};
    auto it = cache.find(k);
    if (it == cache.end()) {
      CUtensorMap t;
      if (!make_tmap3(&t, base, tg.PX, tg.PY, (size_t)nz, FT_W, patch ? FT_H : FT_Y)) return false;
      if (cache.size() > 4096) cache.clear();
      it = cache.emplace(k, t).first;
    }
    out = it->second;
    return true;
  };
  const int Nr = tg.Nr;
  return get(m.u, st.u, Nr, true) && get(m.v, st.v, Nr, true) && get(m.hW, tg.hFacW, Nr, true) && get(m.hS, tg.hFacS, Nr, true) &&
         get(m.hC, tg.hFacC, Nr, true) && get(m.w, st.w, Nr, true) && get(m.mC, tg.maskC, Nr, true) &&
         get(m.phi, phi ? phi : st.u, Nr, true) && get(m.kU, st.kapU, Nr + 1, false) && get(m.kV, st.kapV, Nr + 1, false) &&
         get(m.rhW, tg.recip_hFacW, Nr, false) && get(m.rhS, tg.recip_hFacS, Nr, false) && get(m.guO, guN, Nr, false) &&
         get(m.gvO, gvN, Nr, false) && get(m.mW, tg.maskW, Nr, false) && get(m.mS, tg.maskS, Nr, false);
}

// Thread-block shape of the column-marching kernels (rhs, corr): x extent = contiguous bytes per level
// and array a block touches.  MITGCM_B200_COLBLK="bx,by" overrides (tuning aid).
static dim3 col_block() {
  static dim3 b(0, 0);
  if (b.x == 0) {
    b = dim3(32, 4);
    if (const char *e = getenv("MITGCM_B200_COLBLK")) {
      int x = 0, y = 0;
      if (sscanf(e, "%d,%d", &x, &y) == 2 && x >= 32 && x % 32 == 0 && y >= 1 && x * y <= 128) b = dim3(x, y);
    }
  }
  return b;
}

// part 0: THERMODYNAMICS, DYNAMICS, SOLVE_FOR_PRESSURE up to and including CG2D
// part 1: etaN = recip_Bo * cg2d_x (after the halo update of cg2d_x), MOMENTUM_CORRECTION_STEP, INTEGR_CONTINUITY
// The halo exchanges between and after the parts are done by the caller: locally (one rank) or
// over NCCL (mitgcm_b200/distributed.py).
static bool step_part(int part, int myIter, double *initRes, int *iters, double *lastRes, bool peerHalo = false) {
  Ctx &c = ctx();
  if (!c.ready) return fail(30, "mitgcm_b200_init_ not called");
  const Geom &g = c.g;
  const Params &q = c.p;
  if (q.I(MI_USECDSCHEME)) return fail(61, "forward_step: the CD scheme is not supported");
  const bool semiImpl = q.D(MP_IMPLICSURFPRESS) != 1.0 || q.D(MP_IMPLICDIV2DFLOW) != 1.0;
  // (an exch2 tile graph across ranks runs the single-process sequence: its exchanges reach the other ranks themselves)
  if (semiImpl && g.nPx * g.nPy > 1 && !exch2_active()) return fail(61, "forward_step: implicSurfPress/implicDiv2DFlow < 1 is single-rank for now");
  MomPar mp;
  if (!make_mom_par(mp)) return false;
  const bool vecinv = q.I(MI_VECTORINVARIANTMOMENTUM) != 0;
  ViPar vp{};
  if (vecinv) {
    if (!make_vi_par(vp)) return false;
    if (mp.momViscosity && mp.useBiharmonicVisc) return fail(61, "forward_step: MOM_VECINV with biharmonic viscosity is per-level only (mom_vecinv_b200_)");
    if (g.OLx < 2 || g.OLy < 2) return fail(61, "forward_step: MOM_VECINV needs OLx, OLy >= 2");
    if (!field(MG_RECIP_RAZ, false) || !field(MG_FCORIG, false)) return fail(43, "forward_step: rAz / fCoriG mirrors not set");
    vp.iMin = 0; vp.iMax = g.sNx + 1; vp.jMin = 0; vp.jMax = g.sNy + 1;
    if (exch2_active() && (int)c.csCorners.size() != g.nTiles) return fail(61, "forward_step: mitgcm_b200_set_cs_tiles_ not called");
  }
  const double abFac = (myIter == q.I(MI_NITER0)) ? 0.0 : 0.5 + q.D(MP_ABEPS);
  double *u = field(MG_UVEL), *v = field(MG_VVEL), *w = field(MG_WVEL);
  double *gU = field(MG_GU), *gV = field(MG_GV), *guN = field(MG_GUNM1), *gvN = field(MG_GVNM1);
  double *eta = field(MG_ETAN), *b = field(MG_CG2D_B), *x = field(MG_CG2D_X);
  double *sfU = field(MG_SURFFORCU), *sfV = field(MG_SURFFORCV), *Bo = field(MG_BO_SURF), *rBo = field(MG_RECIP_BO);
  double *kapU = field(MG_KAPPARU), *kapV = field(MG_KAPPARV);
  if (!u || !v || !w || !gU || !gV || !guN || !gvN || !eta || !b || !x || !sfU || !sfV || !Bo || !rBo || !kapU || !kapV) return false;
  const size_t ns = g.slab;
  dim3 blk(32, 4);
  // optional physics of the wider configurations (see include/mitgcm_b200.h, forward step)
  const bool buoy = q.I(MI_BUOYANCYLINEAR) != 0, relaxT = q.I(MI_DOTHETACLIMRELAX) != 0;
  const bool exactConserv = q.I(MI_EXACTCONSERV) != 0, implDiff = q.I(MI_IMPLICITDIFFUSION) != 0;
  // buoyancy and relaxation work on the halo'd slab of each rank (theta halos are exchanged every step);
  // the exactConserv update needs the u, v halo exchange in the middle of the continuity step
  if (exactConserv && g.nPx * g.nPy > 1 && !exch2_active()) return fail(61, "forward_step: exactConserv is single-rank for now");
  if (g.Nr > PHYS_NRMAX && implDiff) return fail(61, "forward_step: implicit diffusion supports Nr <= 64");
  if (g.Nr > PHYS_NRMAX && mp.momViscosity && mp.implicitViscosity) return fail(61, "forward_step: implicit viscosity supports Nr <= 64");
  double *rho = nullptr, *phiHyd = nullptr, *sfT = nullptr, *etaH = nullptr;
  const bool ivdc = q.D(MP_IVDC_KAPPA) != 0.;
  if (buoy && !(phiHyd = field(MG_PHIHYD))) return false;
  if (buoy && ivdc && !(rho = field(MG_RHOINSITU))) return false;
  if (relaxT && !(sfT = field(MG_SURFFORCT))) return false;
  if (exactConserv && !(etaH = field(MG_ETAH))) return false;
  const bool prof = q.I(MI_PROFILE) != 0;
  auto mark = [&](int n) {
    if (!prof) return;
    if (!c.pev[n]) cudaEventCreate(&c.pev[n]);
    cudaEventRecord(c.pev[n], c.stream);
  };
  col_geom_ready();      // (re)checks the z-level form of the geometry mirrors when they changed; attach_col_geom uses the result
  if (part == 0) {
  mark(0);
  // DO_OCEANIC_PHYS: surface relaxation forcing, in-situ density, convective flag -> kappaRT
  if (relaxT || (buoy && ivdc)) {
    double *th = field(MG_THETA), *sa = field(MG_SALT), *sst = field(MG_SST), *lam = field(MG_LAMBDATHETACLIMRELAX);
    double *tRef = field(MG_TREF), *sRef = field(MG_SREF), *kapT = field(MG_KAPPART), *sf = field(MG_SURFFORCT);
    double *rh = (buoy && ivdc) ? field(MG_RHOINSITU) : th;     // written only when the density is needed
    if (!th || !sa || !sst || !lam || !tRef || !sRef || !kapT || !sf || !rh) return false;
    EosLinear e{q.D(MP_RHONIL), q.D(MP_RHOCONST), q.D(MP_TALPHA), q.D(MP_SBETA)};
    dim3 grd((g.PX + 31) / 32, (g.PY + 3) / 4);
    for (int bj = 1; bj <= g.nSy; bj++)
      for (int bi = 1; bi <= g.nSx; bi++) {
        TileGrid tg;
        if (!make_tile_grid(bi, bj, tg)) return false;
        size_t t = (size_t)(bi - 1) + (size_t)g.nSx * (bj - 1);
        size_t o3 = ns * g.Nr * t, o2 = ns * t;
        c.launches++;
        ocean_phys_kernel<<<grd, blk, 0, c.stream>>>(tg, th + o3, sa + o3, sst + o2, lam + o2, tRef, sRef, e, q.D(MP_RKSIGN),
                                                     q.D(MP_IVDC_KAPPA), q.D(MP_DIFFKRT), relaxT ? 1 : 0, (buoy && ivdc) ? 1 : 0,
                                                     sf + o2, rh + o3, kapT + o3, q.D(MP_DIFFKRS),
                                                     (q.I(MI_SALTSTEPPING) && field(MG_KAPPARS)) ? field(MG_KAPPARS) + o3 : nullptr);
      }
    MG_CUDA(cudaGetLastError());
  }
  // CALC_PHI_HYD for all levels (dynamics.F:436-441 marches it level by level inside DYNAMICS; it only
  // depends on the density of DO_OCEANIC_PHYS, i.e. on theta BEFORE the thermodynamics step, so it runs here).
  // Without IVDC nobody else needs rhoInSitu and the EOS is evaluated on the fly.
  // CALC_PHI_HYD rides inside the dynamics kernel when that is the column-geometry TMA kernel in its default shape and the
  // density needs theta only (no IVDC density array, sBeta = 0): the kernel then integrates the potential of its patch
  // cells from the theta of BEFORE the thermodynamics step, which CYCLE_TRACER leaves in the other buffer.  The mirror
  // MG_PHIHYD is not written in that case (nothing else reads it).
  const bool fusePhi = buoy && !ivdc && q.D(MP_SBETA) == 0. && !vecinv && !semiImpl && c.cgState == 1 && dyn_tma_ok(g, mp) &&
                       !getenv("MITGCM_B200_NO_COLGEOM") && !getenv("MITGCM_B200_DYN_NOPIPE") && !getenv("MITGCM_B200_DYN_NOTMA") &&
                       !getenv("MITGCM_B200_DYN_TMA_NOSPLIT") && !getenv("MITGCM_B200_DYN_TMA_STAGES") &&
                       !getenv("MITGCM_B200_DYN_TMA_MINB") && !getenv("MITGCM_B200_NO_PHIFUSE");
  const double *thetaBefore = nullptr;      // theta this step started from (set once the tracer step has run)
  if (buoy && !fusePhi) {
    double *rF = field(MG_RF), *rC = field(MG_RC), *th = field(MG_THETA), *sa = field(MG_SALT);
    double *tRef = field(MG_TREF), *sRef = field(MG_SREF);
    if (!rF || !rC || !th || !sa || !tRef || !sRef) return false;
    EosLinear e{q.D(MP_RHONIL), q.D(MP_RHOCONST), q.D(MP_TALPHA), q.D(MP_SBETA)};
    dim3 grd((g.PX + 31) / 32, (g.PY + 3) / 4);
    for (int bj = 1; bj <= g.nSy; bj++)
      for (int bi = 1; bi <= g.nSx; bi++) {
        TileGrid tg;
        if (!make_tile_grid(bi, bj, tg)) return false;
        size_t o3 = ns * g.Nr * ((size_t)(bi - 1) + (size_t)g.nSx * (bj - 1));
        c.launches++;
        phihyd_kernel<<<grd, blk, 0, c.stream>>>(tg, ivdc ? rho + o3 : nullptr, th + o3, sa + o3, tRef, sRef, e, rF, rC,
                                                 q.D(MP_GRAVITY), 1.0 / q.D(MP_RHOCONST), phiHyd + o3);
      }
    MG_CUDA(cudaGetLastError());
  }
  // THERMODYNAMICS: TEMP_INTEGRATE, then SALT_INTEGRATE (thermodynamics.F:268-300), the same kernels on another field
  auto step_tracer = [&](int idT, int idT2, int idGN, int idKap, int advS, int vertS, double diffKh, double diffK4,
                         const double *sfT) -> bool {
    double *th = field(idT), *th2 = field(idT2), *gtN = field(idGN), *kapT = field(idKap);
    if (!th || !th2 || !gtN || !kapT) return false;
    GadPar p;
    p.k = 0; p.advScheme = advS; p.vertAdvScheme = vertS;
    p.calcAdvection = 1; p.implicitAdvection = 0; p.applyAB = 0; p.useDiffKr4 = 0;
    p.implicitDiffusion = q.I(MI_IMPLICITDIFFUSION);
    p.diffKh = diffKh; p.diffK4 = diffK4; p.rkSign = q.D(MP_RKSIGN);
    p.deltaT = q.D(MP_DELTATTRACER); p.diffKr4k = 0.;
    // gad_init_fixed.F:100-131: Adams-Bashforth on gT and the 1-D form only for schemes 2, 3, 4; every other scheme
    // is stepped forward in time and, unless multiDimAdvection = F, advected by GAD_ADVECTION
    const bool abScheme = p.advScheme == ADV_CENTERED_2ND || p.advScheme == ADV_UPWIND_3RD || p.advScheme == ADV_CENTERED_4TH;
    const bool multiDim = q.I(MI_MULTIDIMADVECTION) && !abScheme;
    double *gTadv = nullptr, *dTdev = nullptr;
    if (multiDim) {
      gTadv = field(MG_GT);
      std::vector<double> dTl((size_t)g.Nr, q.D(MP_DELTATTRACER));
      dTdev = to_device(dTl.data(), (size_t)g.Nr, 56, true);
      if (!gTadv || !dTdev) return false;
      MG_CUDA(cudaStreamSynchronize(c.stream));     // dTl is a local
      p.calcAdvection = 0;
    }
    dim3 grd((g.sNx + 31) / 32, (g.sNy + 3) / 4);
    for (int bj = 1; bj <= g.nSy; bj++)
      for (int bi = 1; bi <= g.nSx; bi++) {
        TileGrid tg;
        if (!make_tile_grid(bi, bj, tg)) return false;
        size_t o2 = ns * ((size_t)(bi - 1) + (size_t)g.nSx * (bj - 1)), o3 = o2 * g.Nr;
        if (multiDim && !gad_advection_tile(tg, o2 / ns, advS, vertS, 0, u + o3, v + o3,
                                            w + o3, th + o3, gTadv + o3, dTdev)) return false;
        c.launches++;
        if (thermo_fast_ok(g, p) && !(getenv("MITGCM_B200_THERMO_NOPIPE") && !multiDim)) {
          if (!c.attrThermo) {      // per device / context: reset by mitgcm_b200_init_
            MG_CUDA(cudaFuncSetAttribute(thermo_pipe_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ThermoPipeSmemT<false>)));
            MG_CUDA(cudaFuncSetAttribute(thermo_pipe_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ThermoPipeSmemT<true>)));
            c.attrThermo = true;
          }
          const bool cgT = attach_col_geom(bi, bj, tg);
          (cgT ? thermo_pipe_kernel<true> : thermo_pipe_kernel<false>)
              <<<dim3((g.sNx + FT_X - 1) / FT_X, (g.sNy + FT_Y - 1) / FT_Y), dim3(FT_X, FT_Y),
                 cgT ? sizeof(ThermoPipeSmemT<true>) : sizeof(ThermoPipeSmemT<false>), c.stream>>>(
              tg, u + o3, v + o3, w + o3, th + o3, kapT + o3, th2 + o3, gtN + o3, p, abFac, sfT ? sfT + o2 : nullptr,
              multiDim ? gTadv + o3 : nullptr, abScheme ? 1 : 0);
        } else if (!multiDim && thermo_fast_ok(g, p))
          thermo_fast_kernel<<<dim3((g.sNx + FT_X - 1) / FT_X, (g.sNy + FT_Y - 1) / FT_Y), dim3(FT_X, FT_Y), 0, c.stream>>>(
              tg, u + o3, v + o3, w + o3, th + o3, kapT + o3, th2 + o3, gtN + o3, p, abFac, sfT ? sfT + o2 : nullptr);
        else
          thermo_kernel<<<grd, blk, 0, c.stream>>>(tg, u + o3, v + o3, w + o3, th + o3, kapT + o3, th2 + o3, gtN + o3, p, abFac,
                                                   sfT ? sfT + o2 : nullptr, multiDim ? gTadv + o3 : nullptr, abScheme ? 1 : 0);
        if (implDiff) {      // GAD_IMPLICIT_R on theta* (temp_integrate.F:495-503)
          c.launches++;
          impldiff_kernel<<<grd, blk, 0, c.stream>>>(tg, kapT + o3, th2 + o3, p.deltaT);
        }
      }
    MG_CUDA(cudaGetLastError());
    // CYCLE_TRACER: theta <- theta** (interior); halos follow in the blocking exchange below
    std::swap(c.fields[idT], c.fields[idT2]);
    return true;
  };
  if (q.I(MI_TEMPSTEPPING) && !step_tracer(MG_THETA, MG_THETA2, MG_GTNM1, MG_KAPPART, q.I(MI_TEMPADVSCHEME), q.I(MI_TEMPVERTADVSCHEME),
                                           q.D(MP_DIFFKHT), q.D(MP_DIFFK4T), sfT)) return false;
  if (q.I(MI_SALTSTEPPING) && !step_tracer(MG_SALT, MG_SALT2, MG_GSNM1, MG_KAPPARS, q.I(MI_SALTADVSCHEME), q.I(MI_SALTVERTADVSCHEME),
                                           q.D(MP_DIFFKHS), q.D(MP_DIFFK4S), nullptr)) return false;
  if (fusePhi) thetaBefore = field(q.I(MI_TEMPSTEPPING) ? MG_THETA2 : MG_THETA);      // CYCLE_TRACER swapped the buffers
  // multi-rank, peer pushes: the new theta / salt halos travel on the side stream while DYNAMICS and CG2D run
  // (the NCCL path of mitgcm_b200/distributed.py exchanges them at the end of the step instead)
  if (peerHalo) {
    int ids[2], n = 0;
    if (q.I(MI_TEMPSTEPPING)) ids[n++] = MG_THETA;
    if (q.I(MI_SALTSTEPPING)) ids[n++] = MG_SALT;
    if (n && !halo_exchange(ids, n, !getenv("MITGCM_B200_HALO_NOOVERLAP"))) return false;
  }
  mark(1);
  // DYNAMICS
  {
    dim3 grd((g.sNx + 2 + 31) / 32, (g.sNy + 2 + 3) / 4);
    for (int bj = 1; bj <= g.nSy; bj++)
      for (int bi = 1; bi <= g.nSx; bi++) {
        TileGrid tg;
        if (!make_tile_grid(bi, bj, tg)) return false;
        size_t t = (size_t)(bi - 1) + (size_t)g.nSx * (bj - 1);
        size_t o3 = ns * g.Nr * t, o3p = ns * (g.Nr + 1) * t, o2 = ns * t;
        MomState st{u + o3, v + o3, w + o3, kapU + o3p, kapV + o3p};
        c.launches++;
        if (vecinv) {
          vp.csCorners = c.csCorners.empty() ? 0 : c.csCorners[t];
          vp.myFace = c.csFace.empty() ? 0 : c.csFace[t];
          if (!semiImpl && vi_fast_ok(g, vp)) {
            if (!c.attrVi) {
              MG_CUDA(cudaFuncSetAttribute(vi_pipe_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ViPipeSmem)));
              MG_CUDA(cudaFuncSetAttribute(vi_pipe_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ViPipeSmem)));
              c.attrVi = true;
            }
            // the column-geometry form of this kernel was measured SLOWER (21.5 vs 19.9 ms at 2048^2 x 50: it is issue-bound,
            // and filling the geometry slots costs instructions): kept for experiments only (MITGCM_B200_VI_COLGEOM=1)
            ((getenv("MITGCM_B200_VI_COLGEOM") && attach_col_geom(bi, bj, tg)) ? vi_pipe_kernel<true> : vi_pipe_kernel<false>)<<<dim3((g.sNx + 2 + FT_X - 1) / FT_X, (g.sNy + 2 + FT_Y - 1) / FT_Y), dim3(FT_X, FT_Y), sizeof(ViPipeSmem),
                             c.stream>>>(tg, st, vp, sfU + o2, sfV + o2, gU + o3, gV + o3, guN + o3, gvN + o3, q.D(MP_DELTATMOM),
                                         abFac, q.I(MI_MOMFORCING), q.I(MI_MOMDISSIP_IN_AB), buoy ? phiHyd + o3 : nullptr);
          } else if (!getenv("MITGCM_B200_VI_NOTILE") && !vp.highOrderVorticity && !vp.upwindVorticity)   // C4 reads j+2: past the patch
            dyn_kernel<2><<<dim3((g.sNx + 2 + 31) / 32, (g.sNy + 2 + 7) / 8), dim3(32, 8), 0, c.stream>>>(
                tg, st, mp, vp, sfU + o2, sfV + o2, gU + o3, gV + o3, guN + o3, gvN + o3, q.D(MP_DELTATMOM), abFac,
                q.I(MI_MOMFORCING), q.I(MI_MOMDISSIP_IN_AB), buoy ? phiHyd + o3 : nullptr,
                q.D(MP_IMPLICSURFPRESS) != 1.0 ? eta + o2 : nullptr, Bo + o2, 1.0 * (1.0 - q.D(MP_IMPLICSURFPRESS)));
          else
          dyn_kernel<1><<<grd, blk, 0, c.stream>>>(tg, st, mp, vp, sfU + o2, sfV + o2, gU + o3, gV + o3, guN + o3, gvN + o3,
                                                      q.D(MP_DELTATMOM), abFac, q.I(MI_MOMFORCING), q.I(MI_MOMDISSIP_IN_AB),
                                                      buoy ? phiHyd + o3 : nullptr,
                                                      q.D(MP_IMPLICSURFPRESS) != 1.0 ? eta + o2 : nullptr, Bo + o2,
                                                      1.0 * (1.0 - q.D(MP_IMPLICSURFPRESS)));
        } else if (!semiImpl && dyn_tma_ok(g, mp) && !getenv("MITGCM_B200_DYN_NOPIPE") && !getenv("MITGCM_B200_DYN_NOTMA")) {
          // operands staged by TMA (dyn_tma.cuh); the cp.async kernel below stays as the A/B reference
          DynTmaMaps maps;
          const bool phiF = fusePhi && thetaBefore && c.cgState == 1;
          if (!dyn_tma_maps(tg, st, guN + o3, gvN + o3, phiF ? thetaBefore + o3 : (buoy ? phiHyd + o3 : nullptr), maps)) return fail(63, "forward_step: cuTensorMapEncodeTiled failed");
          PhiFuse pf{};
          if (phiF) {
            pf.rhoNil = q.D(MP_RHONIL); pf.dRho = q.D(MP_RHONIL) - q.D(MP_RHOCONST); pf.tAlpha = q.D(MP_TALPHA);
            pf.gravity = q.D(MP_GRAVITY); pf.recip_rhoConst = 1.0 / q.D(MP_RHOCONST);
            pf.tRef = field(MG_TREF); pf.rF = field(MG_RF); pf.rC = field(MG_RC);
            if (!pf.tRef || !pf.rF || !pf.rC) return false;
          } else if (fusePhi) return fail(64, "forward_step: fused CALC_PHI_HYD was planned but the dynamics kernel cannot take it");
          const int smemBytes = (int)sizeof(DynTmaSmem) + 128;
          if (!c.attrDynTma) {
            MG_CUDA(cudaFuncSetAttribute(dyn_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smemBytes));
            c.attrDynTma = true;
          }
          if (!getenv("MITGCM_B200_DYN_TMA_NOSPLIT")) {      // U and V on two thread groups of one CTA (16 warps)
            // ring depth: 3 = one 512-thread CTA per SM, 110 registers, no spills (measured at 2048^2 x 50: 8.86 ms; 2 stages x 2 CTAs
            // at 64 registers: 9.06; 4: 8.87; the 256-thread dyn_tma_kernel: 9.75)
            const int nst = getenv("MITGCM_B200_DYN_TMA_STAGES") ? atoi(getenv("MITGCM_B200_DYN_TMA_STAGES")) : 3;
            const dim3 grdT((g.sNx + 2 + FT_X - 1) / FT_X, (g.sNy + 2 + FT_Y - 1) / FT_Y), blkT(FT_X, FT_Y, 2);
#define DYN_UV_LAUNCH(NST, MINB, CGF, PHF)                                                                                    \
  {                                                                                                                           \
    const int smB = (int)(CGF ? sizeof(DynTmaSmemCG<NST>) : sizeof(DynTmaSmemN<NST>)) + 128;                                  \
    const int key = NST + 16 * MINB + (CGF ? 256 : 0) + (PHF ? 512 : 0);                                                      \
    if (c.attrDynTmaUV != key) {                                                                                              \
      MG_CUDA(cudaFuncSetAttribute(dyn_tma_uv_kernel<NST, MINB, CGF, PHF>, cudaFuncAttributeMaxDynamicSharedMemorySize, smB)); \
      c.attrDynTmaUV = key;                                                                                                   \
    }                                                                                                                         \
    dyn_tma_uv_kernel<NST, MINB, CGF, PHF><<<grdT, blkT, smB, c.stream>>>(maps, tg, st, mp, sfU + o2, sfV + o2, gU + o3,      \
                                                                          gV + o3, guN + o3, gvN + o3, q.D(MP_DELTATMOM),     \
                                                                          abFac, q.I(MI_MOMFORCING), q.I(MI_MOMDISSIP_IN_AB), \
                                                                          buoy ? 1 : 0, pf);                                  \
  }
            if (attach_col_geom(bi, bj, tg)) {      // geometry from (kLow, hLow) per column: 8 instead of 16 boxes per level
              // 20 KB stages: three of them fit twice per SM (measured at 2048^2 x 50: <3,2> 8.08 ms, <2,2> 8.12, <3,1> 8.39, <4,1> 8.44;
              // the 3-D-array form <3,1>: 8.76)
              const int minb = getenv("MITGCM_B200_DYN_TMA_MINB") ? atoi(getenv("MITGCM_B200_DYN_TMA_MINB")) : 2;
              if (phiF) DYN_UV_LAUNCH(3, 2, true, true)      // (fusePhi implies the default shape)
              else if (nst == 2) DYN_UV_LAUNCH(2, 2, true, false)
              else if (nst == 4) DYN_UV_LAUNCH(4, 1, true, false)
              else if (nst == 5) DYN_UV_LAUNCH(5, 1, true, false)
              else if (minb == 2) DYN_UV_LAUNCH(3, 2, true, false)
              else DYN_UV_LAUNCH(3, 1, true, false)
            } else if (phiF) return fail(64, "forward_step: fused CALC_PHI_HYD without the column geometry");
            else if (nst == 3) DYN_UV_LAUNCH(3, 1, false, false)
            else if (nst == 4) DYN_UV_LAUNCH(4, 1, false, false)
            else if (nst == 5) DYN_UV_LAUNCH(5, 1, false, false)
            else DYN_UV_LAUNCH(2, 2, false, false)
#undef DYN_UV_LAUNCH
          } else
          dyn_tma_kernel<<<dim3((g.sNx + 2 + FT_X - 1) / FT_X, (g.sNy + 2 + FT_Y - 1) / FT_Y), dim3(FT_X, FT_Y), smemBytes, c.stream>>>(
              maps, tg, st, mp, sfU + o2, sfV + o2, gU + o3, gV + o3, guN + o3, gvN + o3, q.D(MP_DELTATMOM), abFac,
              q.I(MI_MOMFORCING), q.I(MI_MOMDISSIP_IN_AB), buoy ? 1 : 0);
        } else if (!semiImpl && dyn_fast_ok(g, mp) && !getenv("MITGCM_B200_DYN_NOPIPE")) {
          if (!c.attrDyn) {
            MG_CUDA(cudaFuncSetAttribute(dyn_pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(DynPipeSmem)));
            c.attrDyn = true;
          }
          dyn_pipe_kernel<<<dim3((g.sNx + 2 + FT_X - 1) / FT_X, (g.sNy + 2 + FT_Y - 1) / FT_Y), dim3(FT_X, FT_Y), sizeof(DynPipeSmem),
                            c.stream>>>(tg, st, mp, sfU + o2, sfV + o2, gU + o3, gV + o3, guN + o3, gvN + o3, q.D(MP_DELTATMOM),
                                        abFac, q.I(MI_MOMFORCING), q.I(MI_MOMDISSIP_IN_AB), buoy ? phiHyd + o3 : nullptr);
        } else if (!buoy && !semiImpl && dyn_fast_ok(g, mp))
          dyn_fast_kernel<<<dim3((g.sNx + 2 + FT_X - 1) / FT_X, (g.sNy + 2 + FT_Y - 1) / FT_Y), dim3(FT_X, FT_Y), 0, c.stream>>>(
              tg, st, mp, sfU + o2, sfV + o2, gU + o3, gV + o3, guN + o3, gvN + o3, q.D(MP_DELTATMOM), abFac,
              q.I(MI_MOMFORCING), q.I(MI_MOMDISSIP_IN_AB));
        else
          dyn_kernel<0><<<grd, blk, 0, c.stream>>>(tg, st, mp, vp, sfU + o2, sfV + o2, gU + o3, gV + o3, guN + o3, gvN + o3,
                                                q.D(MP_DELTATMOM), abFac, q.I(MI_MOMFORCING), q.I(MI_MOMDISSIP_IN_AB),
                                                buoy ? phiHyd + o3 : nullptr,
                                                q.D(MP_IMPLICSURFPRESS) != 1.0 ? eta + o2 : nullptr, Bo + o2,
                                                1.0 * (1.0 - q.D(MP_IMPLICSURFPRESS)));
        if (mp.momViscosity && mp.implicitViscosity && g.Nr > 1) {     // dynamics.F:572-579: MOM_{U,V}_IMPLICIT_R on u*, v*
          c.launches++;
          momimpl_kernel<<<dim3((g.sNx + 1 + 31) / 32, (g.sNy + 1 + 3) / 4), blk, 0, c.stream>>>(tg, kapU + o3p, kapV + o3p, gU + o3, gV + o3,
                                                                                              q.D(MP_DELTATMOM), 3);
        }
      }
    MG_CUDA(cudaGetLastError());
  }
  mark(2);
  // SOLVE_FOR_PRESSURE
  {
    const dim3 cb = col_block();
    dim3 grd((g.PX + cb.x - 1) / cb.x, (g.PY + cb.y - 1) / cb.y);
    for (int bj = 1; bj <= g.nSy; bj++)
      for (int bi = 1; bi <= g.nSx; bi++) {
        TileGrid tg;
        if (!make_tile_grid(bi, bj, tg)) return false;
        size_t t = (size_t)(bi - 1) + (size_t)g.nSx * (bj - 1);
        size_t o3 = ns * g.Nr * t, o2 = ns * t;
        c.launches++;
        (attach_col_geom(bi, bj, tg) ? rhs_kernel<true> : rhs_kernel<false>)<<<grd, cb, 0, c.stream>>>(tg, gU + o3, gV + o3, eta + o2, (exactConserv ? etaH : eta) + o2, Bo + o2, b + o2, x + o2, q.D(MP_DELTATMOM),
                                              q.D(MP_DELTATFREESURF), q.D(MP_FREESURFFAC), q.D(MP_IMPLICDIV2DFLOW),
                                              exactConserv ? nullptr : u + o3, exactConserv ? nullptr : v + o3);
      }
    MG_CUDA(cudaGetLastError());
    // The side-stream exchange must be over before the cooperative solver starts: cg2d_kernel takes every register
    // of every SM and waits for its peers, so a spinning halo kernel that keeps one of its CTAs off the machine (or
    // a push kernel that cannot get on it) closes a wait cycle across the ranks.  theta's halo has had the whole
    // of DYNAMICS to travel.
    if (peerHalo && !halo_join()) return false;
    mark(3);
    int numIters = q.I(MI_CG2DMAXITERS), nIterMin = q.I(MI_CG2DUSEMINRESSOL) - 1;
    double first, minsq, last;
    if (!cg2d_run(q.I(MI_USESRCGSOLVER) != 0, b, x, &first, &minsq, &last, &numIters, &nIterMin)) return false;
    *initRes = first; *iters = numIters; *lastRes = last;
    mark(4);
  }
  return true;
  }   // part 0
  {
    c.launches++;
    eta_kernel<<<(unsigned)((g.n2 + 255) / 256), 256, 0, c.stream>>>(g.n2, rBo, x, eta);
    MG_CUDA(cudaGetLastError());
  }
  mark(5);
  // MOMENTUM_CORRECTION_STEP + INTEGR_CONTINUITY
  {
    const dim3 cb = col_block();
    dim3 grd((g.sNx + cb.x - 1) / cb.x, (g.sNy + cb.y - 1) / cb.y);
    for (int bj = 1; bj <= g.nSy; bj++)
      for (int bi = 1; bi <= g.nSx; bi++) {
        TileGrid tg;
        if (!make_tile_grid(bi, bj, tg)) return false;
        size_t t = (size_t)(bi - 1) + (size_t)g.nSx * (bj - 1);
        size_t o3 = ns * g.Nr * t, o2 = ns * t;
        c.launches++;
        (attach_col_geom(bi, bj, tg) ? corr_kernel<true> : corr_kernel<false>)<<<grd, cb, 0, c.stream>>>(tg, gU + o3, gV + o3, eta + o2, Bo + o2, u + o3, v + o3, w + o3,
                                               q.D(MP_DELTATMOM), q.D(MP_IMPLICSURFPRESS), q.I(MI_RIGIDLID));
        if (q.I(MI_EXACTCONSERV)) {
          c.launches++;
          corr_edge_kernel<<<dim3((g.sNx + g.sNy + 127) / 128, g.Nr), 128, 0, c.stream>>>(
              tg, gU + o3, gV + o3, eta + o2, Bo + o2, u + o3, v + o3, q.D(MP_DELTATMOM), q.D(MP_IMPLICSURFPRESS));
        }
      }
    MG_CUDA(cudaGetLastError());
  }
  mark(6);
  return true;
}

// INTEGR_CONTINUITY, exactConserv part (integr_continuity.F:120-215) + _EXCH_XY_RL(etaN) (:331-333) +
// UPDATE_ETAH (update_etah.F) with its EXCH_XY_RL(etaH); runs right after the correction kernel, which
// stores u(sNx+1), v(sNy+1) for it, and before the blocking exchanges, as in the reference.
static bool etah_update() {
  Ctx &c = ctx();
  const Geom &g = c.g;
  const Params &q = c.p;
  double *u = field(MG_UVEL), *v = field(MG_VVEL), *eta = field(MG_ETAN), *etaH = field(MG_ETAH), *dEtaHdt = field(MG_DETAHDT);
  if (!u || !v || !eta || !etaH || !dEtaHdt) return false;
  dim3 blk(32, 4), grd((g.sNx + 31) / 32, (g.sNy + 3) / 4);
  for (int bj = 1; bj <= g.nSy; bj++)
    for (int bi = 1; bi <= g.nSx; bi++) {
      TileGrid tg;
      if (!make_tile_grid(bi, bj, tg)) return false;
      size_t o2 = g.slab * ((size_t)(bi - 1) + (size_t)g.nSx * (bj - 1)), o3 = o2 * g.Nr;
      c.launches++;
      etah_kernel<<<grd, blk, 0, c.stream>>>(tg, u + o3, v + o3, etaH + o2, dEtaHdt + o2, eta + o2, q.D(MP_IMPLICDIV2DFLOW),
                                             q.D(MP_DELTATFREESURF));
    }
  MG_CUDA(cudaGetLastError());
  return exch_field(eta, 1) && exch_field(etaH, 1);
}

static bool forward_step(int myIter, double *initRes, int *iters, double *lastRes) {
  Ctx &c = ctx();
  if (!c.ready) return fail(30, "mitgcm_b200_init_ not called");
  const Geom &g = c.g;
  const Params &q = c.p;
  // ranks on the periodic process grid: peer pushes (halo.cu).  An exch2 tile graph spread over ranks takes the
  // single-process sequence below: exch2_field / exch2_uv_field gather across the peer arenas (exch2.cu)
  const bool multi = g.nPx * g.nPy > 1 && !exch2_active();
  if (g.nPx * g.nPy > 1 && !halo_connected())
    return fail(62, "forward_step: multi-rank runs need mitgcm_b200_comm_connect_ (or step through mitgcm_b200_step_part_ + NCCL exchanges)");
  const bool prof = q.I(MI_PROFILE) != 0;
  auto mark = [&](int n) {
    if (!prof) return;
    if (!c.pev[n]) cudaEventCreate(&c.pev[n]);
    cudaEventRecord(c.pev[n], c.stream);
  };
  if (multi) {
    // Across ranks every exchange is a set of peer pushes on the stream (halo.cu).  theta (and salt) are final once
    // THERMODYNAMICS has run, so their halos travel on the side stream while DYNAMICS and the solver run:
    // step_part(0) calls halo_exchange(..., side) right after the tracer kernels (see there).
    if (!step_part(0, myIter, initRes, iters, lastRes, true)) return false;
    const int idx = MG_CG2D_X;
    if (!halo_exchange(&idx, 1)) return false;
    if (!step_part(1, myIter, initRes, iters, lastRes, true)) return false;
    int ids[3] = {MG_UVEL, MG_VVEL, MG_WVEL};
    if (!halo_exchange(ids, 3)) return false;
    if (!halo_join()) return false;
    mark(7);
    return true;
  }
  if (!step_part(0, myIter, initRes, iters, lastRes)) return false;
  if (!exch_field(field(MG_CG2D_X), 1)) return false;
  if (!step_part(1, myIter, initRes, iters, lastRes)) return false;
  if (q.I(MI_EXACTCONSERV) && !etah_update()) return false;
  // DO_FIELDS_BLOCKING_EXCHANGES: EXCH_UV_XYZ_RL(uVel, vVel, .TRUE.) -- on an exch2 tile graph the vector
  // exchange swaps / negates components across rotated facet edges -- then the scalars
  if (exch2_active()) {
    if (!exch2_uv_field(field(MG_UVEL), field(MG_VVEL), g.Nr, true)) return false;
  } else if (!exch_field(field(MG_UVEL), g.Nr) || !exch_field(field(MG_VVEL), g.Nr)) return false;
  if (!exch_field(field(MG_WVEL), g.Nr)) return false;
  if (q.I(MI_TEMPSTEPPING) && !exch_field(field(MG_THETA), g.Nr)) return false;
  if (q.I(MI_SALTSTEPPING) && !exch_field(field(MG_SALT), g.Nr)) return false;
  mark(7);
  return true;
}

// ---- multi-rank halo exchange pieces (one tile per rank) ------------------------------------------
// X phase strips: OLx columns x sNy interior rows; Y phase strips: OLy rows x the full PX width
// (so corners propagate, exch1_rx.template:172-200).  dir: 0 = west, 1 = east, 2 = south, 3 = north.
__global__ void pack_kernel(const double *f, double *buf, int nz, int sNx, int sNy, int OLx, int OLy, int dir, int unpack,
                            double *fw) {
  const int PX = sNx + 2 * OLx, PY = sNy + 2 * OLy;
  const size_t slab = (size_t)PX * PY;
  const int w = dir < 2 ? OLx : PX, h = dir < 2 ? sNy : OLy;
  const size_t total = (size_t)w * h * nz;
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    int a = (int)(t % w), b = (int)((t / w) % h), k = (int)(t / ((size_t)w * h));
    int ii, jj;
    if (dir < 2) {
      jj = OLy + b;
      if (!unpack) ii = dir == 0 ? OLx + a : sNx + a;            // send: first / last OLx interior columns
      else ii = dir == 0 ? a : OLx + sNx + a;                    // receive: west / east halo
    } else {
      ii = a;
      if (!unpack) jj = dir == 2 ? OLy + b : sNy + b;
      else jj = dir == 2 ? b : OLy + sNy + b;
    }
    size_t idx = (size_t)ii + (size_t)PX * jj + slab * k;
    if (unpack) fw[idx] = buf[t];
    else buf[t] = f[idx];
  }
}

__global__ void exch_dir_kernel(double *f, int nz, int sNx, int sNy, int OLx, int OLy, int ydir) {
  const int PX = sNx + 2 * OLx, PY = sNy + 2 * OLy;
  const size_t slab = (size_t)PX * PY;
  const int w = ydir ? PX : OLx, h = ydir ? OLy : sNy;
  const size_t total = (size_t)w * h * nz;
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    int a = (int)(t % w), b = (int)((t / w) % h), k = (int)(t / ((size_t)w * h));
    double *p = f + slab * k;
    if (!ydir) {
      size_t row = (size_t)PX * (OLy + b);
      p[row + a] = p[row + sNx + a];
      p[row + OLx + sNx + a] = p[row + OLx + a];
    } else {
      p[(size_t)PX * b + a] = p[(size_t)PX * (sNy + b) + a];
      p[(size_t)PX * (OLy + sNy + b) + a] = p[(size_t)PX * (OLy + b) + a];
    }
  }
}

static int field_nz(int id) {
  const Geom &g = ctx().g;
  return (id >= 0 && id < MG_N2D) ? 1 : (id >= 100 && id < MG_N3D_END) ? g.Nr : (id >= 200 && id < MG_N3DP_END) ? g.Nr + 1 : 0;
}

}  // namespace mg

using namespace mg;

extern "C" {

void mitgcm_b200_step_part_(const int *part, const int *myIter, double *cg2d_init_res, int *cg2d_iters,
                            double *cg2d_last_res, int *ierr) {
  ctx().lastError = 0;
  *ierr = step_part(*part, *myIter, cg2d_init_res, cg2d_iters, cg2d_last_res) ? 0 : 1;
}

void mitgcm_b200_pack_(const int *id, const int *dir, double *buf, const int *unpack, int *ierr) {
  Ctx &c = ctx();
  c.lastError = 0;
  *ierr = 1;
  if (!c.ready) { fail(30, "mitgcm_b200_init_ not called"); return; }
  const Geom &g = c.g;
  if (g.nTiles != 1) { fail(70, "pack: one tile per rank"); return; }
  double *f = field(*id);
  int nz = field_nz(*id);
  if (!f || nz == 0 || *dir < 0 || *dir > 3) { fail(2, "pack: bad field or direction"); return; }
  c.launches++;
  pack_kernel<<<c.numSMs * 4, 256, 0, c.stream>>>(f, buf, nz, g.sNx, g.sNy, g.OLx, g.OLy, *dir, *unpack, f);
  if (cudaGetLastError() != cudaSuccess) { fail(5, "pack launch"); return; }
  *ierr = 0;
}

void mitgcm_b200_exch_dir_(const int *id, const int *ydir, int *ierr) {
  Ctx &c = ctx();
  c.lastError = 0;
  *ierr = 1;
  if (!c.ready) { fail(30, "mitgcm_b200_init_ not called"); return; }
  const Geom &g = c.g;
  if (g.nTiles != 1) { fail(70, "exch_dir: one tile per rank"); return; }
  double *f = field(*id);
  int nz = field_nz(*id);
  if (!f || nz == 0) { fail(2, "exch_dir: bad field"); return; }
  c.launches++;
  exch_dir_kernel<<<c.numSMs * 4, 256, 0, c.stream>>>(f, nz, g.sNx, g.sNy, g.OLx, g.OLy, *ydir);
  if (cudaGetLastError() != cudaSuccess) { fail(5, "exch_dir launch"); return; }
  *ierr = 0;
}

void mitgcm_b200_forward_step_(const int *myIter, double *cg2d_init_res, int *cg2d_iters, double *cg2d_last_res,
                               int *ierr) {
  ctx().lastError = 0;
  *ierr = forward_step(*myIter, cg2d_init_res, cg2d_iters, cg2d_last_res) ? 0 : 1;
}

void mitgcm_b200_exch_(const int *id, int *ierr) {
  Ctx &c = ctx();
  c.lastError = 0;
  *ierr = 1;
  if (!c.ready) { fail(30, "mitgcm_b200_init_ not called"); return; }
  double *f = field(*id);
  if (!f) return;
  int nz = field_nz(*id);
  if (nz == 0) { fail(2, "exch: not a tile array"); return; }
  if (!exch_field(f, nz)) return;
  if (cudaStreamSynchronize(c.stream) != cudaSuccess) { fail(6, "exch: stream error"); return; }
  *ierr = 0;
}

// MOM_U_IMPLICIT_R / MOM_V_IMPLICIT_R (pkg/mom_common/mom_{u,v}_implicit_r.F:6-8; caller dynamics.F:576-579): reference
// argument list + the COMMON array the routine solves in place (gU / gV).  implicitViscosity only.
static void mom_implicit_r(const double *kappaR, const int *bi, const int *bj, double *gFld, int which) {
  Ctx &c = ctx();
  c.lastError = 0;
  if (!c.ready) { fail(30, "mitgcm_b200_init_ not called"); return; }
  const Geom &g = c.g;
  const Params &q = c.p;
  if (!q.I(MI_IMPLICITVISCOSITY)) { fail(54, "mom_implicit_r_b200_: only implicitViscosity is on the B200 path (momImplVertAdv, selectImplicitDrag are not)"); return; }
  if (g.Nr > PHYS_NRMAX) { fail(54, "mom_implicit_r_b200_: Nr exceeds PHYS_NRMAX"); return; }
  TileGrid tg;
  if (!make_tile_grid(*bi, *bj, tg)) { if (!c.lastError) fail(43, "grid mirrors not set"); return; }
  const size_t ns = g.slab, tile = (size_t)(*bi - 1) + (size_t)g.nSx * (size_t)(*bj - 1), off3 = ns * g.Nr * tile;
  double *dK = to_device(kappaR, ns * (g.Nr + 1), 43, true);
  double *dG = to_device(gFld, g.n3, 45, false);
  if (!dK || !dG) return;
  if (!is_device_ptr(gFld) &&
      cudaMemcpyAsync(dG + off3, gFld + off3, ns * g.Nr * sizeof(double), cudaMemcpyHostToDevice, c.stream) != cudaSuccess) { fail(4, "H2D copy failed"); return; }
  c.launches++;
  momimpl_kernel<<<dim3((g.sNx + 1 + 31) / 32, (g.sNy + 1 + 3) / 4), dim3(32, 4), 0, c.stream>>>(tg, dK, dK, dG + off3, dG + off3,
                                                                                              q.D(MP_DELTATMOM), which);
  if (cudaGetLastError() != cudaSuccess) { fail(5, "momimpl_kernel launch failed"); return; }
  if (!is_device_ptr(gFld)) cudaMemcpyAsync(gFld + off3, dG + off3, ns * g.Nr * sizeof(double), cudaMemcpyDeviceToHost, c.stream);
  if (cudaStreamSynchronize(c.stream) != cudaSuccess) fail(6, "mom_implicit_r_b200_: stream error");
}
void mom_u_implicit_r_b200_(const double *kappaRU, const int *bi, const int *bj, const double *myTime, const int *myIter,
                            const int *myThid, double *gU) {
  (void)myTime; (void)myIter; (void)myThid;
  mom_implicit_r(kappaRU, bi, bj, gU, 1);
}
void mom_v_implicit_r_b200_(const double *kappaRV, const int *bi, const int *bj, const double *myTime, const int *myIter,
                            const int *myThid, double *gV) {
  (void)myTime; (void)myIter; (void)myThid;
  mom_implicit_r(kappaRV, bi, bj, gV, 2);
}

void mitgcm_b200_set_cs_tiles_(const int *csCorners, const int *myFace, const int *edges, int *ierr) {
  Ctx &c = ctx();
  *ierr = 1;
  if (!c.ready) { fail(30, "mitgcm_b200_init_ not called"); return; }
  for (int t = 0; t < c.g.nTiles; t++)
    if (csCorners[t] < 0 || csCorners[t] > 15 || (csCorners[t] && myFace[t] < 1)) { fail(70, "set_cs_tiles: bad corner mask / facet number"); return; }
  c.csCorners.assign(csCorners, csCorners + c.g.nTiles);
  c.csFace.assign(myFace, myFace + c.g.nTiles);
  c.csEdges.assign(edges, edges + c.g.nTiles);
  *ierr = 0;
}

void mitgcm_b200_exch_uv_(const int *idU, const int *idV, const int *withSigns, int *ierr) {
  Ctx &c = ctx();
  c.lastError = 0;
  *ierr = 1;
  if (!c.ready) { fail(30, "mitgcm_b200_init_ not called"); return; }
  double *u = field(*idU), *v = field(*idV);
  if (!u || !v) return;
  const int nz = field_nz(*idU);
  if (nz == 0 || nz != field_nz(*idV)) { fail(2, "exch_uv: not a pair of like tile arrays"); return; }
  if (exch2_active()) {
    if (!exch2_uv_field(u, v, nz, *withSigns != 0)) return;
  } else {   // no vector rotation on a plain periodic tiling: two scalar exchanges (exch_uv_xyz_rx.template:86-96)
    if (!exch_field(u, nz) || !exch_field(v, nz)) return;
  }
  if (cudaStreamSynchronize(c.stream) != cudaSuccess) { fail(6, "exch_uv: stream error"); return; }
  *ierr = 0;
}

}  // extern "C"
