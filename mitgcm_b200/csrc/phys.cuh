// phys.cuh -- column kernels for the physics that sits either side of the hot path in the wider
// configurations (verification/tutorial_baroclinic_gyre).
// All are one thread per (i,j) column marching in k, coalesced along i, HBM-bound:
//   ocean_phys_kernel : FORCING_SURF_RELAX (theta) (model/src/forcing_surf_relax.F:60-75) + the
//                       surfaceForcingT line of EXTERNAL_FORCING_SURF (:205-212, Qnet = 0);
//                       FIND_RHO_2D 'LINEAR' (find_rho.F) for every level; GRAD_SIGMA's sigmaR
//                       (grad_sigma.F:95-103) + CALC_IVDC (calc_ivdc.F:60-70) + CALC_3D_DIFFUSIVITY
//                       (calc_3d_diffusivity.F:95-130) -> kappaRT               R{T,S,maskC} W{rho,kappaRT}
//   impldiff_kernel   : GAD_IMPLICIT_R, diffusion only (pkg/generic_advdiff/gad_implicit_r.F:110-140)
//                       + SOLVE_TRIDIAGONAL default branch (model/src/solve_tridiagonal.F:205-270)
//                                                                         R{T*,kappaRT,recip_hFacC,maskC} W{T}
//   phihyd_kernel     : CALC_PHI_HYD 'OCEANIC', integr_GeoPot = 2, uniformFreeSurfLev
//                       (calc_phi_hyd.F:240-262)                                     R{rho} W{phiHyd}
//   etah_kernel       : INTEGR_CONTINUITY, exactConserv part (integr_continuity.F:120-215)
//                                                                  R{u,v,hFacW,hFacS,maskC} W{dEtaHdt,etaN}
// Every expression keeps the Fortran operation order (-fmad=false): results are bit-identical to the
// oracle restatement in oracle/phys_oracle.c, which is pinned to the experiment's golden output.
#pragma once
#include "gad.cuh"

namespace mg {

#ifndef PHYS_NRMAX
#define PHYS_NRMAX 64
#endif

struct EosLinear {
  double rhoNil, rhoConst, tAlpha, sBeta;
};

__global__ void __launch_bounds__(128)
ocean_phys_kernel(TileGrid g, const double *__restrict__ theta, const double *__restrict__ salt,
                  const double *__restrict__ SST, const double *__restrict__ lambdaT,
                  const double *__restrict__ tRef, const double *__restrict__ sRef, EosLinear e, double rkSign,
                  double ivdc_kappa, double diffKrT, int doRelax, int doRho,
                  double *__restrict__ sfT, double *__restrict__ rho, double *__restrict__ kapT,
                  double diffKrS, double *__restrict__ kapS) {
  const int i = 1 - g.OLx + blockIdx.x * 32 + threadIdx.x;
  const int j = 1 - g.OLy + blockIdx.y * 4 + threadIdx.y;
  if (i > g.sNx + g.OLx || j > g.sNy + g.OLy) return;
  const size_t s = g.s(i, j);
  if (doRelax) {
    const size_t s1 = g.s3(i, j, 1);
    double f = -lambdaT[s] * (theta[s1] - SST[s]) * g.drF[0] * g.hFacC[s1];
    f = f - 0.;
    sfT[s] = f;
  }
  if (!doRho) return;
  const double dRho = e.rhoNil - e.rhoConst;
  const double gravitySign = -1.;
  // levels from the bottom up, as DO_OCEANIC_PHYS does for the convective flag (k = Nr..2)
  double Tk = theta[g.s3(i, j, g.Nr)], Sk = salt[g.s3(i, j, g.Nr)];
  for (int k = g.Nr; k >= 1; k--) {
    const size_t s3 = g.s3(i, j, k);
    const double rhoK = e.rhoNil * (e.sBeta * (Sk - sRef[k - 1]) - e.tAlpha * (Tk - tRef[k - 1])) + dRho;
    rho[s3] = rhoK;
    double conv = 0.;
    double Tm = 0., Sm = 0.;
    if (k > 1) {
      const size_t sm = s3 - g.slab;
      Tm = theta[sm]; Sm = salt[sm];
      const double rhoKm1 = e.rhoNil * (e.sBeta * (Sm - sRef[k - 1]) - e.tAlpha * (Tm - tRef[k - 1])) + dRho;
      const double sigmaR = g.maskC[s3] * g.maskC[sm] * g.recip_drC[k - 1] * rkSign * (rhoK - rhoKm1);
      conv = (-sigmaR * gravitySign > 0.) ? 1. : 0.;
    }
    double kap = conv * ivdc_kappa + 0.;     // + KbryanLewis79 (diffKrBL79surf = diffKrBL79deep = 0)
    const double kap0 = kap;
    kap = kap + diffKrT;                     // + diffKrNrT(k)
    kapT[s3] = kap;
    if (kapS) kapS[s3] = kap0 + diffKrS;     // CALC_3D_DIFFUSIVITY for salt: + diffKrNrS(k)
    Tk = Tm; Sk = Sm;
  }
}

// theta* (after TIMESTEP_TRACER) -> theta(n+1): tridiagonal solve down the column.
__global__ void __launch_bounds__(128)
impldiff_kernel(TileGrid g, const double *__restrict__ kapT, double *__restrict__ T, double deltaT) {
  const int i = 1 + blockIdx.x * 32 + threadIdx.x;
  const int j = 1 + blockIdx.y * 4 + threadIdx.y;
  if (i > g.sNx || j > g.sNy) return;
  const int Nr = g.Nr;
  if (Nr <= 1) return;
  double cp[PHYS_NRMAX], yp[PHYS_NRMAX];
  double cpm = 0., ypm = 0.;
  double mCm1 = 0.;                                   // maskC(k-1)
  size_t s3 = g.s3(i, j, 1);
  double mC = g.maskC[s3], rh = g.recip_hFacC[s3], kapK = kapT[s3];
  for (int k = 1; k <= Nr; k++) {
    double mCp1 = 0., rhp = 0., kapP = 0.;
    if (k < Nr) { mCp1 = g.maskC[s3 + g.slab]; rhp = g.recip_hFacC[s3 + g.slab]; kapP = kapT[s3 + g.slab]; }
    double b5 = 0., d5 = 0.;
    if (k >= 2) b5 = -deltaT * mCm1 * rh * g.recip_drF[k - 1] * kapK * g.recip_drC[k - 1];
    if (k <= Nr - 1) d5 = -deltaT * mCp1 * rh * g.recip_drF[k - 1] * kapP * g.recip_drC[k];
    const double c5 = 1. - (b5 + d5);
    const double y = T[s3];
    double cpk, ypk;
    if (k == 1) {
      if (c5 != 0.) { const double rec = 1. / c5; cpk = d5 * rec; ypk = y * rec; }
      else { cpk = 0.; ypk = 0.; }
    } else {
      const double tmp = c5 - b5 * cpm;
      if (tmp != 0.) { const double rec = 1. / tmp; cpk = d5 * rec; ypk = (y - b5 * ypm) * rec; }
      else { cpk = 0.; ypk = 0.; }
    }
    cp[k - 1] = cpk; yp[k - 1] = ypk;
    cpm = cpk; ypm = ypk;
    mCm1 = mC; mC = mCp1; rh = rhp; kapK = kapP;
    s3 += g.slab;
  }
  s3 = g.s3(i, j, Nr);
  double yk = yp[Nr - 1];
  T[s3] = yk;
  for (int k = Nr - 1; k >= 1; k--) {
    s3 -= g.slab;
    yk = yp[k - 1] - cp[k - 1] * yk;
    T[s3] = yk;
  }
}

// MOM_U_IMPLICIT_R / MOM_V_IMPLICIT_R with implicitViscosity (pkg/mom_common/mom_{u,v}_implicit_r.F:100-300;
// momImplVertAdv = F, selectImplicitDrag = 0) + SOLVE_TRIDIAGONAL: u* (in gU) on i = 1..sNx+1, j = 1..sNy and
// v* (in gV) on i = 1..sNx, j = 1..sNy+1, one thread per column.   R{g?,kappaR?,recip_hFac?,mask?} W{g?}
__device__ inline void momimpl_column(const TileGrid &g, const double *__restrict__ kap, const double *__restrict__ rhF,
                                      const double *__restrict__ mask, double *__restrict__ y, size_t s, double deltaTMom) {
  const int Nr = g.Nr;
  double cp[PHYS_NRMAX], yp[PHYS_NRMAX];
  double cpm = 0., ypm = 0.;
  size_t s3 = s;
  double mKm1 = 0., mK = mask[s3];
  for (int k = 1; k <= Nr; k++) {
    const double mKp1 = k < Nr ? mask[s3 + g.slab] : 0.;
    const double rh = rhF[s3];
    double b5 = 0., d5 = 0.;
    if (k >= 2 && mKm1 == 1.) b5 = -deltaTMom * rh * g.recip_drF[k - 1] * kap[s3] * g.recip_drC[k - 1];
    if (k <= Nr - 1 && mKp1 == 1.) d5 = -deltaTMom * rh * g.recip_drF[k - 1] * kap[s3 + g.slab] * g.recip_drC[k];
    const double c5 = 1. - (b5 + d5);
    const double yk = y[s3];
    double cpk, ypk;
    if (k == 1) {
      if (c5 != 0.) { const double rec = 1. / c5; cpk = d5 * rec; ypk = yk * rec; }
      else { cpk = 0.; ypk = 0.; }
    } else {
      const double tmp = c5 - b5 * cpm;
      if (tmp != 0.) { const double rec = 1. / tmp; cpk = d5 * rec; ypk = (yk - b5 * ypm) * rec; }
      else { cpk = 0.; ypk = 0.; }
    }
    cp[k - 1] = cpk; yp[k - 1] = ypk;
    cpm = cpk; ypm = ypk;
    mKm1 = mK; mK = mKp1;
    s3 += g.slab;
  }
  s3 = s + g.slab * (size_t)(Nr - 1);
  double yk = yp[Nr - 1];
  y[s3] = yk;
  for (int k = Nr - 1; k >= 1; k--) {
    s3 -= g.slab;
    yk = yp[k - 1] - cp[k - 1] * yk;
    y[s3] = yk;
  }
}
// which: 1 = U, 2 = V, 3 = both
__global__ void __launch_bounds__(128)
momimpl_kernel(TileGrid g, const double *__restrict__ kapU, const double *__restrict__ kapV, double *__restrict__ gU,
               double *__restrict__ gV, double deltaTMom, int which) {
  const int i = 1 + blockIdx.x * 32 + threadIdx.x;
  const int j = 1 + blockIdx.y * 4 + threadIdx.y;
  if (i > g.sNx + 1 || j > g.sNy + 1 || g.Nr <= 1) return;
  const size_t s = g.s(i, j);
  if ((which & 1) && j <= g.sNy) momimpl_column(g, kapU, g.recip_hFacW, g.maskW, gU, s, deltaTMom);
  if ((which & 2) && i <= g.sNx) momimpl_column(g, kapV, g.recip_hFacS, g.maskS, gV, s, deltaTMom);
}

// phiHydC for every level; range = the whole slab (the gradient needs (i-1,j) and (i,j-1)).
// rho == nullptr: the in-situ density anomaly is evaluated on the fly from theta (and salt when
// sBeta != 0) with FIND_RHO_2D 'LINEAR' -- same expression as ocean_phys_kernel, so nothing but
// theta is read when no other consumer of rhoInSitu (IVDC) is active: R{theta} W{phiHyd}.
#ifndef PHI_MINB
#define PHI_MINB 8
#endif
#ifndef PHI_UNROLL
#define PHI_UNROLL 5
#endif
#define PHI_DO_PRAGMA(x) _Pragma(#x)
#define PHI_UNROLL_N(n) PHI_DO_PRAGMA(unroll n)
__global__ void __launch_bounds__(128, PHI_MINB)
phihyd_kernel(TileGrid g, const double *__restrict__ rho, const double *__restrict__ theta,
              const double *__restrict__ salt, const double *__restrict__ tRef, const double *__restrict__ sRef,
              EosLinear e, const double *__restrict__ rF, const double *__restrict__ rC,
              double gravity, double recip_rhoConst, double *__restrict__ phiHyd) {
  const int i = 1 - g.OLx + blockIdx.x * 32 + threadIdx.x;
  const int j = 1 - g.OLy + blockIdx.y * 4 + threadIdx.y;
  if (i > g.sNx + g.OLx || j > g.sNy + g.OLy) return;
  const double dRho = e.rhoNil - e.rhoConst;
  double phiF = 0.;
  PHI_UNROLL_N(PHI_UNROLL)
  for (int k = 1; k <= g.Nr; k++) {
    const size_t s3 = g.s3(i, j, k);
    double dRlocM = 0.5 * g.drC[k - 1] * 1.;
    if (k == 1) dRlocM = (rF[0] - rC[0]) * 1.;
    double dRlocP;
    if (k == g.Nr) dRlocP = (rC[k - 1] - rF[k]) * 1.;
    else dRlocP = 0.5 * g.drC[k] * 1.;
    double a;
    if (rho) a = rho[s3];
    else {
      const double sTerm = e.sBeta != 0. ? e.sBeta * (salt[s3] - sRef[k - 1]) : 0.;
      a = e.rhoNil * (sTerm - e.tAlpha * (theta[s3] - tRef[k - 1])) + dRho;
    }
    const double phiC = phiF + dRlocM * gravity * a * recip_rhoConst;
    phiF = phiC + dRlocP * gravity * a * recip_rhoConst;
    phiHyd[s3] = phiC;
  }
}

// exactConserv: dEtaHdt, the conserving update of etaN (integr_continuity.F:120-215) and UPDATE_ETAH
// (update_etah.F:52-68) on the interior; levels summed k = 1..Nr.  Reads u(sNx+1), v(sNy+1), which the
// correction kernel has stored.  The caller exchanges etaN and etaH afterwards.
__global__ void __launch_bounds__(128)
etah_kernel(TileGrid g, const double *__restrict__ u, const double *__restrict__ v, double *__restrict__ etaH,
            double *__restrict__ dEtaHdt, double *__restrict__ etaN, double implicDiv2DFlow, double deltaTFreeSurf) {
  const int i = 1 + blockIdx.x * 32 + threadIdx.x;
  const int j = 1 + blockIdx.y * 4 + threadIdx.y;
  if (i > g.sNx || j > g.sNy) return;
  const size_t s = g.s(i, j);
  const double dyG0 = g.dyG[s], dyG1 = g.dyG[s + 1], dxG0 = g.dxG[s], dxG1 = g.dxG[s + g.PX];
  double h = 0.;
#pragma unroll 5
  for (int k = 1; k <= g.Nr; k++) {
    const size_t q = s + g.slab * (size_t)(k - 1);
    const double drFk = g.drF[k - 1];
    const double u0 = u[q] * dyG0 * drFk * g.hFacW[q], u1 = u[q + 1] * dyG1 * drFk * g.hFacW[q + 1];
    const double v0 = v[q] * dxG0 * drFk * g.hFacS[q], v1 = v[q + g.PX] * dxG1 * drFk * g.hFacS[q + g.PX];
    h = h + g.maskC[q] * (u1 - u0 + v1 - v0);
  }
  const double d = -h * g.recip_rA[s] * 1. - 0.;
  dEtaHdt[s] = d;
  const double eN = etaH[s] + implicDiv2DFlow * d * deltaTFreeSurf;
  etaN[s] = eN;
  // UPDATE_ETAH: etaH = etaN for implicDiv2DFlow = 1, else etaN + (1 - implicDiv2DFlow)*dEtaHdt*deltaTFreeSurf
  etaH[s] = implicDiv2DFlow == 1. ? eN : eN + (1. - implicDiv2DFlow) * d * deltaTFreeSurf;
}

}  // namespace mg
