// gad.cuh -- point-wise restatement of GAD_CALC_RHS (pkg/generic_advdiff/gad_calc_rhs.F:193-781)
// and its leaf stencils as device functions.  A flux at a cell face is a pure function of the
// inputs, so a thread that needs fZon(i+1,j) or fMer(i,j+1) for its flux divergence simply
// evaluates it again: the result is bit-identical wherever it is computed, and the ~20 slab
// sweeps and 2-D temporaries of the reference (gad_calc_rhs.F:142-148) disappear.
// Expression order follows the Fortran exactly (compiled with -fmad=false).
//
// The accessor type A supplies the level-k inputs: either the slabs of the reference argument
// list (SlabAcc: drop-in, per-level call) or values derived on the fly from the resident 3-D
// state (FusedAcc in step.cu: CALC_ADV_FLOW fused in, model/src/calc_adv_flow.F).
#pragma once
#include "context.h"

namespace mg {

enum { ADV_UPWIND_1RST = 1, ADV_CENTERED_2ND = 2, ADV_UPWIND_3RD = 3, ADV_CENTERED_4TH = 4, ADV_DST2 = 20,
       ADV_FLUX_LIMIT = 77, ADV_DST3 = 30, ADV_DST3_FLUX_LIMIT = 33, ADV_OS7MP = 7 };

// Device views of the GRID.h mirrors for one tile (pointers already offset to the tile).
struct TileGrid {
  int sNx, sNy, OLx, OLy, Nr, PX, PY;
  size_t slab;
  const double *dxC, *dyC, *dxG, *dyG, *dxF, *dyF, *dxV, *dyU, *rA, *rAw, *rAs;
  const double *recip_dxC, *recip_dyC, *recip_dxF, *recip_dyF, *recip_dxV, *recip_dyU, *recip_rA, *recip_rAw,
      *recip_rAs;
  const double *recip_dxG, *recip_dyG, *recip_rAz, *fCoriG;   // vorticity-point metrics (MOM_VECINV)
  const double *fCori, *tanPhiAtU, *tanPhiAtV;
  const double *cosFacU, *cosFacV;                        // (PY) per tile
  const double *drF, *drC, *recip_drF, *recip_drC;        // vertical
  const double *hFacC, *hFacW, *hFacS, *recip_hFacC, *recip_hFacW, *recip_hFacS, *maskC, *maskW, *maskS;
  // column geometry (colgeom.cu; nullptr unless the resident step attached it): tile2d arrays
  const int *kLowC, *kLowW, *kLowS;
  const double *hLowC, *hLowW, *hLowS, *rhLowC, *rhLowW, *rhLowS;
  __device__ __forceinline__ size_t s(int i, int j) const {
    return (size_t)(i + OLx - 1) + (size_t)PX * (size_t)(j + OLy - 1);
  }
  __device__ __forceinline__ size_t s3(int i, int j, int k) const { return s(i, j) + slab * (size_t)(k - 1); }
};

// hFac (or, with the reciprocal at kLow, recip_hFac) and mask of level k of a z-level column
__device__ __forceinline__ double cg_hfac(int k, int kLow, double hLow) { return k < kLow ? 1. : (k == kLow ? hLow : 0.); }
__device__ __forceinline__ double cg_mask(int k, int kLow) { return k <= kLow ? 1. : 0.; }

bool make_tile_grid(int bi, int bj, TileGrid &tg);   // host: fills pointers from the mirrors
bool attach_col_geom(int bi, int bj, TileGrid &tg);  // host: adds the column geometry when it is valid (colgeom.cu)
// host: GAD_ADVECTION of one tile on device pointers (gad_advection.cu); dT = deltaTLev(Nr) on the device
bool gad_advection_tile(TileGrid tg, size_t tile, int advScheme, int vertScheme, int implicitAdvection, const double *u,
                        const double *v, const double *w, const double *tr, double *gT, const double *dT);

struct GadPar {
  int k, advScheme, vertAdvScheme, calcAdvection, implicitAdvection, applyAB, useDiffKr4, implicitDiffusion;
  double diffKh, diffK4, rkSign, deltaT;   // deltaT = deltaTLev(k)
  double diffKr4k;
};

__device__ __forceinline__ double gad_limiter(double Cr) { return fmax(0., fmax(fmin(1., 2. * Cr), fmin(2., Cr))); }

// 7th-order one-step method with monotonicity-preserving limiter: the arithmetic shared by
// GAD_OS7MP_ADV_X / _Y / _R (pkg/generic_advdiff/gad_os7mp_adv_x.F:132-206).  q[0..6] = the seven
// values along the flow (q[3] = Qi, the cell upstream of the face; q[0] = Qippp ... q[6] = Qimmm),
// m[0..5] = MskIpp, MskIp, MskI, MskIm, MskImm, MskImmm.  Returns trans*(Qi + Psi*DelIp).
__device__ inline double gad_os7mp(double trans, double cfl, const double q[7], const double m[6]) {
  const double Eps = 1.e-20;
  const double Qipp = q[1], Qip = q[2], Qi = q[3], Qim = q[4], Qimm = q[5], Qimmm = q[6];
  const double MskIp = m[1], MskI = m[2], MskIm = m[3], MskImm = m[4], MskImmm = m[5];
  double Fac = 1.;
  const double DelP = (Qip - Qi) * MskI;
  double Phi = Fac * DelP;
  Fac = Fac * (cfl + 1.) / 3.;
  const double DelM = (Qi - Qim) * MskIm;
  const double Del2 = DelP - DelM;
  Phi = Phi - Fac * Del2;
  Fac = Fac * (cfl - 2.) / 4.;
  const double DelPP = (Qipp - Qip) * MskIp * MskI;
  const double Del2P = DelPP - DelP;
  const double Del3P = Del2P - Del2;
  Phi = Phi + Fac * Del3P;
  Fac = Fac * (cfl - 3.) / 5.;
  const double DelMM = (Qim - Qimm) * MskImm * MskIm;
  const double Del2M = DelM - DelMM;
  const double Del3M = Del2 - Del2M;
  const double Del4 = Del3P - Del3M;
  Phi = Phi + Fac * Del4;
  Fac = Fac * (cfl + 2.) / 6.;
  const double Del2PP = DelPP - DelP;          // as coded in the reference (:151): DelPPP is not used
  const double Del3PP = Del2PP - Del2P;
  const double Del4P = Del3PP - Del3P;
  const double Del5P = Del4P - Del4;
  Phi = Phi + Fac * Del5P;
  Fac = Fac * (cfl + 2.) / 7.;
  const double DelMMM = (Qimm - Qimmm) * MskImmm * MskImm * MskIm;
  const double Del2MM = DelMM - DelMMM;
  const double Del3MM = Del2M - Del2MM;
  const double Del4M = Del3M - Del3MM;
  const double Del5M = Del4 - Del4M;
  const double Del6 = Del5P - Del5M;
  Phi = Phi - Fac * Del6;
  const double DelIp = (Qip - Qi) * MskI;
  const double recip_DelIp = copysign(1., DelIp) / fmax(fabs(DelIp), Eps);
  Phi = Phi * recip_DelIp;
  const double DelI = (Qi - Qim) * MskIm;
  const double recip_DelI = copysign(1., DelI) / fmax(fabs(DelI), Eps);
  const double rp1h = DelI * recip_DelIp;
  const double rp1h_cfl = rp1h / (cfl + Eps);
  const double d2 = Del2, d2p1 = Del2P, d2m1 = Del2M;
  double A = 4. * d2 - d2p1, B = 4. * d2p1 - d2, C = d2, D = d2p1;
  const double dp1h = fmax(fmin(fmin(A, B), fmin(C, D)), 0.) + fmin(fmax(fmax(A, B), fmax(C, D)), 0.);
  A = 4. * d2m1 - d2; B = 4. * d2 - d2m1; C = d2m1; D = d2;
  const double dm1h = fmax(fmin(fmin(A, B), fmin(C, D)), 0.) + fmin(fmax(fmax(A, B), fmax(C, D)), 0.);
  const double PhiMD = 1. / (1. - cfl) * (DelIp - dp1h) * recip_DelIp;
  const double PhiLC = rp1h_cfl * (1. + dm1h * recip_DelI);
  const double PhiMin = fmax(fmin(0., PhiMD), fmin(fmin(0., 2. * rp1h_cfl), PhiLC));
  const double PhiMax = fmin(fmax(2. / (1. - cfl), PhiMD), fmax(fmax(0., 2. * rp1h_cfl), PhiLC));
  Phi = fmax(PhiMin, fmin(Phi, PhiMax));
  const double Psi = Phi * 0.5 * (1. - cfl);
  return trans * (Qi + Psi * DelIp);
}

// gad_*_adv_x.F / gad_*_adv_y.F: advective flux through the west (dir 0) / south (dir 1) face.
template <class A>
__device__ double gad_adv_h(const TileGrid &g, const A &a, const GadPar &p, int dir, int i, int j) {
  const int di = dir == 0, dj = dir == 1;
  const int scheme = p.advScheme;
  if (scheme == ADV_OS7MP) {   // gad_os7mp_adv_x.F:96-214 / gad_os7mp_adv_y.F: zero outside 5-OL .. sN+OL-3
    const int c = dir == 0 ? i : j, lo = 1 - (dir == 0 ? g.OLx : g.OLy) + 4, hi = (dir == 0 ? g.sNx + g.OLx : g.sNy + g.OLy) - 3;
    if (c < lo || c > hi) return 0.;
    const double uT = dir == 0 ? a.uTrans(i, j) : a.vTrans(i, j);
    if (uT == 0.) return 0.;
    const double vel = dir == 0 ? a.uFld(i, j) : a.vFld(i, j);
    const double cfl = fabs(vel * p.deltaT * (dir == 0 ? g.recip_dxC[g.s(i, j)] : g.recip_dyC[g.s(i, j)]));
    const double *mk = dir == 0 ? g.maskW : g.maskS;
    double q[7], m[6];
    const int sg = uT > 0. ? 1 : -1;     // upstream side: q[n] = Q(c + sg*(2-n)) (+1 shift when flowing backwards)
    const int sh = uT > 0. ? 0 : -1;
#pragma unroll
    for (int n = 0; n < 7; n++) {
      const int o = sg * (2 - n) + sh;
      q[n] = a.TA(i + o * di, j + o * dj, p.k);
    }
#pragma unroll
    for (int n = 0; n < 6; n++) {
      const int o = sg * (2 - n);
      m[n] = mk[g.s3(i + o * di, j + o * dj, p.k)];
    }
    return gad_os7mp(uT, cfl, q, m);
  }
  const bool narrow = scheme == ADV_CENTERED_2ND || scheme == ADV_UPWIND_1RST || scheme == ADV_DST2;
  // rows/columns the reference leaves at zero
  if (dir == 0) {
    if (i == 1 - g.OLx) return 0.;
    if (!narrow && (i == 2 - g.OLx || i == g.sNx + g.OLx)) return 0.;
  } else {
    if (j == 1 - g.OLy) return 0.;
    if (!narrow && (j == 2 - g.OLy || j == g.sNy + g.OLy)) return 0.;
  }
  const double oneSixth = 1.0 / 6.0;
  const double T0 = a.TA(i, j, p.k), Tm1 = a.TA(i - di, j - dj, p.k);
  const double uT = dir == 0 ? a.uTrans(i, j) : a.vTrans(i, j);
  if (scheme == ADV_CENTERED_2ND) return uT * (T0 + Tm1) * 0.5;
  const double vel = dir == 0 ? a.uFld(i, j) : a.vFld(i, j);
  const double rdC = dir == 0 ? g.recip_dxC[g.s(i, j)] : g.recip_dyC[g.s(i, j)];
  if (scheme == ADV_UPWIND_1RST || scheme == ADV_DST2) {
    const double xLimit = scheme == ADV_DST2 ? 1. : 0.;
    double uCFL = fabs(vel * p.deltaT * rdC);
    double uAbs = fabs(uT) * (1. - xLimit * (1. - uCFL));
    return (uT + uAbs) * 0.5 * Tm1 + (uT - uAbs) * 0.5 * T0;
  }
  const double *mk = dir == 0 ? g.maskW : g.maskS;
  const double Tp1 = a.TA(i + di, j + dj, p.k), Tm2 = a.TA(i - 2 * di, j - 2 * dj, p.k);
  const double Rjp = (Tp1 - T0) * mk[g.s3(i + di, j + dj, p.k)];
  const double Rj = (T0 - Tm1) * mk[g.s3(i, j, p.k)];
  const double Rjm = (Tm1 - Tm2) * mk[g.s3(i - di, j - dj, p.k)];
  if (scheme == ADV_UPWIND_3RD || scheme == ADV_CENTERED_4TH) {
    const double Rjjp = Rjp - Rj, Rjjm = Rj - Rjm;
    double v = uT * (T0 + Tm1 - oneSixth * (Rjjp + Rjjm)) * 0.5;
    if (scheme == ADV_UPWIND_3RD) return v + fabs(uT) * 0.5 * oneSixth * (Rjjp - Rjjm);
    return v + fabs(uT) * 0.5 * oneSixth * (Rjjp - Rjjm) *
                   (1. - mk[g.s3(i - di, j - dj, p.k)] * mk[g.s3(i + di, j + dj, p.k)]);
  }
  const double uCFL = fabs(vel * p.deltaT * rdC);
  if (scheme == ADV_FLUX_LIMIT) {
    const double CrMax = 1.e6;
    double Cr = (uT > 0.) ? Rjm : Rjp;
    if (fabs(Rj) * CrMax <= fabs(Cr)) Cr = copysign(CrMax, Cr) * copysign(1., Rj);
    else Cr = Cr / Rj;
    Cr = gad_limiter(Cr);
    return uT * (T0 + Tm1) * 0.5 - fabs(uT) * ((1. - Cr) + uCFL * Cr) * Rj * 0.5;
  }
  const double d0 = (2. - uCFL) * (1. - uCFL) * oneSixth;
  const double d1 = (1. - uCFL * uCFL) * oneSixth;
  if (scheme == ADV_DST3)
    return 0.5 * (uT + fabs(uT)) * (Tm1 + (d0 * Rj + d1 * Rjm)) + 0.5 * (uT - fabs(uT)) * (T0 - (d0 * Rj + d1 * Rjp));
  // ADV_DST3_FLUX_LIMIT
  const double thetaMax = 1.e20;
  double thetaP, thetaM;
  if (fabs(Rj) * thetaMax <= fabs(Rjm)) thetaP = copysign(thetaMax, Rjm * Rj);
  else thetaP = Rjm / Rj;
  if (fabs(Rj) * thetaMax <= fabs(Rjp)) thetaM = copysign(thetaMax, Rjp * Rj);
  else thetaM = Rjp / Rj;
  double psiP = d0 + d1 * thetaP;
  psiP = fmax(0., fmin(fmin(1., psiP), thetaP * (1. - uCFL) / (uCFL + 1.e-20)));
  double psiM = d0 + d1 * thetaM;
  psiM = fmax(0., fmin(fmin(1., psiM), thetaM * (1. - uCFL) / (uCFL + 1.e-20)));
  return 0.5 * (uT + fabs(uT)) * (Tm1 + psiP * Rj) + 0.5 * (uT - fabs(uT)) * (T0 - psiM * Rj);
}

// GAD_GRAD_X / GAD_GRAD_Y / GAD_DEL2 (gad_grad_x.F, gad_grad_y.F, gad_del2.F)
template <class A>
__device__ __forceinline__ double gad_dTdx(const TileGrid &g, const A &a, int k, int i, int j) {
  if (i == 1 - g.OLx) return 0.;
  return a.xA(i, j) * g.recip_dxC[g.s(i, j)] * (a.T(i, j, k) - a.T(i - 1, j, k));
}
template <class A>
__device__ __forceinline__ double gad_dTdy(const TileGrid &g, const A &a, int k, int i, int j) {
  if (j == 1 - g.OLy) return 0.;
  return a.yA(i, j) * g.recip_dyC[g.s(i, j)] * (a.T(i, j, k) - a.T(i, j - 1, k));
}
template <class A>
__device__ double gad_del2(const TileGrid &g, const A &a, int k, int i, int j) {
  if (i > g.sNx + g.OLx - 1 || j > g.sNy + g.OLy - 1) return 0.;
  return g.recip_rA[g.s(i, j)] * g.recip_drF[k - 1] * g.recip_hFacC[g.s3(i, j, k)] *
         ((gad_dTdx(g, a, k, i + 1, j) - gad_dTdx(g, a, k, i, j)) + (gad_dTdy(g, a, k, i, j + 1) - gad_dTdy(g, a, k, i, j)));
}

// Net flux through the west face of cell (i,j): gad_calc_rhs.F:245-355.
template <class A>
__device__ double gad_fzon(const TileGrid &g, const A &a, const GadPar &p, int i, int j) {
  double f = 0.;
  if (p.calcAdvection) f = f + gad_adv_h(g, a, p, 0, i, j);
  double df = 0.;
  if (p.diffKh != 0. && i != 1 - g.OLx)   // GAD_DIFF_X
    df = -p.diffKh * a.xA(i, j) * g.recip_dxC[g.s(i, j)] * (a.T(i, j, p.k) - a.T(i - 1, j, p.k)) * g.cosFacU[j + g.OLy - 1];
  if (p.diffK4 != 0. && i != 1 - g.OLx)   // GAD_BIHARM_X
    df = df + p.diffK4 * a.xA(i, j) * g.recip_dxC[g.s(i, j)] *
                  (gad_del2(g, a, p.k, i, j) - gad_del2(g, a, p.k, i - 1, j)) * g.cosFacU[j + g.OLy - 1];
  return f + df;
}

// Net flux through the south face: gad_calc_rhs.F:374-484.
template <class A>
__device__ double gad_fmer(const TileGrid &g, const A &a, const GadPar &p, int i, int j) {
  double f = 0.;
  if (p.calcAdvection) f = f + gad_adv_h(g, a, p, 1, i, j);
  double df = 0.;
  if (p.diffKh != 0. && j != 1 - g.OLy)   // GAD_DIFF_Y
    df = -p.diffKh * a.yA(i, j) * g.recip_dyC[g.s(i, j)] * (a.T(i, j, p.k) - a.T(i, j - 1, p.k));
  if (p.diffK4 != 0. && j != 1 - g.OLy)   // GAD_BIHARM_Y
    df = df + p.diffK4 * a.yA(i, j) * g.recip_dyC[g.s(i, j)] * (gad_del2(g, a, p.k, i, j) - gad_del2(g, a, p.k, i, j - 1));
  return f + df;
}

// Vertical advective flux at the upper interface of level k: gad_*_adv_r.F.
template <class A>
__device__ double gad_adv_r(const TileGrid &g, const A &a, const GadPar &p, int i, int j) {
  const int k = p.k, Nr = g.Nr, scheme = p.vertAdvScheme;
  const int km2 = max(1, k - 2), km1 = max(1, k - 1), kp1 = min(Nr, k + 1);
  const double oneSixth = 1.0 / 6.0;
  const double Tk = a.TA(i, j, k), Tkm1 = a.TA(i, j, km1);
  const double rT = a.rTrans(i, j);
  if (scheme == ADV_OS7MP) {   // gad_os7mp_adv_r.F:96-210 (rT < 0: upward flow, upstream cell is k-1)
    if (rT == 0.) return 0.;
    const double cfl = fabs(a.wFld(i, j) * p.deltaT * g.recip_drC[k - 1]);
    auto cl = [&](int kk) { return min(Nr, max(1, kk)); };
    double q[7], m[6];
    if (rT < 0.) {
      const int ks[8] = {cl(k + 2), cl(k + 1), k, cl(k - 1), cl(k - 2), cl(k - 3), cl(k - 4), 0};
#pragma unroll
      for (int n = 0; n < 7; n++) q[n] = a.TA(i, j, ks[n]);
#pragma unroll
      for (int n = 0; n < 6; n++) m[n] = g.maskC[g.s3(i, j, ks[n])] * (double)(ks[n] - ks[n + 1]);
    } else {
      const int ks[8] = {cl(k - 3), cl(k - 2), cl(k - 1), k, cl(k + 1), cl(k + 2), cl(k + 3), 0};
#pragma unroll
      for (int n = 0; n < 7; n++) q[n] = a.TA(i, j, ks[n]);
      // MskIpp = maskC(km2)*(km2-km3), MskIp = maskC(km1)*(km1-km2), MskI = maskC(k)*(k-km1),
      // MskIm = maskC(kp1)*(kp1-k), MskImm = maskC(kp2)*(kp2-kp1), MskImmm = maskC(kp3)*(kp3-kp2)
      m[0] = g.maskC[g.s3(i, j, ks[1])] * (double)(ks[1] - ks[0]);
      m[1] = g.maskC[g.s3(i, j, ks[2])] * (double)(ks[2] - ks[1]);
      m[2] = g.maskC[g.s3(i, j, ks[3])] * (double)(ks[3] - ks[2]);
      m[3] = g.maskC[g.s3(i, j, ks[4])] * (double)(ks[4] - ks[3]);
      m[4] = g.maskC[g.s3(i, j, ks[5])] * (double)(ks[5] - ks[4]);
      m[5] = g.maskC[g.s3(i, j, ks[6])] * (double)(ks[6] - ks[5]);
    }
    return gad_os7mp(rT, cfl, q, m);
  }
  const double mkm1 = g.maskC[g.s3(i, j, km1)];
  if (scheme == ADV_CENTERED_2ND) return mkm1 * rT * (Tk + Tkm1) * 0.5;
  if (scheme == ADV_UPWIND_1RST || scheme == ADV_DST2) {
    const double rLimit = scheme == ADV_DST2 ? 1. : 0.;
    double wCFL = fabs(a.wFld(i, j) * p.deltaT * g.recip_drC[k - 1]);
    double wAbs = fabs(rT) * p.rkSign * (1. - rLimit * (1. - wCFL));
    return mkm1 * ((rT + wAbs) * 0.5 * Tkm1 + (rT - wAbs) * 0.5 * Tk);
  }
  const double Tkm2 = a.TA(i, j, km2), Tkp1 = a.TA(i, j, kp1);
  if (scheme == ADV_UPWIND_3RD || scheme == ADV_CENTERED_4TH) {
    const double Rjp = (Tkp1 - Tk) * g.maskC[g.s3(i, j, kp1)];
    const double Rj = (Tk - Tkm1);
    const double Rjm = (Tkm1 - Tkm2) * (scheme == ADV_UPWIND_3RD ? g.maskC[g.s3(i, j, km2)] : mkm1);
    const double Rjjp = Rjp - Rj, Rjjm = Rj - Rjm;
    if (scheme == ADV_UPWIND_3RD)
      return mkm1 * (rT * ((Tk + Tkm1) * 0.5 - oneSixth * (Rjjm + Rjjp) * 0.5) + fabs(rT) * oneSixth * (Rjjm - Rjjp) * 0.5);
    double maskPM = 1.;
    if (k <= 2 || k >= Nr) maskPM = 0.;
    double maskBound = maskPM * g.maskC[g.s3(i, j, km2)] * g.maskC[g.s3(i, j, kp1)];
    return mkm1 * (rT * ((Tk + Tkm1) * 0.5 - oneSixth * (Rjjm + Rjjp) * 0.5) +
                   fabs(rT) * oneSixth * (Rjjm - Rjjp) * 0.5 * (1. - maskBound));
  }
  if (scheme == ADV_FLUX_LIMIT) {
    const double CrMax = 1.e6;
    double wCFL = fabs(a.wFld(i, j) * p.deltaT * g.recip_drC[k - 1]);
    const double Rjp = (Tkp1 - Tk) * g.maskC[g.s3(i, j, kp1)];
    const double Rj = (Tk - Tkm1);
    const double Rjm = (Tkm1 - Tkm2) * g.maskC[g.s3(i, j, km2)];
    double Cr = (rT < 0.) ? Rjm : Rjp;
    if (fabs(Rj) * CrMax <= fabs(Cr)) Cr = copysign(CrMax, Cr) * copysign(1., Rj);
    else Cr = Cr / Rj;
    Cr = gad_limiter(Cr);
    return mkm1 * (rT * (Tk + Tkm1) * 0.5 + fabs(rT) * ((1. - Cr) + wCFL * Cr) * Rj * 0.5);
  }
  const double Rjp = (Tk - Tkp1) * g.maskC[g.s3(i, j, kp1)];
  const double Rj = (Tkm1 - Tk) * g.maskC[g.s3(i, j, k)] * mkm1;
  const double Rjm = (Tkm2 - Tkm1) * mkm1;
  double cfl = fabs(a.wFld(i, j) * p.deltaT * g.recip_drC[k - 1]);
  const double d0 = (2. - cfl) * (1. - cfl) * oneSixth;
  const double d1 = (1. - cfl * cfl) * oneSixth;
  if (scheme == ADV_DST3)
    return 0.5 * (rT + fabs(rT)) * (Tk + (d0 * Rj + d1 * Rjp)) + 0.5 * (rT - fabs(rT)) * (Tkm1 - (d0 * Rj + d1 * Rjm));
  const double thetaMax = 1.e20;
  double thetaP, thetaM;
  if (fabs(Rj) * thetaMax <= fabs(Rjm)) thetaP = copysign(thetaMax, Rjm * Rj);
  else thetaP = Rjm / Rj;
  if (fabs(Rj) * thetaMax <= fabs(Rjp)) thetaM = copysign(thetaMax, Rjp * Rj);
  else thetaM = Rjp / Rj;
  double psiP = d0 + d1 * thetaP;
  psiP = fmax(0., fmin(fmin(1., psiP), thetaP * (1. - cfl) / (cfl + 1.e-20)));
  double psiM = d0 + d1 * thetaM;
  psiM = fmax(0., fmin(fmin(1., psiM), thetaM * (1. - cfl) / (cfl + 1.e-20)));
  return 0.5 * (rT + fabs(rT)) * (Tk + psiM * Rj) + 0.5 * (rT - fabs(rT)) * (Tkm1 - psiP * Rj);
}

// Net vertical flux at the upper interface of level k: gad_calc_rhs.F:502-632.
template <class A>
__device__ double gad_fver(const TileGrid &g, const A &a, const GadPar &p, int i, int j) {
  const int k = p.k, Nr = g.Nr;
  double f = 0.;
  if (p.calcAdvection && !p.implicitAdvection && k >= 2) f = f + gad_adv_r(g, a, p, i, j);
  double df = 0.;
  if (!p.implicitDiffusion && k != 1 && k <= Nr)   // GAD_DIFF_R
    df = -a.KappaR(i, j) * a.maskUp(i, j) * g.rA[g.s(i, j)] * g.recip_drC[k - 1] *
         (a.T(i, j, k) - a.T(i, j, k - 1)) * p.rkSign;
  if (p.useDiffKr4 && k >= 2) {                    // GAD_BIHARM_R
    double gradR[3], del2T[2];
    for (int n = 1; n <= 3; n++) {
      int km = k + n - 3, kl = k + n - 2;
      if (km < 1 || kl > Nr) gradR[n - 1] = 0.;
      else
        gradR[n - 1] = (a.T(i, j, kl) - a.T(i, j, km)) * g.recip_drC[kl - 1] * g.maskC[g.s3(i, j, kl)] * g.maskC[g.s3(i, j, km)];
    }
    for (int n = 1; n <= 2; n++) {
      int kl = k + n - 2;
      del2T[n - 1] = (gradR[n] - gradR[n - 1]) * g.recip_hFacC[g.s3(i, j, kl)];
    }
    double tmpFac = p.rkSign * g.recip_drC[k - 1];
    df = df + p.diffKr4k * (del2T[1] - del2T[0]) * tmpFac * g.rA[g.s(i, j)] * a.maskUp(i, j);
  }
  return f + df;
}

// Flux divergence, gad_calc_rhs.F:767-781; returns the new gTracer(i,j,k).
template <class A>
__device__ __forceinline__ double gad_tendency(const TileGrid &g, const A &a, const GadPar &p, int i, int j, double gT,
                                               double fZ0, double fZ1, double fM0, double fM1, double fVup,
                                               double fVdn) {
  double advFac = p.calcAdvection ? 1. : 0.;
  double rAdvFac = p.rkSign * advFac;
  if (p.implicitAdvection) rAdvFac = p.rkSign;
  const int k = p.k;
  return gT - g.recip_hFacC[g.s3(i, j, k)] * g.recip_drF[k - 1] * g.recip_rA[g.s(i, j)] *
                  ((fZ1 - fZ0) + (fM1 - fM0) + (fVdn - fVup) * p.rkSign -
                   a.T(i, j, k) * ((a.uTrans(i + 1, j) - a.uTrans(i, j)) * advFac + (a.vTrans(i, j + 1) - a.vTrans(i, j)) * advFac +
                                   (a.rTransKp1(i, j) - a.rTrans(i, j)) * rAdvFac));
}

}  // namespace mg
