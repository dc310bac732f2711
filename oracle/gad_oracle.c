/* gad_oracle.c -- CPU restatement of GAD_CALC_RHS, its leaf stencils and
 * CALC_ADV_FLOW.  TEST INFRASTRUCTURE ONLY (see mitgcm_oracle.h).
 * Follows pkg/generic_advdiff/gad_calc_rhs.F:193-781 and the leaves
 * gad_{c2,dst2u1,u3,c4,dst3,dst3fl,fluxlimit,os7mp}_adv_{x,y,r}.F, gad_diff_{x,y,r}.F,
 * gad_grad_{x,y}.F, gad_del2.F, gad_biharm_{x,y,r}.F (loop ranges as in the
 * reference).  Not restated: GM/Redi, KPP, OBCS, Smolarkiewicz hack,
 * SMAG_3D diffusivity, cubed-sphere corner fill; maskInC == 1; deepFac/rhoFac == 1.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "mitgcm_oracle.h"

#define S(i, j) ((size_t)((i) + OLx - 1) + (size_t)px * (size_t)((j) + OLy - 1))
#define G2(a, i, j) (a)[S(i, j) + off2]
#define G3(a, i, j, k) (a)[S(i, j) + (size_t)px * py * (size_t)((k)-1) + off3]
#define K3(a, i, j, k) (a)[S(i, j) + (size_t)px * py * (size_t)((k)-1)]
#define FORALL for (int j = 1 - OLy; j <= sNy + OLy; j++) for (int i = 1 - OLx; i <= sNx + OLx; i++)
#define SETUP                                                                          \
  const og_dims *d = &g->d;                                                            \
  const int sNx = d->sNx, sNy = d->sNy, OLx = d->OLx, OLy = d->OLy, Nr = d->Nr;        \
  const size_t px = (size_t)(sNx + 2 * OLx), py = (size_t)(sNy + 2 * OLy);             \
  const size_t tile = (size_t)(bi - 1) + (size_t)d->nSx * (size_t)(bj - 1);            \
  const size_t off2 = px * py * tile, off3 = px * py * (size_t)Nr * tile;              \
  const size_t offc = py * tile;                                                       \
  (void)Nr; (void)off2; (void)off3; (void)offc;

enum { UPWIND_1RST = 1, CENTERED_2ND = 2, UPWIND_3RD = 3, CENTERED_4TH = 4, DST2 = 20,
       FLUX_LIMIT = 77, DST3 = 30, DST3_FLUX_LIMIT = 33, OS7MP = 7 };

static const double oneSixth = 1.0 / 6.0;

/* GAD_FLUX_LIMITER.h: Superbee */
static double Limiter(double Cr) { return fmax(0., fmax(fmin(1., 2. * Cr), fmin(2., Cr))); }

void og_calc_adv_flow(const og_grid *g, int bi, int bj, int k,
                      const double *uVel, const double *vVel, const double *wVel,
                      double *xA, double *yA, double *maskUp,
                      double *uFld, double *vFld, double *wFld,
                      double *uTrans, double *vTrans, double *rTrans, double *rTransKp1) {
  SETUP
  /* uVel.. are the global tile3d arrays; uFld.. receive the level-k slabs the
   * callers pass on (temp_integrate.F:357-370). rTrans is in/out: on entry the
   * value left by level k+1 (calc_adv_flow.F, non-AUTODIFF branch). */
  FORALL {
    xA[S(i, j)] = G2(g->dyG, i, j) * g->drF[k - 1] * G3(g->hFacW, i, j, k);
    yA[S(i, j)] = G2(g->dxG, i, j) * g->drF[k - 1] * G3(g->hFacS, i, j, k);
  }
  if (k == Nr) { FORALL rTransKp1[S(i, j)] = 0.; }
  else { FORALL rTransKp1[S(i, j)] = rTrans[S(i, j)]; }
  FORALL {
    uFld[S(i, j)] = G3(uVel, i, j, k);
    vFld[S(i, j)] = G3(vVel, i, j, k);
    wFld[S(i, j)] = G3(wVel, i, j, k);
    uTrans[S(i, j)] = G3(uVel, i, j, k) * xA[S(i, j)];
    vTrans[S(i, j)] = G3(vVel, i, j, k) * yA[S(i, j)];
  }
  if (k == 1) {
    FORALL { maskUp[S(i, j)] = 0.; rTrans[S(i, j)] = 0.; }
  } else {
    FORALL {
      maskUp[S(i, j)] = G3(g->maskC, i, j, k - 1) * G3(g->maskC, i, j, k);
      rTrans[S(i, j)] = G3(wVel, i, j, k) * G2(g->rA, i, j) * maskUp[S(i, j)];
    }
  }
}

/* Core of the 7th-order one-step method with the monotonicity-preserving limiter, shared by
 * GAD_OS7MP_ADV_X / _Y / _R (gad_os7mp_adv_x.F:132-206): Q are the seven values along the flow
 * (Qi = upstream of the face), M the six face masks, trans the volume transport through the face.
 * Returns trans*(Qi + Psi*DelIp). */
static double os7mp_core(double trans, double cfl, double Qippp, double Qipp, double Qip, double Qi, double Qim,
                         double Qimm, double Qimmm, double MskIpp, double MskIp, double MskI, double MskIm,
                         double MskImm, double MskImmm) {
  const double Eps = 1.e-20;
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
  const double DelPPP = (Qippp - Qipp) * MskIpp * MskIp * MskI;
  (void)DelPPP;                       /* computed but unused by the reference (:150) */
  const double Del2PP = DelPP - DelP;
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

/* horizontal advective flux, direction dir = 0 (x) / 1 (y).
 * gad_*_adv_x.F / gad_*_adv_y.F */
static void adv_h(const og_grid *g, int bi, int bj, int k, int dir, int scheme, double deltaTloc,
                  const double *trans, const double *vel, const double *maskLoc,
                  const double *tr, double *af) {
  SETUP
  const int di = dir == 0, dj = dir == 1;
  const double *recip_dC = dir == 0 ? g->recip_dxC : g->recip_dyC;
  if (scheme == OS7MP) {   /* gad_os7mp_adv_x.F:96-214, gad_os7mp_adv_y.F: seven-point stencil along dir */
    FORALL af[S(i, j)] = 0.;
    const int i0 = 1 - OLx + (di ? 4 : 0), i1 = sNx + OLx - (di ? 3 : 0);
    const int j0 = 1 - OLy + (dj ? 4 : 0), j1 = sNy + OLy - (dj ? 3 : 0);
    for (int j = j0; j <= j1; j++)
      for (int i = i0; i <= i1; i++) {
        const double uT = trans[S(i, j)];
        if (uT == 0.) { af[S(i, j)] = 0.; continue; }
        const double cfl = fabs(vel[S(i, j)] * deltaTloc * G2(recip_dC, i, j));
#define Q_(o) tr[S(i + (o) * di, j + (o) * dj)]
#define M_(o) maskLoc[S(i + (o) * di, j + (o) * dj)]
        if (uT > 0.)
          af[S(i, j)] = os7mp_core(uT, cfl, Q_(2), Q_(1), Q_(0), Q_(-1), Q_(-2), Q_(-3), Q_(-4),
                                   M_(2), M_(1), M_(0), M_(-1), M_(-2), M_(-3));
        else
          af[S(i, j)] = os7mp_core(uT, cfl, Q_(-3), Q_(-2), Q_(-1), Q_(0), Q_(1), Q_(2), Q_(3),
                                   M_(-2), M_(-1), M_(0), M_(1), M_(2), M_(3));
#undef Q_
#undef M_
      }
    return;
  }
  /* leading rows/columns the reference zeroes explicitly */
  if (scheme == CENTERED_2ND || scheme == UPWIND_1RST || scheme == DST2) {
    if (dir == 0) for (int j = 1 - OLy; j <= sNy + OLy; j++) af[S(1 - OLx, j)] = 0.;
    else for (int i = 1 - OLx; i <= sNx + OLx; i++) af[S(i, 1 - OLy)] = 0.;
  } else {
    if (dir == 0) for (int j = 1 - OLy; j <= sNy + OLy; j++) {
      af[S(1 - OLx, j)] = 0.; af[S(2 - OLx, j)] = 0.; af[S(sNx + OLx, j)] = 0.;
    } else for (int i = 1 - OLx; i <= sNx + OLx; i++) {
      af[S(i, 1 - OLy)] = 0.; af[S(i, 2 - OLy)] = 0.; af[S(i, sNy + OLy)] = 0.;
    }
  }
  const int wide = !(scheme == CENTERED_2ND || scheme == UPWIND_1RST || scheme == DST2);
  const int iLo = 1 - OLx + (di ? (wide ? 2 : 1) : 0), iHi = sNx + OLx - (di && wide ? 1 : 0);
  const int jLo = 1 - OLy + (dj ? (wide ? 2 : 1) : 0), jHi = sNy + OLy - (dj && wide ? 1 : 0);
  for (int j = jLo; j <= jHi; j++)
    for (int i = iLo; i <= iHi; i++) {
      const double T0 = tr[S(i, j)], Tm1 = tr[S(i - di, j - dj)];
      const double uT = trans[S(i, j)];
      if (scheme == CENTERED_2ND) {
        af[S(i, j)] = uT * (T0 + Tm1) * 0.5;
        continue;
      }
      if (scheme == UPWIND_1RST || scheme == DST2) {
        const double xLimit = scheme == DST2 ? 1. : 0.;
        double uCFL = fabs(vel[S(i, j)] * deltaTloc * G2(recip_dC, i, j));
        double uAbs = fabs(uT) * (1. - xLimit * (1. - uCFL));
        af[S(i, j)] = (uT + uAbs) * 0.5 * Tm1 + (uT - uAbs) * 0.5 * T0;
        continue;
      }
      const double Tp1 = tr[S(i + di, j + dj)], Tm2 = tr[S(i - 2 * di, j - 2 * dj)];
      const double Rjp = (Tp1 - T0) * maskLoc[S(i + di, j + dj)];
      const double Rj = (T0 - Tm1) * maskLoc[S(i, j)];
      const double Rjm = (Tm1 - Tm2) * maskLoc[S(i - di, j - dj)];
      if (scheme == UPWIND_3RD || scheme == CENTERED_4TH) {
        const double Rjjp = Rjp - Rj, Rjjm = Rj - Rjm;
        double v = uT * (T0 + Tm1 - oneSixth * (Rjjp + Rjjm)) * 0.5;
        if (scheme == UPWIND_3RD)
          v = v + fabs(uT) * 0.5 * oneSixth * (Rjjp - Rjjm);
        else {
          const double *mk = dir == 0 ? g->maskW : g->maskS;
          v = v + fabs(uT) * 0.5 * oneSixth * (Rjjp - Rjjm)
                  * (1. - G3(mk, i - di, j - dj, k) * G3(mk, i + di, j + dj, k));
        }
        af[S(i, j)] = v;
        continue;
      }
      const double uCFL = fabs(vel[S(i, j)] * deltaTloc * G2(recip_dC, i, j));
      if (scheme == FLUX_LIMIT) {
        const double CrMax = 1.e6;
        double Cr = (uT > 0.) ? Rjm : Rjp;
        if (fabs(Rj) * CrMax <= fabs(Cr)) Cr = copysign(CrMax, Cr) * copysign(1., Rj);
        else Cr = Cr / Rj;
        Cr = Limiter(Cr);
        af[S(i, j)] = uT * (T0 + Tm1) * 0.5 - fabs(uT) * ((1. - Cr) + uCFL * Cr) * Rj * 0.5;
        continue;
      }
      const double d0 = (2. - uCFL) * (1. - uCFL) * oneSixth;
      const double d1 = (1. - uCFL * uCFL) * oneSixth;
      if (scheme == DST3) {
        af[S(i, j)] = 0.5 * (uT + fabs(uT)) * (Tm1 + (d0 * Rj + d1 * Rjm))
                    + 0.5 * (uT - fabs(uT)) * (T0 - (d0 * Rj + d1 * Rjp));
      } else { /* DST3_FLUX_LIMIT */
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
        af[S(i, j)] = 0.5 * (uT + fabs(uT)) * (Tm1 + psiP * Rj)
                    + 0.5 * (uT - fabs(uT)) * (T0 - psiM * Rj);
      }
    }
}

/* vertical advective flux at interface k, gad_*_adv_r.F */
static void adv_r(const og_grid *g, const og_params *p, int bi, int bj, int k, int scheme,
                  double dTarg, const double *rTrans, const double *wFld, const double *tr,
                  double *wT) {
  SETUP
  const int km2 = k - 2 > 1 ? k - 2 : 1, km1 = k - 1 > 1 ? k - 1 : 1, kp1 = k + 1 < Nr ? k + 1 : Nr;
  if (scheme == OS7MP) {   /* gad_os7mp_adv_r.F:96-210 */
    const int km4 = k - 4 > 1 ? k - 4 : 1, km3 = k - 3 > 1 ? k - 3 : 1, kp2 = k + 2 < Nr ? k + 2 : Nr,
              kp3 = k + 3 < Nr ? k + 3 : Nr;
    FORALL {
      const double rT = rTrans[S(i, j)];
      if (rT == 0.) { wT[S(i, j)] = 0.; continue; }
      const double cfl = fabs(wFld[S(i, j)] * dTarg * g->recip_drC[k - 1]);
#define QK(kk) K3(tr, i, j, kk)
#define MK(kk, f) (G3(g->maskC, i, j, kk) * (double)(f))
      if (rT < 0.)
        wT[S(i, j)] = os7mp_core(rT, cfl, QK(kp2), QK(kp1), QK(k), QK(km1), QK(km2), QK(km3), QK(km4),
                                 MK(kp2, kp2 - kp1), MK(kp1, kp1 - k), MK(k, k - km1), MK(km1, km1 - km2),
                                 MK(km2, km2 - km3), MK(km3, km3 - km4));
      else
        wT[S(i, j)] = os7mp_core(rT, cfl, QK(km3), QK(km2), QK(km1), QK(k), QK(kp1), QK(kp2), QK(kp3),
                                 MK(km2, km2 - km3), MK(km1, km1 - km2), MK(k, k - km1), MK(kp1, kp1 - k),
                                 MK(kp2, kp2 - kp1), MK(kp3, kp3 - kp2));
#undef QK
#undef MK
    }
    return;
  }
  const int zeroTop = (k == 1 || k > Nr);
  if ((scheme == CENTERED_2ND || scheme == UPWIND_1RST || scheme == DST2 || scheme == UPWIND_3RD ||
       scheme == CENTERED_4TH) && zeroTop) {
    FORALL wT[S(i, j)] = 0.;
    return;
  }
  if (scheme == FLUX_LIMIT && k > Nr) { FORALL wT[S(i, j)] = 0.; return; }
  FORALL {
    const double Tk = K3(tr, i, j, k), Tkm1 = K3(tr, i, j, km1), Tkm2 = K3(tr, i, j, km2), Tkp1 = K3(tr, i, j, kp1);
    const double rT = rTrans[S(i, j)];
    const double mkm1 = G3(g->maskC, i, j, km1);
    if (scheme == CENTERED_2ND) {
      wT[S(i, j)] = mkm1 * rT * (Tk + Tkm1) * 0.5;
    } else if (scheme == UPWIND_1RST || scheme == DST2) {
      const double rLimit = scheme == DST2 ? 1. : 0.;
      double wCFL = fabs(wFld[S(i, j)] * dTarg * g->recip_drC[k - 1]);
      double wAbs = fabs(rT) * p->rkSign * (1. - rLimit * (1. - wCFL));
      wT[S(i, j)] = mkm1 * ((rT + wAbs) * 0.5 * Tkm1 + (rT - wAbs) * 0.5 * Tk);
    } else if (scheme == UPWIND_3RD || scheme == CENTERED_4TH) {
      const double Rjp = (Tkp1 - Tk) * G3(g->maskC, i, j, kp1);
      const double Rj = (Tk - Tkm1);
      const double Rjm = (Tkm1 - Tkm2) * (scheme == UPWIND_3RD ? G3(g->maskC, i, j, km2) : mkm1);
      const double Rjjp = Rjp - Rj, Rjjm = Rj - Rjm;
      if (scheme == UPWIND_3RD)
        wT[S(i, j)] = mkm1 * (rT * ((Tk + Tkm1) * 0.5 - oneSixth * (Rjjm + Rjjp) * 0.5)
                              + fabs(rT) * oneSixth * (Rjjm - Rjjp) * 0.5);
      else {
        double maskPM = 1.;
        if (k <= 2 || k >= Nr) maskPM = 0.;
        double maskBound = maskPM * G3(g->maskC, i, j, km2) * G3(g->maskC, i, j, kp1);
        wT[S(i, j)] = mkm1 * (rT * ((Tk + Tkm1) * 0.5 - oneSixth * (Rjjm + Rjjp) * 0.5)
                              + fabs(rT) * oneSixth * (Rjjm - Rjjp) * 0.5 * (1. - maskBound));
      }
    } else if (scheme == FLUX_LIMIT) {
      const double CrMax = 1.e6;
      double wCFL = fabs(wFld[S(i, j)] * dTarg * g->recip_drC[k - 1]);
      const double Rjp = (Tkp1 - Tk) * G3(g->maskC, i, j, kp1);
      const double Rj = (Tk - Tkm1);
      const double Rjm = (Tkm1 - Tkm2) * G3(g->maskC, i, j, km2);
      double Cr = (rT < 0.) ? Rjm : Rjp;
      if (fabs(Rj) * CrMax <= fabs(Cr)) Cr = copysign(CrMax, Cr) * copysign(1., Rj);
      else Cr = Cr / Rj;
      Cr = Limiter(Cr);
      wT[S(i, j)] = mkm1 * (rT * (Tk + Tkm1) * 0.5 + fabs(rT) * ((1. - Cr) + wCFL * Cr) * Rj * 0.5);
    } else { /* DST3 / DST3_FLUX_LIMIT */
      const double Rjp = (Tk - Tkp1) * G3(g->maskC, i, j, kp1);
      const double Rj = (Tkm1 - Tk) * G3(g->maskC, i, j, k) * mkm1;
      const double Rjm = (Tkm2 - Tkm1) * mkm1;
      double cfl = fabs(wFld[S(i, j)] * dTarg * g->recip_drC[k - 1]);
      const double d0 = (2. - cfl) * (1. - cfl) * oneSixth;
      const double d1 = (1. - cfl * cfl) * oneSixth;
      if (scheme == DST3) {
        wT[S(i, j)] = 0.5 * (rT + fabs(rT)) * (Tk + (d0 * Rj + d1 * Rjp))
                    + 0.5 * (rT - fabs(rT)) * (Tkm1 - (d0 * Rj + d1 * Rjm));
      } else {
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
        wT[S(i, j)] = 0.5 * (rT + fabs(rT)) * (Tk + psiM * Rj)
                    + 0.5 * (rT - fabs(rT)) * (Tkm1 - psiP * Rj);
      }
    }
  }
}

void og_gad_calc_rhs(const og_grid *g, const og_params *p, int bi, int bj,
                     int iMin, int iMax, int jMin, int jMax, int k, int kM1, int kUp, int kDown,
                     const double *xA, const double *yA, const double *maskUp,
                     const double *uFld, const double *vFld, const double *wFld,
                     const double *uTrans, const double *vTrans, const double *rTrans,
                     const double *rTransKp1, double diffKh, double diffK4,
                     const double *KappaR, const double *diffKr4,
                     const double *TracerN, const double *TracAB, const double *deltaTLev,
                     int advectionScheme, int vertAdvecScheme,
                     int calcAdvection, int implicitAdvection, int applyAB_onTracer,
                     int trUseDiffKr4,
                     double *fZon, double *fMer, double *fVerT, double *gTracer) {
  SETUP
  (void)iMin; (void)iMax; (void)jMin; (void)jMax; (void)kM1;
  const size_t ns = px * py;
  double *buf = (double *)calloc(ns * 7, sizeof(double));
  double *df4 = buf, *af = buf + ns, *df = buf + 2 * ns, *localT = buf + 3 * ns, *locABT = buf + 4 * ns,
         *maskLocW = buf + 5 * ns, *maskLocS = buf + 6 * ns;
  double *fVerUp = fVerT + ns * (size_t)(kUp - 1);
  const double *fVerDn = fVerT + ns * (size_t)(kDown - 1);

  double advFac = 0.;
  if (calcAdvection) advFac = 1.;
  double rAdvFac = p->rkSign * advFac;
  if (implicitAdvection) rAdvFac = p->rkSign;

  FORALL { fZon[S(i, j)] = 0.; fMer[S(i, j)] = 0.; fVerUp[S(i, j)] = 0.; }
  FORALL {
    localT[S(i, j)] = K3(TracerN, i, j, k);
    locABT[S(i, j)] = applyAB_onTracer ? K3(TracAB, i, j, k) : K3(TracerN, i, j, k);
  }
  /* del^2 T for the bi-harmonic term: GAD_GRAD_X/Y + GAD_DEL2, :226-242 */
  if (diffK4 != 0.) {
    for (int j = 1 - OLy; j <= sNy + OLy; j++) {
      fZon[S(1 - OLx, j)] = 0.;
      for (int i = 2 - OLx; i <= sNx + OLx; i++)
        fZon[S(i, j)] = xA[S(i, j)] * G2(g->recip_dxC, i, j) * (localT[S(i, j)] - localT[S(i - 1, j)]);
    }
    for (int i = 1 - OLx; i <= sNx + OLx; i++) fMer[S(i, 1 - OLy)] = 0.;
    for (int j = 2 - OLy; j <= sNy + OLy; j++)
      for (int i = 1 - OLx; i <= sNx + OLx; i++)
        fMer[S(i, j)] = yA[S(i, j)] * G2(g->recip_dyC, i, j) * (localT[S(i, j)] - localT[S(i, j - 1)]);
    for (int j = 1 - OLy; j <= sNy + OLy - 1; j++)
      for (int i = 1 - OLx; i <= sNx + OLx - 1; i++)
        df4[S(i, j)] = G2(g->recip_rA, i, j) * g->recip_drF[k - 1] * G3(g->recip_hFacC, i, j, k)
                       * ((fZon[S(i + 1, j)] - fZon[S(i, j)]) + (fMer[S(i, j + 1)] - fMer[S(i, j)]));
  }
  /* ---- X ---- :245-355 */
  FORALL fZon[S(i, j)] = 0.;
  if (calcAdvection) {
    FORALL maskLocW[S(i, j)] = G3(g->maskW, i, j, k);
    adv_h(g, bi, bj, k, 0, advectionScheme, deltaTLev[k - 1], uTrans, uFld, maskLocW, locABT, af);
    FORALL fZon[S(i, j)] = fZon[S(i, j)] + af[S(i, j)];
  }
  if (diffKh != 0.) {
    for (int j = 1 - OLy; j <= sNy + OLy; j++) {
      df[S(1 - OLx, j)] = 0.;
      for (int i = 2 - OLx; i <= sNx + OLx; i++)
        df[S(i, j)] = -diffKh * xA[S(i, j)] * G2(g->recip_dxC, i, j)
                      * (localT[S(i, j)] - localT[S(i - 1, j)]) * g->cosFacU[(j + OLy - 1) + offc];
    }
  } else {
    FORALL df[S(i, j)] = 0.;
  }
  if (diffK4 != 0.)
    for (int j = 1 - OLy; j <= sNy + OLy; j++)
      for (int i = 2 - OLx; i <= sNx + OLx; i++)
        df[S(i, j)] = df[S(i, j)] + diffK4 * xA[S(i, j)] * G2(g->recip_dxC, i, j)
                      * (df4[S(i, j)] - df4[S(i - 1, j)]) * g->cosFacU[(j + OLy - 1) + offc];
  FORALL fZon[S(i, j)] = fZon[S(i, j)] + df[S(i, j)];
  /* ---- Y ---- :374-484 */
  FORALL fMer[S(i, j)] = 0.;
  if (calcAdvection) {
    FORALL maskLocS[S(i, j)] = G3(g->maskS, i, j, k);
    adv_h(g, bi, bj, k, 1, advectionScheme, deltaTLev[k - 1], vTrans, vFld, maskLocS, locABT, af);
    FORALL fMer[S(i, j)] = fMer[S(i, j)] + af[S(i, j)];
  }
  if (diffKh != 0.) {
    for (int i = 1 - OLx; i <= sNx + OLx; i++) df[S(i, 1 - OLy)] = 0.;
    for (int j = 2 - OLy; j <= sNy + OLy; j++)
      for (int i = 1 - OLx; i <= sNx + OLx; i++)
        df[S(i, j)] = -diffKh * yA[S(i, j)] * G2(g->recip_dyC, i, j) * (localT[S(i, j)] - localT[S(i, j - 1)]);
  } else {
    FORALL df[S(i, j)] = 0.;
  }
  if (diffK4 != 0.)
    for (int j = 2 - OLy; j <= sNy + OLy; j++)
      for (int i = 1 - OLx; i <= sNx + OLx; i++)
        df[S(i, j)] = df[S(i, j)] + diffK4 * yA[S(i, j)] * G2(g->recip_dyC, i, j) * (df4[S(i, j)] - df4[S(i, j - 1)]);
  FORALL fMer[S(i, j)] = fMer[S(i, j)] + df[S(i, j)];
  /* ---- R ---- :502-632 */
  if (calcAdvection && !implicitAdvection && k >= 2) {
    adv_r(g, p, bi, bj, k, vertAdvecScheme, deltaTLev[k - 1], rTrans, wFld,
          applyAB_onTracer ? TracAB : TracerN, af);
    FORALL fVerUp[S(i, j)] = fVerUp[S(i, j)] + af[S(i, j)];
  }
  if (p->implicitDiffusion) {
    FORALL df[S(i, j)] = 0.;
  } else {
    /* GAD_DIFF_R */
    const int km1 = k - 1 > 1 ? k - 1 : 1;
    if (k == 1 || k > Nr) { FORALL df[S(i, j)] = 0.; }
    else FORALL df[S(i, j)] = -KappaR[S(i, j)] * maskUp[S(i, j)] * G2(g->rA, i, j) * g->recip_drC[k - 1]
                              * (K3(TracerN, i, j, k) - K3(TracerN, i, j, km1)) * p->rkSign;
  }
  if (trUseDiffKr4 && k >= 2) {
    /* GAD_BIHARM_R (tmpFac of the del2T loop is computed but unused in the reference) */
    double *gradR = (double *)calloc(ns * 3, sizeof(double));
    double *del2T = (double *)calloc(ns * 2, sizeof(double));
    for (int n = 1; n <= 3; n++) {
      int km = k + n - 3, kl = k + n - 2;
      if (km < 1 || kl > Nr) { FORALL gradR[S(i, j) + ns * (n - 1)] = 0.; }
      else {
        double tmpFac = g->recip_drC[kl - 1];
        FORALL gradR[S(i, j) + ns * (n - 1)] = (K3(TracerN, i, j, kl) - K3(TracerN, i, j, km))
            * tmpFac * G3(g->maskC, i, j, kl) * G3(g->maskC, i, j, km);
      }
    }
    for (int n = 1; n <= 2; n++) {
      int kl = k + n - 2;
      FORALL del2T[S(i, j) + ns * (n - 1)] = (gradR[S(i, j) + ns * n] - gradR[S(i, j) + ns * (n - 1)])
          * G3(g->recip_hFacC, i, j, kl);
    }
    double tmpFac = p->rkSign * g->recip_drC[k - 1];
    FORALL df[S(i, j)] = df[S(i, j)] + diffKr4[k - 1] * (del2T[S(i, j) + ns] - del2T[S(i, j)])
                         * tmpFac * G2(g->rA, i, j) * maskUp[S(i, j)];
    free(gradR); free(del2T);
  }
  FORALL fVerUp[S(i, j)] = fVerUp[S(i, j)] + df[S(i, j)];
  /* ---- divergence ---- :767-781 */
  for (int j = 1 - OLy; j <= sNy + OLy - 1; j++)
    for (int i = 1 - OLx; i <= sNx + OLx - 1; i++)
      K3(gTracer, i, j, k) = K3(gTracer, i, j, k)
          - G3(g->recip_hFacC, i, j, k) * g->recip_drF[k - 1] * G2(g->recip_rA, i, j)
            * ((fZon[S(i + 1, j)] - fZon[S(i, j)])
             + (fMer[S(i, j + 1)] - fMer[S(i, j)])
             + (fVerDn[S(i, j)] - fVerUp[S(i, j)]) * p->rkSign
             - localT[S(i, j)] * ((uTrans[S(i + 1, j)] - uTrans[S(i, j)]) * advFac
                                + (vTrans[S(i, j + 1)] - vTrans[S(i, j)]) * advFac
                                + (rTransKp1[S(i, j)] - rTrans[S(i, j)]) * rAdvFac));
  free(buf);
}

/* FILL_CS_CORNER_TR_RL (eesupp/src/fill_cs_corner_tr_rl.F:74-156), withSigns = .FALSE.; corners: 1 SW, 2 SE, 4 NE, 8 NW */
static void gad_fill_cs_corner_tr(const og_dims *d, int fill4dir, int corners, double *f) {
  const int sNx = d->sNx, sNy = d->sNy, OLx = d->OLx, OLy = d->OLy;
  const size_t px = (size_t)(sNx + 2 * OLx);
  if (!corners) return;
  for (int j = 1; j <= OLy; j++)
    for (int i = 1; i <= OLx; i++) {
      if (fill4dir == 1) {
        if (corners & 1) f[S(1 - i, 1 - j)] = f[S(1 - j, i)];
        if (corners & 2) f[S(sNx + i, 1 - j)] = f[S(sNx + j, i)];
        if (corners & 8) f[S(1 - i, sNy + j)] = f[S(1 - j, sNy + 1 - i)];
        if (corners & 4) f[S(sNx + i, sNy + j)] = f[S(sNx + j, sNy + 1 - i)];
      } else {
        if (corners & 1) f[S(1 - i, 1 - j)] = f[S(j, 1 - i)];
        if (corners & 2) f[S(sNx + i, 1 - j)] = f[S(sNx + 1 - j, 1 - i)];
        if (corners & 8) f[S(1 - i, sNy + j)] = f[S(j, sNy + i)];
        if (corners & 4) f[S(sNx + i, sNy + j)] = f[S(sNx + 1 - j, sNy + i)];
      }
    }
}

/* FILL_CS_CORNER_UV_RS (eesupp/src/fill_cs_corner_uv_rs.F:46-108), withSigns = .FALSE. (negOne = 1) */
static void gad_fill_cs_corner_uv(const og_dims *d, int corners, double *uF, double *vF) {
  const int sNx = d->sNx, sNy = d->sNy, OLx = d->OLx, OLy = d->OLy;
  const size_t px = (size_t)(sNx + 2 * OLx);
  if (corners & 1) {
    for (int j = 1; j <= OLy; j++) for (int i = 1; i <= OLx; i++) uF[S(1 - i, 1 - j)] = vF[S(1 - j, 1 + i)];
    for (int j = 1; j <= OLy; j++) for (int i = 1; i <= OLx; i++) vF[S(1 - i, 1 - j)] = uF[S(1 + j, 1 - i)];
  }
  if (corners & 2) {
    for (int j = 1; j <= OLy; j++) for (int i = 2; i <= OLx; i++) uF[S(sNx + i, 1 - j)] = vF[S(sNx + j, i)];
    for (int j = 1; j <= OLy; j++) for (int i = 1; i <= OLx; i++) vF[S(sNx + i, 1 - j)] = uF[S(sNx + 1 - j, 1 - i)];
  }
  if (corners & 8) {
    for (int j = 1; j <= OLy; j++) for (int i = 1; i <= OLx; i++) uF[S(1 - i, sNy + j)] = vF[S(1 - j, sNy + 1 - i)];
    for (int j = 2; j <= OLy; j++) for (int i = 1; i <= OLx; i++) vF[S(1 - i, sNy + j)] = uF[S(j, sNy + i)];
  }
  if (corners & 4) {
    for (int j = 1; j <= OLy; j++) for (int i = 2; i <= OLx; i++) uF[S(sNx + i, sNy + j)] = vF[S(sNx + j, sNy + 2 - i)];
    for (int j = 2; j <= OLy; j++) for (int i = 1; i <= OLx; i++) vF[S(sNx + i, sNy + j)] = uF[S(sNx + 2 - j, sNy + i)];
  }
}

/* ---- GAD_ADVECTION (pkg/generic_advdiff/gad_advection.F:240-1097), multi-dimensional
 * direct-space-time advection of one tracer on one tile, non-cube topology (npass = 2: X pass then
 * Y pass, each updating localTij over the halo'd slab), then the vertical flux pass k = Nr..1.
 * compressible != 0 selects the GAD_MULTIDIM_COMPRESSIBLE form (local volume updated with the
 * transport divergence, :480-490, :1018-1032), else the default form (-T*div(U) correction,
 * :491-499, :1034-1046).  Schemes: 1, 20, 77, 30, 33, 7.  uFld, vFld, wFld, tracer are tile3d;
 * gTracer is a per-tile (slab, Nr) array.  implicitAdvection (vertical part left to GAD_IMPLICIT_R)
 * only in the default form, as in the reference.  Returns 0, or 1 for an unsupported scheme. */
int og_gad_advection(const og_grid *g, const og_params *p, int bi, int bj, int advectionScheme,
                     int vertAdvecScheme, int implicitAdvection, int compressible,
                     const double *deltaTLev, const double *uFld, const double *vFld, const double *wFld,
                     const double *tracer, double *gTracer, int nCFace, int edges) {
  SETUP
  const size_t ns = px * py;
  /* cubed sphere (gad_advection.F:249-272): nCFace = exch2_myFace, edges = 1 N | 2 S | 4 E | 8 W of the facet */
  const int cube = nCFace > 0, npass = cube ? 3 : 2;
  const int N_edge = cube && (edges & 1), S_edge = cube && (edges & 2), E_edge = cube && (edges & 4), W_edge = cube && (edges & 8);
  const int corners = (W_edge && S_edge ? 1 : 0) | (E_edge && S_edge ? 2 : 0) | (E_edge && N_edge ? 4 : 0) | (W_edge && N_edge ? 8 : 0);
  const int ok = advectionScheme == UPWIND_1RST || advectionScheme == DST2 || advectionScheme == FLUX_LIMIT ||
                 advectionScheme == DST3 || advectionScheme == DST3_FLUX_LIMIT || advectionScheme == OS7MP;
  const int okv = vertAdvecScheme == UPWIND_1RST || vertAdvecScheme == DST2 || vertAdvecScheme == FLUX_LIMIT ||
                  vertAdvecScheme == DST3 || vertAdvecScheme == DST3_FLUX_LIMIT || vertAdvecScheme == OS7MP;
  if (!ok || (!implicitAdvection && !okv) || (implicitAdvection && compressible)) return 1;
  double *buf = (double *)calloc(ns * (12 + 2 * (size_t)Nr), sizeof(double));
  double *xA = buf, *yA = buf + ns, *uTrans = buf + 2 * ns, *vTrans = buf + 3 * ns, *localTij = buf + 4 * ns,
         *localVol = buf + 5 * ns, *maskLocW = buf + 6 * ns, *maskLocS = buf + 7 * ns, *af = buf + 8 * ns,
         *rTrans = buf + 9 * ns, *rTransKp = buf + 10 * ns;
  double *localT3d = buf + 12 * ns, *locVol3d = buf + (12 + (size_t)Nr) * ns;
  double *fVerT = (double *)calloc(2 * ns, sizeof(double));
  for (int k = 1; k <= Nr; k++) {
    const double *uK = &G3(uFld, 1 - OLx, 1 - OLy, k), *vK = &G3(vFld, 1 - OLx, 1 - OLy, k);
    FORALL {
      xA[S(i, j)] = G2(g->dyG, i, j) * 1. * g->drF[k - 1] * G3(g->hFacW, i, j, k);
      yA[S(i, j)] = G2(g->dxG, i, j) * 1. * g->drF[k - 1] * G3(g->hFacS, i, j, k);
    }
    FORALL {
      uTrans[S(i, j)] = G3(uFld, i, j, k) * xA[S(i, j)] * 1.;
      vTrans[S(i, j)] = G3(vFld, i, j, k) * yA[S(i, j)] * 1.;
    }
    FORALL {
      localTij[S(i, j)] = G3(tracer, i, j, k);
      localVol[S(i, j)] = G2(g->rA, i, j) * 1. * 1. * g->drF[k - 1] * G3(g->hFacC, i, j, k) + (1. - G3(g->maskC, i, j, k));
      maskLocW[S(i, j)] = G3(g->maskW, i, j, k);
      maskLocS[S(i, j)] = G3(g->maskS, i, j, k);
    }
    if (cube) gad_fill_cs_corner_uv(d, corners, maskLocW, maskLocS);       /* :329-334 */
    for (int ipass = 1; ipass <= npass; ipass++) {
      int interiorOnly = 0, overlapOnly = 0, fluxX, fluxY;
      if (cube) {                                                          /* :346-362 */
        if (ipass == 1) {
          overlapOnly = nCFace % 3 == 0; interiorOnly = nCFace % 3 != 0;
          fluxX = nCFace == 6 || nCFace == 1 || nCFace == 2; fluxY = nCFace == 3 || nCFace == 4 || nCFace == 5;
        } else if (ipass == 2) {
          overlapOnly = nCFace % 3 == 2; interiorOnly = nCFace % 3 == 1;
          fluxX = nCFace == 2 || nCFace == 3 || nCFace == 4; fluxY = nCFace == 5 || nCFace == 6 || nCFace == 1;
        } else {
          interiorOnly = 1;
          fluxX = nCFace == 5 || nCFace == 6; fluxY = nCFace == 2 || nCFace == 3;
        }
      } else { fluxX = ipass % 2 == 1; fluxY = !fluxX; }
      const double dT = deltaTLev[k - 1];
#define MD_UPDATE(i, j, trA, trB, afA, afB)                                                                          \
      do {                                                                                                           \
        if (compressible) {                                                                                          \
          const double tmpTrac = localTij[S(i, j)] * localVol[S(i, j)] - dT * ((afB) - (afA)) * 1.;                  \
          localVol[S(i, j)] = localVol[S(i, j)] - dT * ((trB) - (trA)) * 1.;                                         \
          localTij[S(i, j)] = tmpTrac / localVol[S(i, j)];                                                           \
        } else {                                                                                                     \
          localTij[S(i, j)] = localTij[S(i, j)]                                                                      \
              - dT * 1. * G3(g->recip_hFacC, i, j, k) * g->recip_drF[k - 1] * G2(g->recip_rA, i, j) * 1.            \
                    * ((afB) - (afA) - G3(tracer, i, j, k) * ((trB) - (trA))) * 1.;                                  \
        }                                                                                                            \
      } while (0)
      FORALL af[S(i, j)] = 0.;
      if (fluxX) {
        if (!overlapOnly || N_edge || S_edge) {
          if (overlapOnly) gad_fill_cs_corner_tr(d, 1, corners, localTij);
          FORALL af[S(i, j)] = 0.;
          adv_h(g, bi, bj, k, 0, advectionScheme, dT, uTrans, uK, maskLocW, localTij, af);
          if (overlapOnly && ipass == 1) gad_fill_cs_corner_tr(d, 2, corners, localTij);
        }
        if (overlapOnly) {                                                 /* :471-546 */
          const int iMinUpd = W_edge ? 1 : 1 - OLx + 1, iMaxUpd = E_edge ? sNx : sNx + OLx - 1;
          if (S_edge)
            for (int j = 1 - OLy; j <= 0; j++)
              for (int i = iMinUpd; i <= iMaxUpd; i++) MD_UPDATE(i, j, uTrans[S(i, j)], uTrans[S(i + 1, j)], af[S(i, j)], af[S(i + 1, j)]);
          if (N_edge)
            for (int j = sNy + 1; j <= sNy + OLy; j++)
              for (int i = iMinUpd; i <= iMaxUpd; i++) MD_UPDATE(i, j, uTrans[S(i, j)], uTrans[S(i + 1, j)], af[S(i, j)], af[S(i + 1, j)]);
        } else {                                                           /* :547-600 */
          const int jMinUpd = (interiorOnly && S_edge) ? 1 : 1 - OLy, jMaxUpd = (interiorOnly && N_edge) ? sNy : sNy + OLy;
          for (int j = jMinUpd; j <= jMaxUpd; j++)
            for (int i = 1 - OLx + 1; i <= sNx + OLx - 1; i++) MD_UPDATE(i, j, uTrans[S(i, j)], uTrans[S(i + 1, j)], af[S(i, j)], af[S(i + 1, j)]);
        }
      }
      FORALL af[S(i, j)] = 0.;
      if (fluxY) {
        if (!overlapOnly || E_edge || W_edge) {
          if (overlapOnly) gad_fill_cs_corner_tr(d, 2, corners, localTij);
          FORALL af[S(i, j)] = 0.;
          adv_h(g, bi, bj, k, 1, advectionScheme, dT, vTrans, vK, maskLocS, localTij, af);
          if (overlapOnly && ipass == 1) gad_fill_cs_corner_tr(d, 1, corners, localTij);
        }
        if (overlapOnly) {                                                 /* :692-768 */
          const int jMinUpd = S_edge ? 1 : 1 - OLy + 1, jMaxUpd = N_edge ? sNy : sNy + OLy - 1;
          if (W_edge)
            for (int j = jMinUpd; j <= jMaxUpd; j++)
              for (int i = 1 - OLx; i <= 0; i++) MD_UPDATE(i, j, vTrans[S(i, j)], vTrans[S(i, j + 1)], af[S(i, j)], af[S(i, j + 1)]);
          if (E_edge)
            for (int j = jMinUpd; j <= jMaxUpd; j++)
              for (int i = sNx + 1; i <= sNx + OLx; i++) MD_UPDATE(i, j, vTrans[S(i, j)], vTrans[S(i, j + 1)], af[S(i, j)], af[S(i, j + 1)]);
        } else {                                                           /* :769-812 */
          const int iMinUpd = (interiorOnly && W_edge) ? 1 : 1 - OLx, iMaxUpd = (interiorOnly && E_edge) ? sNx : sNx + OLx;
          for (int j = 1 - OLy + 1; j <= sNy + OLy - 1; j++)
            for (int i = iMinUpd; i <= iMaxUpd; i++) MD_UPDATE(i, j, vTrans[S(i, j)], vTrans[S(i, j + 1)], af[S(i, j)], af[S(i, j + 1)]);
        }
      }
#undef MD_UPDATE
    }
    if (implicitAdvection) {
      FORALL K3(gTracer, i, j, k) = (localTij[S(i, j)] - G3(tracer, i, j, k)) / deltaTLev[k - 1];
    } else {
      FORALL { K3(locVol3d, i, j, k) = localVol[S(i, j)]; K3(localT3d, i, j, k) = localTij[S(i, j)]; }
    }
  }
  if (!implicitAdvection) {
    FORALL { fVerT[S(i, j)] = 0.; fVerT[ns + S(i, j)] = 0.; rTrans[S(i, j)] = 0.; }
    for (int k = Nr; k >= 1; k--) {
      const int kUp = 1 + (k + 1) % 2, kDown = 1 + k % 2;
      const double kp1Msk = k == Nr ? 0. : 1.;
      double *fUp = fVerT + ns * (size_t)(kUp - 1), *fDn = fVerT + ns * (size_t)(kDown - 1);
      if (k == 1) {
        FORALL { rTransKp[S(i, j)] = kp1Msk * rTrans[S(i, j)]; rTrans[S(i, j)] = 0.; fUp[S(i, j)] = 0.; }
      } else {
        FORALL {
          rTransKp[S(i, j)] = kp1Msk * rTrans[S(i, j)];
          rTrans[S(i, j)] = G3(wFld, i, j, k) * G2(g->rA, i, j) * 1. * 1. * G3(g->maskC, i, j, k - 1);
          fUp[S(i, j)] = 0.;
        }
        /* GAD_DST2U1_ADV_R is called with advectionScheme, not vertAdvecScheme (gad_advection.F:976) */
        const int vs = (vertAdvecScheme == UPWIND_1RST || vertAdvecScheme == DST2) ? advectionScheme : vertAdvecScheme;
        adv_r(g, p, bi, bj, k, vs, deltaTLev[k - 1], rTrans, &G3(wFld, 1 - OLx, 1 - OLy, k), localT3d, fUp);
      }
      FORALL {
        if (compressible) {
          const double tmpTrac = K3(localT3d, i, j, k) * K3(locVol3d, i, j, k)
                                 - deltaTLev[k - 1] * (fDn[S(i, j)] - fUp[S(i, j)]) * p->rkSign * 1.;
          localVol[S(i, j)] = K3(locVol3d, i, j, k) - deltaTLev[k - 1] * (rTransKp[S(i, j)] - rTrans[S(i, j)]) * p->rkSign * 1.;
          K3(gTracer, i, j, k) = (tmpTrac - G3(tracer, i, j, k) * localVol[S(i, j)])
                                 * G2(g->recip_rA, i, j) * 1. * g->recip_drF[k - 1] * G3(g->recip_hFacC, i, j, k) * 1. / deltaTLev[k - 1];
        } else {
          const double lt = K3(localT3d, i, j, k)
              - deltaTLev[k - 1] * 1. * G3(g->recip_hFacC, i, j, k) * g->recip_drF[k - 1] * G2(g->recip_rA, i, j) * 1.
                    * (fDn[S(i, j)] - fUp[S(i, j)] - G3(tracer, i, j, k) * (rTransKp[S(i, j)] - rTrans[S(i, j)])) * p->rkSign * 1.;
          K3(gTracer, i, j, k) = (lt - G3(tracer, i, j, k)) / deltaTLev[k - 1];
        }
      }
    }
  }
  free(buf);
  free(fVerT);
  return 0;
}
