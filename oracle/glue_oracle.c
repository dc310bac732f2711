/* glue_oracle.c -- CPU restatement of the step glue either side of the hot path
 * (SURVEY.md §8(f) rank 1): TIMESTEP + ADAMS_BASHFORTH2 + APPLY_FORCING_U/V,
 * CALC_DIV_GHAT + the SOLVE_FOR_PRESSURE right-hand side, CALC_GRAD_PHI_SURF +
 * CORRECTION_STEP, INTEGRATE_FOR_W, TIMESTEP_TRACER, CALC_PHI_HYD (linear EOS,
 * z coordinates, finite-volume form).  TEST INFRASTRUCTURE ONLY.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "mitgcm_oracle.h"
#include "glue_oracle.h"

#define S(i, j) ((size_t)((i) + OLx - 1) + (size_t)px * (size_t)((j) + OLy - 1))
#define G2(a, i, j) (a)[S(i, j) + off2]
#define G3(a, i, j, k) (a)[S(i, j) + (size_t)px * py * (size_t)((k)-1) + off3]
#define FORALL for (int j = 1 - OLy; j <= sNy + OLy; j++) for (int i = 1 - OLx; i <= sNx + OLx; i++)
#define SETUP                                                                          \
  const og_dims *d = &g->d;                                                            \
  const int sNx = d->sNx, sNy = d->sNy, OLx = d->OLx, OLy = d->OLy, Nr = d->Nr;        \
  const size_t px = (size_t)(sNx + 2 * OLx), py = (size_t)(sNy + 2 * OLy);             \
  const size_t tile = (size_t)(bi - 1) + (size_t)d->nSx * (size_t)(bj - 1);            \
  const size_t off2 = px * py * tile, off3 = px * py * (size_t)Nr * tile;              \
  (void)Nr; (void)off2; (void)off3;

/* TIMESTEP for one tile, one level: model/src/timestep.F:95-385 with
 * staggerTimeStep = F, implicSurfPress = 1 (gUdPx = 0), no CD scheme, linear
 * free surface; momForcing = surface stress at k = 1 (apply_forcing.F:142-148);
 * ADAMS_BASHFORTH2 (adams_bashforth2.F:62-86). abFac = 0 on the first step. */
void og_timestep(const og_grid *g, const og_params *p, int bi, int bj, int k,
                 int iMin, int iMax, int jMin, int jMax,
                 const double *dPhiHydX, const double *dPhiHydY,
                 const double *guDissip, const double *gvDissip,
                 const double *surfaceForcingU, const double *surfaceForcingV,
                 int momForcing, int momDissip_In_AB, double abFac,
                 const double *uVel, const double *vVel,
                 double *gU, double *gV, double *guNm1, double *gvNm1,
                 const double *phiSurfX, const double *phiSurfY) {
  SETUP
  const size_t ns = px * py;
  double *guExt = (double *)calloc(ns * 4, sizeof(double));
  double *gvExt = guExt + ns, *gUtmp = guExt + 2 * ns, *gVtmp = guExt + 3 * ns;
  const double phFac = 1.; /* pfFacMom */
  if (momForcing && k == 1) {
    for (int j = 0; j <= sNy + 1; j++)
      for (int i = 1; i <= sNx + 1; i++)
        guExt[S(i, j)] = guExt[S(i, j)] + G2(surfaceForcingU, i, j) * g->recip_drF[k - 1] * G3(g->recip_hFacW, i, j, k);
    for (int j = 1; j <= sNy + 1; j++)
      for (int i = 0; i <= sNx + 1; i++)
        gvExt[S(i, j)] = gvExt[S(i, j)] + G2(surfaceForcingV, i, j) * g->recip_drF[k - 1] * G3(g->recip_hFacS, i, j, k);
  }
  for (int j = jMin; j <= jMax; j++)
    for (int i = iMin; i <= iMax; i++) {
      G3(gU, i, j, k) = G3(gU, i, j, k) - phFac * dPhiHydX[S(i, j)];
      G3(gV, i, j, k) = G3(gV, i, j, k) - phFac * dPhiHydY[S(i, j)];
    }
  if (p->momViscosity && momDissip_In_AB)
    for (int j = jMin; j <= jMax; j++)
      for (int i = iMin; i <= iMax; i++) {
        G3(gU, i, j, k) = G3(gU, i, j, k) + guDissip[S(i, j)];
        G3(gV, i, j, k) = G3(gV, i, j, k) + gvDissip[S(i, j)];
      }
  if (momForcing)
    for (int j = jMin; j <= jMax; j++)
      for (int i = iMin; i <= iMax; i++) {
        G3(gU, i, j, k) = G3(gU, i, j, k) + guExt[S(i, j)];
        G3(gV, i, j, k) = G3(gV, i, j, k) + gvExt[S(i, j)];
      }
  /* ADAMS_BASHFORTH2 over the whole slab */
  FORALL {
    double ab = abFac * (G3(gU, i, j, k) - G3(guNm1, i, j, k));
    G3(guNm1, i, j, k) = G3(gU, i, j, k);
    G3(gU, i, j, k) = G3(gU, i, j, k) + ab;
    ab = abFac * (G3(gV, i, j, k) - G3(gvNm1, i, j, k));
    G3(gvNm1, i, j, k) = G3(gV, i, j, k);
    G3(gV, i, j, k) = G3(gV, i, j, k) + ab;
  }
  for (int j = jMin; j <= jMax; j++)
    for (int i = iMin; i <= iMax; i++) {
      gUtmp[S(i, j)] = G3(gU, i, j, k);
      gVtmp[S(i, j)] = G3(gV, i, j, k);
    }
  if (p->momViscosity && !momDissip_In_AB)
    for (int j = jMin; j <= jMax; j++)
      for (int i = iMin; i <= iMax; i++) {
        gUtmp[S(i, j)] = gUtmp[S(i, j)] + guDissip[S(i, j)];
        gVtmp[S(i, j)] = gVtmp[S(i, j)] + gvDissip[S(i, j)];
      }
  for (int j = jMin; j <= jMax; j++)
    for (int i = iMin; i <= iMax; i++) {
      /* gUdPx = -psFac*phiSurfX with psFac = pfFacMom*(1 - implicSurfPress) (timestep.F:95-97, 230-236);
       * phiSurf is only computed by the caller when implicSurfPress != 1 (dynamics.F:249-255) */
      const double psFac = 1. * (1. - p->implicSurfPress) * 1. * 1.;
      const double gUdPx = phiSurfX ? -psFac * phiSurfX[S(i, j)] : 0.;
      const double gVdPy = phiSurfY ? -psFac * phiSurfY[S(i, j)] : 0.;
      G3(gU, i, j, k) = G3(uVel, i, j, k) + p->deltaTMom * (gUtmp[S(i, j)] + gUdPx) * G3(g->maskW, i, j, k);
      G3(gV, i, j, k) = G3(vVel, i, j, k) + p->deltaTMom * (gVtmp[S(i, j)] + gVdPy) * G3(g->maskS, i, j, k);
    }
  free(guExt);
}

/* Right-hand side of the surface-pressure equation for one tile:
 * solve_for_pressure.F:120-238 (cg2d_x = Bo_surf*etaN, cg2d_b = 0, CALC_DIV_GHAT
 * for k = Nr..1 with implicDiv2DFlow = 1 (calc_div_ghat.F:64-167), free-surface
 * term on etaFS = etaH (exactConserv, :213-222) or etaN (:224-233)). */
void og_solve_rhs(const og_grid *g, const og_params *p, int bi, int bj, const double *Bo_surf,
                  const double *etaN, const double *etaFS, const double *gU, const double *gV,
                  double *cg2d_b, double *cg2d_x) {
  SETUP
  const size_t ns = px * py;
  double *xA = (double *)calloc(ns * 3, sizeof(double));
  double *yA = xA + ns, *pf = xA + 2 * ns;
  FORALL {
    G2(cg2d_x, i, j) = G2(Bo_surf, i, j) * G2(etaN, i, j);
    G2(cg2d_b, i, j) = 0.;
  }
  for (int k = Nr; k >= 1; k--) {
    for (int j = 1; j <= sNy + 1; j++)
      for (int i = 1; i <= sNx + 1; i++) {
        xA[S(i, j)] = G2(g->dyG, i, j) * g->drF[k - 1] * G3(g->hFacW, i, j, k);
        yA[S(i, j)] = G2(g->dxG, i, j) * g->drF[k - 1] * G3(g->hFacS, i, j, k);
      }
    /* calc_div_ghat.F:75-98: implicDiv2DFlow = 1, or < 1 with exactConserv (the (1-implicDiv2DFlow)*uVel part
     * then lives in etaH); the third branch (no exactConserv) is not restated */
    const int full = p->implicDiv2DFlow == 1.;
    for (int j = 1; j <= sNy; j++)
      for (int i = 1; i <= sNx + 1; i++)
        pf[S(i, j)] = full ? xA[S(i, j)] * G3(gU, i, j, k) / p->deltaTMom
                           : p->implicDiv2DFlow * xA[S(i, j)] * G3(gU, i, j, k) / p->deltaTMom;
    for (int j = 1; j <= sNy; j++)
      for (int i = 1; i <= sNx; i++) G2(cg2d_b, i, j) = G2(cg2d_b, i, j) + pf[S(i + 1, j)] - pf[S(i, j)];
    for (int j = 1; j <= sNy + 1; j++)
      for (int i = 1; i <= sNx; i++)
        pf[S(i, j)] = full ? yA[S(i, j)] * G3(gV, i, j, k) / p->deltaTMom
                           : p->implicDiv2DFlow * yA[S(i, j)] * G3(gV, i, j, k) / p->deltaTMom;
    for (int j = 1; j <= sNy; j++)
      for (int i = 1; i <= sNx; i++) G2(cg2d_b, i, j) = G2(cg2d_b, i, j) + pf[S(i, j + 1)] - pf[S(i, j)];
  }
  for (int j = 1; j <= sNy; j++)
    for (int i = 1; i <= sNx; i++)
      G2(cg2d_b, i, j) = G2(cg2d_b, i, j)
          - p->freeSurfFac * G2(g->rA, i, j) * 1. / p->deltaTMom / p->deltaTFreeSurf * G2(etaFS, i, j);
  free(xA);
}

/* MOMENTUM_CORRECTION_STEP for one tile: CALC_GRAD_PHI_SURF (calc_grad_phi_surf.F)
 * on 2-OL..sN+OL and CORRECTION_STEP (correction_step.F:152-231). */
void og_correction_step(const og_grid *g, const og_params *p, int bi, int bj, const double *Bo_surf,
                        const double *etaN, const double *gU, const double *gV,
                        double *uVel, double *vVel) {
  SETUP
  const size_t ns = px * py;
  double *phiSurfX = (double *)calloc(ns * 2, sizeof(double));
  double *phiSurfY = phiSurfX + ns;
  const int iMin = 2 - OLx, iMax = sNx + OLx, jMin = 2 - OLy, jMax = sNy + OLy;
  for (int j = jMin; j <= jMax; j++)
    for (int i = iMin; i <= iMax; i++) {
      phiSurfX[S(i, j)] = G2(g->recip_dxC, i, j)
          * (G2(Bo_surf, i, j) * G2(etaN, i, j) - G2(Bo_surf, i - 1, j) * G2(etaN, i - 1, j));
      phiSurfY[S(i, j)] = G2(g->recip_dyC, i, j)
          * (G2(Bo_surf, i, j) * G2(etaN, i, j) - G2(Bo_surf, i, j - 1) * G2(etaN, i, j - 1));
    }
  const double psFac = 1. * p->implicSurfPress;
  for (int k = 1; k <= Nr; k++)
    for (int j = jMin; j <= jMax; j++)
      for (int i = iMin; i <= iMax; i++) {
        double gU_dpx = -psFac * phiSurfX[S(i, j)] * G3(g->maskW, i, j, k);
        double gV_dpy = -psFac * phiSurfY[S(i, j)] * G3(g->maskS, i, j, k);
        G3(uVel, i, j, k) = (G3(gU, i, j, k) + p->deltaTMom * gU_dpx) * G3(g->maskW, i, j, k);
        G3(vVel, i, j, k) = (G3(gV, i, j, k) + p->deltaTMom * gV_dpy) * G3(g->maskS, i, j, k);
      }
  free(phiSurfX);
}

/* INTEGRATE_FOR_W for one tile, k = Nr..1 (integrate_for_w.F, free-surface,
 * r coordinate branch). */
void og_integrate_for_w(const og_grid *g, const og_params *p, int bi, int bj,
                        const double *uVel, const double *vVel, double *wVel) {
  SETUP
  const size_t ns = px * py;
  double *uTrans = (double *)calloc(ns * 2, sizeof(double));
  double *vTrans = uTrans + ns;
  for (int k = Nr; k >= 1; k--) {
    for (int j = 1; j <= sNy + 1; j++)
      for (int i = 1; i <= sNx + 1; i++) {
        uTrans[S(i, j)] = G3(uVel, i, j, k) * G2(g->dyG, i, j) * g->drF[k - 1] * G3(g->hFacW, i, j, k);
        vTrans[S(i, j)] = G3(vVel, i, j, k) * G2(g->dxG, i, j) * g->drF[k - 1] * G3(g->hFacS, i, j, k);
      }
    for (int j = 1; j <= sNy; j++)
      for (int i = 1; i <= sNx; i++) {
        double conv2d = -(uTrans[S(i + 1, j)] - uTrans[S(i, j)] + vTrans[S(i, j + 1)] - vTrans[S(i, j)]);
        if (p->rigidLid) {
          if (k == 1) G3(wVel, i, j, k) = 0.;
          else if (k == Nr)
            G3(wVel, i, j, k) = conv2d * G2(g->recip_rA, i, j) * G3(g->maskC, i, j, k) * G3(g->maskC, i, j, k - 1);
          else
            G3(wVel, i, j, k) = (G3(wVel, i, j, k + 1) + conv2d * G2(g->recip_rA, i, j))
                                * G3(g->maskC, i, j, k) * G3(g->maskC, i, j, k - 1);
        } else {
          if (k == Nr) G3(wVel, i, j, k) = conv2d * G2(g->recip_rA, i, j) * G3(g->maskC, i, j, k);
          else G3(wVel, i, j, k) = (G3(wVel, i, j, k + 1) + conv2d * G2(g->recip_rA, i, j)) * G3(g->maskC, i, j, k);
        }
      }
  }
  free(uTrans);
}

/* CALC_GRAD_PHI_SURF (model/src/calc_grad_phi_surf.F) for one tile on iMin..iMax x jMin..jMax. */
void og_calc_grad_phi_surf(const og_grid *g, int bi, int bj, int iMin, int iMax, int jMin, int jMax,
                           const double *Bo_surf, const double *etaFld, double *phiSurfX, double *phiSurfY) {
  SETUP
  for (int j = jMin; j <= jMax; j++)
    for (int i = iMin; i <= iMax; i++) {
      phiSurfX[S(i, j)] = G2(g->recip_dxC, i, j)
          * (G2(Bo_surf, i, j) * G2(etaFld, i, j) - G2(Bo_surf, i - 1, j) * G2(etaFld, i - 1, j));
      phiSurfY[S(i, j)] = G2(g->recip_dyC, i, j)
          * (G2(Bo_surf, i, j) * G2(etaFld, i, j) - G2(Bo_surf, i, j - 1) * G2(etaFld, i, j - 1));
    }
}
