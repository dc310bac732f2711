/* nh_oracle.c -- CPU restatement of the non-hydrostatic step around CG3D, so that the
 * CG3D solver lines of verification/tutorial_deep_convection (cg3d: Sum(rhs),rhsMax / cg3d_init_res / cg3d_last_res)
 * can be reproduced and the CG3D restatement (cg3d_oracle.c) and kernel (csrc/cg3d.cu) PINNED:
 *   CALC_GW          model/src/calc_gw.F:156-640   vertical momentum tendency (flux-form advection, harmonic
 *                                                  viscosity, free-slip sides) + ADAMS_BASHFORTH2 on gW
 *   TIMESTEP_WVEL    model/src/timestep_wvel.F:60-118
 *   the NH right-hand sides: CALC_DIV_GHAT with cg3d_b (calc_div_ghat.F:60-150), the old-style free-surface term
 *                    of SOLVE_FOR_PRESSURE (solve_for_pressure.F:194-218, use3Dsolver without exactConserv) and
 *                    PRE_CG3D (pre_cg3d.F:60-250, uniformFreeSurfLev branch)
 *   CORRECTION_STEP  with the gradient of phi_nh (correction_step.F:152-200)
 * TEST INFRASTRUCTURE ONLY.  z coordinates, Boussinesq (deepFac* = rhoFac* = wUnit2rVel = rVel2wUnit = 1: omitted,
 * multiplying by 1.0 is exact), selectNHfreeSurf = 0, implicitNHPress = 1, nh_Am2 = 1, no OBCS, no SMAG_3D,
 * no biharmonic viscosity in W (viscA4W = 0), free-slip sides (MOM_W_SIDEDRAG not restated: returns 1),
 * useNHMTerms = F, select3dCoriScheme = 0 -- what the experiment runs. */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "mitgcm_oracle.h"
#include "glue_oracle.h"

#define S(i, j) ((size_t)((i) + OLx - 1) + (size_t)px * (size_t)((j) + OLy - 1))
#define G2(a, i, j) (a)[S(i, j) + off2]
#define G3(a, i, j, k) (a)[S(i, j) + (size_t)px * py * (size_t)((k)-1) + off3]
#define K3(a, i, j, k) (a)[S(i, j) + (size_t)px * py * (size_t)((k)-1)] /* per-tile (Nr+1)-level array */
#define FORALL for (int j = 1 - OLy; j <= sNy + OLy; j++) for (int i = 1 - OLx; i <= sNx + OLx; i++)
#define SETUP                                                                          \
  const og_dims *d = &g->d;                                                            \
  const int sNx = d->sNx, sNy = d->sNy, OLx = d->OLx, OLy = d->OLy, Nr = d->Nr;        \
  const size_t px = (size_t)(sNx + 2 * OLx), py = (size_t)(sNy + 2 * OLy);             \
  const size_t tile = (size_t)(bi - 1) + (size_t)d->nSx * (size_t)(bj - 1);            \
  const size_t off2 = px * py * tile, off3 = px * py * (size_t)Nr * tile;              \
  (void)Nr; (void)off2; (void)off3;

static double dmin(double a, double b) { return a < b ? a : b; }
static double dmax(double a, double b) { return a > b ? a : b; }

/* CALC_GW for one tile (calc_gw.F:156-640, CALC_GW_NEW_THICK).  R_low, Ro_surf, rLowW, rSurfW, rLowS, rSurfS:
 * tile2d (GRID.h, ini_depths.F / ini_masks_etc.F); rC: Nr cell-centre positions; kappaRU/V: per-tile (Nr+1)
 * levels; viscAhW / viscA4W: the constants viscAh_W / viscA4_W hold (no variable viscosity).
 * abFac: ADAMS_BASHFORTH2 factor (0 on the first step when nHydStartAB = 0, else 0.5 + abEps). */
int og_calc_gw(const og_grid *g, const og_params *p, int bi, int bj, const double *R_low, const double *Ro_surf,
               const double *rLowW, const double *rSurfW, const double *rLowS, const double *rSurfS,
               const double *rC, const double *kappaRU, const double *kappaRV, double viscAhW, double viscA4W,
               int momDissip_In_AB, double abFac, const double *uVel, const double *vVel, const double *wVel,
               double *gW, double *gwNm1) {
  SETUP
  if (p->momViscosity && p->no_slip_sides) return 1;       /* MOM_W_SIDEDRAG */
  if (viscA4W != 0.) return 1;                              /* biharmonic block (calc_gw.F:288-345) */
  const size_t ns = px * py;
  double *buf = (double *)calloc(ns * 13, sizeof(double));
  double *xA = buf, *yA = buf + ns, *rThickC_W = buf + 2 * ns, *rThickC_S = buf + 3 * ns, *rThickC_C = buf + 4 * ns;
  double *recip_rThickC = buf + 5 * ns, *flx_NS = buf + 6 * ns, *flx_EW = buf + 7 * ns, *flx_Dn = buf + 8 * ns;
  double *flxAdvUp = buf + 9 * ns, *flxDisUp = buf + 10 * ns, *gwDiss = buf + 11 * ns, *del2w = buf + 12 * ns;
  const int iMin = 1, iMax = sNx, jMin = 1, jMax = sNy;
  const double halfRL = 0.5, rkSign = p->rkSign;
  for (int k = 1; k <= Nr; k++) FORALL G3(gW, i, j, k) = 0.;
  for (int k = 1; k <= Nr; k++) {
    const int km1 = k - 1 > 1 ? k - 1 : 1, kp1 = k + 1 < Nr ? k + 1 : Nr;
    const double mskM1 = k == 1 ? 0. : 1., mskP1 = k == Nr ? 0. : 1.;
    if (k > 1) {
      FORALL {
        if (G3(g->maskC, i, j, k - 1) == 0. || G3(g->maskC, i, j, k) == 0.) recip_rThickC[S(i, j)] = 0.;
        else recip_rThickC[S(i, j)] = 1. / (dmin(G2(Ro_surf, i, j), rC[k - 2]) - dmax(G2(R_low, i, j), rC[k - 1]));
      }
      if (p->momViscosity) {
        FORALL rThickC_C[S(i, j)] = dmax(0., dmin(G2(Ro_surf, i, j), rC[k - 2]) - dmax(G2(R_low, i, j), rC[k - 1]));
        for (int j = 1 - OLy; j <= sNy + OLy; j++)
          for (int i = 2 - OLx; i <= sNx + OLx; i++) {
            rThickC_W[S(i, j)] = dmax(0., dmin(G2(rSurfW, i, j), rC[k - 2]) - dmax(G2(rLowW, i, j), rC[k - 1]));
            xA[S(i, j)] = G2(g->dyG, i, j) * rThickC_W[S(i, j)];
          }
        for (int j = 2 - OLy; j <= sNy + OLy; j++)
          for (int i = 1 - OLx; i <= sNx + OLx; i++) {
            rThickC_S[S(i, j)] = dmax(0., dmin(G2(rSurfS, i, j), rC[k - 2]) - dmax(G2(rLowS, i, j), rC[k - 1]));
            yA[S(i, j)] = G2(g->dxG, i, j) * rThickC_S[S(i, j)];
          }
      }
    }
    /* viscous fluxes and dissipation tendency (calc_gw.F:347-435) */
    if (p->momViscosity && k > 1) {
      for (int j = jMin; j <= jMax; j++)
        for (int i = iMin; i <= iMax + 1; i++)
          flx_EW[S(i, j)] =
              -(viscAhW + viscAhW) * halfRL * (G3(wVel, i, j, k) - G3(wVel, i - 1, j, k)) * G2(g->recip_dxC, i, j) * xA[S(i, j)]
                  * g->cosFacU[(size_t)(j + OLy - 1) + py * tile]
              + (viscA4W + viscA4W) * halfRL * (del2w[S(i, j)] - del2w[S(i - 1, j)]) * G2(g->recip_dxC, i, j) * xA[S(i, j)]
                    * g->cosFacU[(size_t)(j + OLy - 1) + py * tile];
      for (int j = jMin; j <= jMax + 1; j++)
        for (int i = iMin; i <= iMax; i++)
          flx_NS[S(i, j)] =
              -(viscAhW + viscAhW) * halfRL * (G3(wVel, i, j, k) - G3(wVel, i, j - 1, k)) * G2(g->recip_dyC, i, j) * yA[S(i, j)]
              + (viscA4W + viscA4W) * halfRL * (del2w[S(i, j)] - del2w[S(i, j - 1)]) * G2(g->recip_dyC, i, j) * yA[S(i, j)];
      for (int j = jMin; j <= jMax; j++)
        for (int i = iMin; i <= iMax; i++) {
          const double viscLoc = (K3(kappaRU, i, j, k) + K3(kappaRU, i + 1, j, k) + K3(kappaRU, i, j, k + 1) + K3(kappaRU, i + 1, j, k + 1)
                                  + K3(kappaRV, i, j, k) + K3(kappaRV, i, j + 1, k) + K3(kappaRV, i, j, k + 1) + K3(kappaRV, i, j + 1, k + 1))
                                 * 0.125;
          flx_Dn[S(i, j)] = -viscLoc * (G3(wVel, i, j, kp1) * mskP1 - G3(wVel, i, j, k)) * rkSign * g->recip_drF[k - 1] * G2(g->rA, i, j);
        }
      if (k == 2)
        for (int j = jMin; j <= jMax; j++)
          for (int i = iMin; i <= iMax; i++) {
            const double viscLoc = (K3(kappaRU, i, j, k) + K3(kappaRU, i + 1, j, k) + K3(kappaRV, i, j, k) + K3(kappaRV, i, j + 1, k)) * 0.25;
            flxDisUp[S(i, j)] = -viscLoc * (G3(wVel, i, j, k) - G3(wVel, i, j, k - 1)) * rkSign * g->recip_drF[k - 2] * G2(g->rA, i, j);
          }
      for (int j = jMin; j <= jMax; j++)
        for (int i = iMin; i <= iMax; i++) {
          gwDiss[S(i, j)] = -((flx_EW[S(i + 1, j)] - flx_EW[S(i, j)]) + (flx_NS[S(i, j + 1)] - flx_NS[S(i, j)])
                              + (flx_Dn[S(i, j)] - flxDisUp[S(i, j)]) * rkSign)
                            * G2(g->recip_rA, i, j) * recip_rThickC[S(i, j)];
          flxDisUp[S(i, j)] = flx_Dn[S(i, j)];
        }
    }
    /* advective fluxes and tendency (calc_gw.F:470-560) */
    if (p->momAdvection) {
      if (k > 1) {
        for (int j = jMin; j <= jMax; j++)
          for (int i = iMin; i <= iMax + 1; i++) {
            const double uTrans = (g->drF[km1 - 1] * G3(g->hFacW, i, j, km1) * G3(uVel, i, j, km1) * mskM1
                                   + g->drF[k - 1] * G3(g->hFacW, i, j, k) * G3(uVel, i, j, k))
                                  * halfRL * G2(g->dyG, i, j);
            flx_EW[S(i, j)] = uTrans * (G3(wVel, i, j, k) + G3(wVel, i - 1, j, k)) * halfRL;
          }
        for (int j = jMin; j <= jMax + 1; j++)
          for (int i = iMin; i <= iMax; i++) {
            const double vTrans = (g->drF[km1 - 1] * G3(g->hFacS, i, j, km1) * G3(vVel, i, j, km1) * mskM1
                                   + g->drF[k - 1] * G3(g->hFacS, i, j, k) * G3(vVel, i, j, k))
                                  * halfRL * G2(g->dxG, i, j);
            flx_NS[S(i, j)] = vTrans * (G3(wVel, i, j, k) + G3(wVel, i, j - 1, k)) * halfRL;
          }
      }
      for (int j = jMin; j <= jMax; j++)
        for (int i = iMin; i <= iMax; i++) {
          const double tmp_WbarZ = halfRL * (G3(wVel, i, j, k) + G3(wVel, i, j, kp1) * mskP1);
          const double rTrans = halfRL * (G3(wVel, i, j, k) + G3(wVel, i, j, kp1) * mskP1) * G2(g->rA, i, j);
          flx_Dn[S(i, j)] = rTrans * tmp_WbarZ;
        }
      if (k > 1)
        for (int j = jMin; j <= jMax; j++)
          for (int i = iMin; i <= iMax; i++)
            G3(gW, i, j, k) = -((flx_EW[S(i + 1, j)] - flx_EW[S(i, j)]) + (flx_NS[S(i, j + 1)] - flx_NS[S(i, j)])
                                + (flx_Dn[S(i, j)] - flxAdvUp[S(i, j)]) * rkSign)
                              * G2(g->recip_rA, i, j) * recip_rThickC[S(i, j)];
      for (int j = jMin; j <= jMax; j++)
        for (int i = iMin; i <= iMax; i++) flxAdvUp[S(i, j)] = flx_Dn[S(i, j)];
    }
    if (p->momViscosity && momDissip_In_AB)
      for (int j = jMin; j <= jMax; j++)
        for (int i = iMin; i <= iMax; i++) G3(gW, i, j, k) = G3(gW, i, j, k) + gwDiss[S(i, j)];
    /* ADAMS_BASHFORTH2, tendency form (adams_bashforth2.F:76-84), whole slab */
    FORALL {
      const double ab = abFac * (G3(gW, i, j, k) - G3(gwNm1, i, j, k));
      G3(gwNm1, i, j, k) = G3(gW, i, j, k);
      G3(gW, i, j, k) = G3(gW, i, j, k) + ab;
    }
    if (p->momViscosity && !momDissip_In_AB)
      for (int j = jMin; j <= jMax; j++)
        for (int i = iMin; i <= iMax; i++) G3(gW, i, j, k) = G3(gW, i, j, k) + gwDiss[S(i, j)];
  }
  free(buf);
  return 0;
}

/* TIMESTEP_WVEL for one tile (timestep_wvel.F:60-118): nonHydrostatic, implicitNHPress = 1, nh_Am2 = 1, no
 * implicitIntGravWave.  gW returns the OLD wVel, wVel the stepped one (interior). */
void og_timestep_wvel(const og_grid *g, const og_params *p, int bi, int bj, double *gW, double *wVel) {
  SETUP
  const size_t ns = px * py;
  double *gWtmp = (double *)calloc(ns, sizeof(double));
  const double nh_Fac = 1. / 1.;
  for (int k = 1; k <= Nr; k++) {
    const int km1 = k - 1 > 1 ? k - 1 : 1;
    FORALL {
      gWtmp[S(i, j)] = G3(gW, i, j, k) * G3(g->maskC, i, j, k) * G3(g->maskC, i, j, km1);
      G3(gW, i, j, k) = G3(wVel, i, j, k);
    }
    double tmpFac = nh_Fac + 0.;
    if (tmpFac > 0.) tmpFac = 1. / tmpFac;
    for (int j = 1; j <= sNy; j++)
      for (int i = 1; i <= sNx; i++) G3(wVel, i, j, k) = G3(wVel, i, j, k) + p->deltaTMom * tmpFac * gWtmp[S(i, j)];
  }
  free(gWtmp);
}

/* The two right-hand sides before the solvers, one tile (solve_for_pressure.F:120-218 with use3Dsolver and not
 * exactConserv = oldFreeSurfTerm; CALC_DIV_GHAT calc_div_ghat.F:60-150, implicDiv2DFlow = 1; kSurfC = 1 on wet
 * columns, Nr+1 on land).  cg3d_b is tile3d. */
void og_solve_rhs_nh(const og_grid *g, const og_params *p, int bi, int bj, const double *Bo_surf, const double *etaN,
                     const double *phi_nh, const double *gU, const double *gV, double *cg2d_b, double *cg2d_x,
                     double *cg3d_b) {
  SETUP
  const size_t ns = px * py;
  double *xA = (double *)calloc(ns * 3, sizeof(double));
  double *yA = xA + ns, *pf = xA + 2 * ns;
  FORALL {
    G2(cg2d_x, i, j) = G2(Bo_surf, i, j) * G2(etaN, i, j);
    G2(cg2d_b, i, j) = 0.;
  }
  for (int k = 1; k <= Nr; k++) FORALL G3(cg3d_b, i, j, k) = 0.;
  for (int k = Nr; k >= 1; k--) {
    for (int j = 1; j <= sNy + 1; j++)
      for (int i = 1; i <= sNx + 1; i++) {
        xA[S(i, j)] = G2(g->dyG, i, j) * g->drF[k - 1] * G3(g->hFacW, i, j, k);
        yA[S(i, j)] = G2(g->dxG, i, j) * g->drF[k - 1] * G3(g->hFacS, i, j, k);
      }
    for (int j = 1; j <= sNy; j++)
      for (int i = 1; i <= sNx + 1; i++) pf[S(i, j)] = xA[S(i, j)] * G3(gU, i, j, k) / p->deltaTMom;
    for (int j = 1; j <= sNy; j++)
      for (int i = 1; i <= sNx; i++) {
        G2(cg2d_b, i, j) = G2(cg2d_b, i, j) + pf[S(i + 1, j)] - pf[S(i, j)];
        G3(cg3d_b, i, j, k) = (pf[S(i + 1, j)] - pf[S(i, j)]);
      }
    for (int j = 1; j <= sNy + 1; j++)
      for (int i = 1; i <= sNx; i++) pf[S(i, j)] = yA[S(i, j)] * G3(gV, i, j, k) / p->deltaTMom;
    for (int j = 1; j <= sNy; j++)
      for (int i = 1; i <= sNx; i++) {
        G2(cg2d_b, i, j) = G2(cg2d_b, i, j) + pf[S(i, j + 1)] - pf[S(i, j)];
        G3(cg3d_b, i, j, k) = G3(cg3d_b, i, j, k) + (pf[S(i, j + 1)] - pf[S(i, j)]);
      }
  }
  for (int j = 1; j <= sNy; j++)
    for (int i = 1; i <= sNx; i++) {
      int ks = Nr + 1;
      for (int k = 1; k <= Nr; k++)
        if (G3(g->maskC, i, j, k) != 0.) { ks = k; break; }
      if (ks <= Nr) {
        const double t = p->freeSurfFac * G2(g->rA, i, j) * 1. / p->deltaTMom / p->deltaTFreeSurf
                         * (G2(etaN, i, j) + G3(phi_nh, i, j, ks) * G2(g->recip_Bo, i, j));
        G2(cg2d_b, i, j) = G2(cg2d_b, i, j) - t;
        G3(cg3d_b, i, j, ks) = G3(cg3d_b, i, j, ks) - t;
      }
    }
  free(xA);
}

/* PRE_CG3D for one tile (pre_cg3d.F:60-250): oldFreeSurfTerm, uniformFreeSurfLev (surfFac = freeSurfFac),
 * no fresh-water flux, no OBCS.  cg2d_x: the surface-pressure solution WITH its halo; wVel: the stepped w*. */
void og_pre_cg3d(const og_grid *g, const og_params *p, int bi, int bj, const double *cg2d_x, const double *etaN,
                 const double *wVel, double *cg3d_b) {
  SETUP
  const size_t ns = px * py;
  double *uf = (double *)calloc(ns * 2, sizeof(double));
  double *vf = uf + ns;
  const double surfFac = p->freeSurfFac * 1.;
  for (int j = 1; j <= sNy + 1; j++)
    for (int i = 1; i <= sNx + 1; i++) {
      uf[S(i, j)] = -G2(g->recip_dxC, i, j) * p->implicSurfPress * p->implicDiv2DFlow * (G2(cg2d_x, i, j) - G2(cg2d_x, i - 1, j));
      vf[S(i, j)] = -G2(g->recip_dyC, i, j) * p->implicSurfPress * p->implicDiv2DFlow * (G2(cg2d_x, i, j) - G2(cg2d_x, i, j - 1));
    }
  for (int k = 1; k <= Nr; k++) {
    const int kp1 = k + 1 < Nr ? k + 1 : Nr;
    const double wFacKm = p->implicDiv2DFlow * 1. * 1.;
    double wFacKp = p->implicDiv2DFlow * 1. * 1.;
    if (k >= Nr) wFacKp = 0.;
    for (int j = 1; j <= sNy; j++)
      for (int i = 1; i <= sNx; i++) {
        double last;
        if (k == 1) last = (surfFac * G2(etaN, i, j) / p->deltaTFreeSurf - G3(wVel, i, j, kp1) * wFacKp) * G2(g->rA, i, j) / p->deltaTMom;
        else last = (G3(wVel, i, j, k) * wFacKm * G3(g->maskC, i, j, k - 1) - G3(wVel, i, j, kp1) * wFacKp) * G2(g->rA, i, j) / p->deltaTMom;
        G3(cg3d_b, i, j, k) = G3(cg3d_b, i, j, k)
            + g->drF[k - 1] * G2(g->dyG, i + 1, j) * G3(g->hFacW, i + 1, j, k) * uf[S(i + 1, j)]
            - g->drF[k - 1] * G2(g->dyG, i, j) * G3(g->hFacW, i, j, k) * uf[S(i, j)]
            + g->drF[k - 1] * G2(g->dxG, i, j + 1) * G3(g->hFacS, i, j + 1, k) * vf[S(i, j + 1)]
            - g->drF[k - 1] * G2(g->dxG, i, j) * G3(g->hFacS, i, j, k) * vf[S(i, j)]
            + last;
      }
  }
  free(uf);
}

/* MOMENTUM_CORRECTION_STEP for one tile with the 3-D solver: CALC_GRAD_PHI_SURF on 2-OL..sN+OL, then
 * CORRECTION_STEP (correction_step.F:152-231) with nhFac = pfFacMom*implicitNHPress = 1. */
void og_correction_step_nh(const og_grid *g, const og_params *p, int bi, int bj, const double *Bo_surf, const double *etaN,
                           const double *phi_nh, const double *gU, const double *gV, double *uVel, double *vVel) {
  SETUP
  const size_t ns = px * py;
  double *phiSurfX = (double *)calloc(ns * 2, sizeof(double));
  double *phiSurfY = phiSurfX + ns;
  const int iMin = 2 - OLx, iMax = sNx + OLx, jMin = 2 - OLy, jMax = sNy + OLy;
  for (int j = jMin; j <= jMax; j++)
    for (int i = iMin; i <= iMax; i++) {
      phiSurfX[S(i, j)] = G2(g->recip_dxC, i, j) * (G2(Bo_surf, i, j) * G2(etaN, i, j) - G2(Bo_surf, i - 1, j) * G2(etaN, i - 1, j));
      phiSurfY[S(i, j)] = G2(g->recip_dyC, i, j) * (G2(Bo_surf, i, j) * G2(etaN, i, j) - G2(Bo_surf, i, j - 1) * G2(etaN, i, j - 1));
    }
  const double psFac = 1. * p->implicSurfPress, nhFac = 1. * 1.;
  for (int k = 1; k <= Nr; k++)
    for (int j = jMin; j <= jMax; j++)
      for (int i = iMin; i <= iMax; i++) {
        const double gU_dpx = -(psFac * phiSurfX[S(i, j)]
                                + nhFac * G2(g->recip_dxC, i, j) * (G3(phi_nh, i, j, k) - G3(phi_nh, i - 1, j, k)))
                              * G3(g->maskW, i, j, k);
        const double gV_dpy = -(psFac * phiSurfY[S(i, j)]
                                + nhFac * G2(g->recip_dyC, i, j) * (G3(phi_nh, i, j, k) - G3(phi_nh, i, j - 1, k)))
                              * G3(g->maskS, i, j, k);
        G3(uVel, i, j, k) = (G3(gU, i, j, k) + p->deltaTMom * gU_dpx) * G3(g->maskW, i, j, k);
        G3(vVel, i, j, k) = (G3(gV, i, j, k) + p->deltaTMom * gV_dpy) * G3(g->maskS, i, j, k);
      }
  free(phiSurfX);
}
