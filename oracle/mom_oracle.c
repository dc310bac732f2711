/* mom_oracle.c -- CPU restatement of MOM_FLUXFORM and the leaves it calls.
 * TEST INFRASTRUCTURE ONLY (see mitgcm_oracle.h).  Each block cites the
 * reference routine it follows; loop ranges are the reference's so that whole
 * slabs can be compared.  Options not restated (and rejected by the product as
 * unsupported): variable viscosity, OBCS, shelf ice, NH / 3-D Coriolis metric
 * terms, cylindrical grid, r* / non-linear free surface, MOM_BOUNDARY_CONSERVE,
 * deep-atmosphere / anelastic factors (== 1).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "mitgcm_oracle.h"

#define PXd (d->sNx + 2 * d->OLx)
#define PYd (d->sNy + 2 * d->OLy)
#define S(i, j) ((size_t)((i) + OLx - 1) + (size_t)px * (size_t)((j) + OLy - 1))
#define G2(a, i, j) (a)[S(i, j) + off2]
#define G3(a, i, j, k) (a)[S(i, j) + (size_t)px * py * (size_t)((k)-1) + off3]
#define K3(a, i, j, k) (a)[S(i, j) + (size_t)px * py * (size_t)((k)-1)] /* per-tile (slab,Nr[+1]) */
#define FORALL for (int j = 1 - OLy; j <= sNy + OLy; j++) for (int i = 1 - OLx; i <= sNx + OLx; i++)

#define SETUP                                                                          \
  const og_dims *d = &g->d;                                                            \
  const int sNx = d->sNx, sNy = d->sNy, OLx = d->OLx, OLy = d->OLy, Nr = d->Nr;        \
  const size_t px = PXd, py = PYd;                                                     \
  const size_t tile = (size_t)(bi - 1) + (size_t)d->nSx * (size_t)(bj - 1);            \
  const size_t off2 = px * py * tile, off3 = px * py * (size_t)Nr * tile;              \
  const size_t offc = py * tile;                                                       \
  (void)Nr; (void)off2; (void)off3; (void)offc;

/* MOM_CALC_HFACZ, pkg/mom_common/mom_calc_hfacz.F:160-377 (hZoption = 0) */
static void mom_calc_hfacz(const og_grid *g, int bi, int bj, int k, double *hFacZ, double *r_hFacZ) {
  SETUP
  for (int i = 1 - OLx; i <= sNx + OLx; i++) hFacZ[S(i, 1 - OLy)] = 0.;
  for (int j = 2 - OLy; j <= sNy + OLy; j++) hFacZ[S(1 - OLx, j)] = 0.;
  for (int j = 2 - OLy; j <= sNy + OLy; j++)
    for (int i = 2 - OLx; i <= sNx + OLx; i++) {
      double h = fmin(G3(g->hFacW, i, j, k), G3(g->hFacW, i, j - 1, k));
      h = fmin(G3(g->hFacS, i, j, k), h);
      h = fmin(G3(g->hFacS, i - 1, j, k), h);
      hFacZ[S(i, j)] = h;
    }
  FORALL {
    if (hFacZ[S(i, j)] == 0.) r_hFacZ[S(i, j)] = 0.;
    else r_hFacZ[S(i, j)] = 1. / hFacZ[S(i, j)];
  }
}

/* MOM_CALC_KE, pkg/mom_common/mom_calc_ke.F (KEscheme -1..3) */
static void mom_calc_ke(const og_grid *g, int bi, int bj, int k, int KEscheme,
                        const double *uFld, const double *vFld, double *KE) {
  SETUP
  for (int j = 1 - OLy; j <= sNy + OLy - 1; j++)
    for (int i = 1 - OLx; i <= sNx + OLx - 1; i++) {
      double u0 = uFld[S(i, j)], u1 = uFld[S(i + 1, j)], v0 = vFld[S(i, j)], v1 = vFld[S(i, j + 1)];
      if (KEscheme == -1)
        KE[S(i, j)] = 0.125 * ((u0 + u1) * (u0 + u1) + (v0 + v1) * (v0 + v1));
      else if (KEscheme == 0)
        KE[S(i, j)] = 0.25 * ((u0 * u0 + u1 * u1) + (v0 * v0 + v1 * v1));
      else if (KEscheme == 1)
        KE[S(i, j)] = 0.25 * ((u0 * u0 * G2(g->rAw, i, j) + u1 * u1 * G2(g->rAw, i + 1, j))
                            + (v0 * v0 * G2(g->rAs, i, j) + v1 * v1 * G2(g->rAs, i, j + 1))) * G2(g->recip_rA, i, j);
      else if (KEscheme == 2)
        KE[S(i, j)] = 0.25 * ((u0 * u0 * G3(g->hFacW, i, j, k) + u1 * u1 * G3(g->hFacW, i + 1, j, k))
                            + (v0 * v0 * G3(g->hFacS, i, j, k) + v1 * v1 * G3(g->hFacS, i, j + 1, k)))
                      * G3(g->recip_hFacC, i, j, k);
      else
        KE[S(i, j)] = 0.25 * ((u0 * u0 * G3(g->hFacW, i, j, k) * G2(g->rAw, i, j)
                             + u1 * u1 * G3(g->hFacW, i + 1, j, k) * G2(g->rAw, i + 1, j))
                            + (v0 * v0 * G3(g->hFacS, i, j, k) * G2(g->rAs, i, j)
                             + v1 * v1 * G3(g->hFacS, i, j + 1, k) * G2(g->rAs, i, j + 1)))
                      * G3(g->recip_hFacC, i, j, k) * G2(g->recip_rA, i, j);
    }
}

/* MOM_CALC_RTRANS, pkg/mom_fluxform/mom_calc_rtrans.F (select_rStar = 0) */
static void mom_calc_rtrans(const og_grid *g, int bi, int bj, int k, const double *wVel,
                            double *rTransU, double *rTransV) {
  SETUP
  for (int j = 2 - OLy; j <= sNy + OLy; j++)
    for (int i = 2 - OLx; i <= sNx + OLx; i++) {
      if (k > Nr) {
        rTransU[S(i, j)] = 0.;
        rTransV[S(i, j)] = 0.;
      } else {
        rTransU[S(i, j)] = 0.5 * (G3(wVel, i - 1, j, k) * G2(g->rA, i - 1, j) + G3(wVel, i, j, k) * G2(g->rA, i, j));
        rTransV[S(i, j)] = 0.5 * (G3(wVel, i, j - 1, k) * G2(g->rA, i, j - 1) + G3(wVel, i, j, k) * G2(g->rA, i, j));
      }
    }
}

/* MOM_U_ADV_WU / MOM_V_ADV_WV, pkg/mom_fluxform/mom_{u_adv_wu,v_adv_wv}.F.
 * isV = 0: U component (neighbour i-1), isV = 1: V component (neighbour j-1). */
static void mom_adv_w(const og_grid *g, const og_params *p, int bi, int bj, int k, int isV,
                      const double *fld, const double *wVel, const double *rTrans, double *flux) {
  SETUP
  if (k > Nr || (k == 1 && p->rigidLid)) {
    FORALL flux[S(i, j)] = 0.;
  } else if (k == 1) {
    for (int j = 2 - OLy; j <= sNy + OLy; j++)
      for (int i = 2 - OLx; i <= sNx + OLx; i++) flux[S(i, j)] = rTrans[S(i, j)] * G3(fld, i, j, k);
  } else {
    for (int j = 2 - OLy; j <= sNy + OLy; j++)
      for (int i = 2 - OLx; i <= sNx + OLx; i++)
        flux[S(i, j)] = rTrans[S(i, j)] * 0.5 * (G3(fld, i, j, k) + G3(fld, i, j, k - 1));
    if (p->select_rStar == 0 && !p->rigidLid) {
      const int di = isV ? 0 : 1, dj = isV ? 1 : 0;
      for (int j = 2 - OLy; j <= sNy + OLy; j++)
        for (int i = 2 - OLx; i <= sNx + OLx; i++)
          flux[S(i, j)] = flux[S(i, j)]
              + 0.25 * (G3(wVel, i, j, k) * G2(g->rA, i, j) * (G3(g->maskC, i, j, k) - G3(g->maskC, i, j, k - 1))
                      + G3(wVel, i - di, j - dj, k) * G2(g->rA, i - di, j - dj)
                          * (G3(g->maskC, i - di, j - dj, k) - G3(g->maskC, i - di, j - dj, k - 1)))
                  * G3(fld, i, j, k);
    }
  }
}

/* MOM_U_RVISCFLUX / MOM_V_RVISCFLUX, pkg/mom_common/mom_{u,v}_rviscflux.F */
static void mom_rviscflux(const og_grid *g, const og_params *p, int bi, int bj, int k, int isV,
                          const double *fld, const double *kappaR, double *flux) {
  SETUP
  if (k <= 1 || k > Nr) {
    FORALL flux[S(i, j)] = 0.;
  } else {
    const double *rAx = isV ? g->rAs : g->rAw;
    const double *mask = isV ? g->maskS : g->maskW;
    for (int j = 1 - OLy; j <= sNy + OLy - 1; j++)
      for (int i = 1 - OLx; i <= sNx + OLx - 1; i++)
        flux[S(i, j)] = -K3(kappaR, i, j, k) * G2(rAx, i, j)
            * (G3(fld, i, j, k) - G3(fld, i, j, k - 1)) * p->rkSign * g->recip_drC[k - 1]
            * G3(mask, i, j, k) * G3(mask, i, j, k - 1);
  }
}

/* MOM_U_DEL2U, pkg/mom_fluxform/mom_u_del2u.F */
static void mom_u_del2u(const og_grid *g, const og_params *p, int bi, int bj, int k,
                        const double *uFld, const double *hFacZ, const double *h0FacZ, double *del2u) {
  SETUP
  double *fZon = (double *)calloc(px * py, sizeof(double));
  double *fMer = (double *)calloc(px * py, sizeof(double));
  for (int j = 2 - OLy; j <= sNy + OLy - 1; j++)
    for (int i = 1 - OLx; i <= sNx + OLx - 1; i++)
      fZon[S(i, j)] = g->drF[k - 1] * G3(g->hFacC, i, j, k) * G2(g->dyF, i, j) * G2(g->recip_dxF, i, j)
                      * (uFld[S(i + 1, j)] - uFld[S(i, j)]);
  for (int j = 2 - OLy; j <= sNy + OLy; j++)
    for (int i = 2 - OLx; i <= sNx + OLx - 1; i++)
      fMer[S(i, j)] = g->drF[k - 1] * hFacZ[S(i, j)] * G2(g->dxV, i, j) * G2(g->recip_dyU, i, j)
                      * (uFld[S(i, j)] - uFld[S(i, j - 1)]);
  for (int j = 2 - OLy; j <= sNy + OLy - 1; j++)
    for (int i = 2 - OLx; i <= sNx + OLx - 1; i++)
      del2u[S(i, j)] = g->recip_drF[k - 1] * G3(g->recip_hFacW, i, j, k) * G2(g->recip_rAw, i, j)
                       * (fZon[S(i, j)] - fZon[S(i - 1, j)] + fMer[S(i, j + 1)] - fMer[S(i, j)])
                       * G3(g->maskW, i, j, k);
  if (p->no_slip_sides)
    for (int j = 2 - OLy; j <= sNy + OLy - 1; j++)
      for (int i = 2 - OLx; i <= sNx + OLx - 1; i++) {
        double hS = G3(g->hFacW, i, j, k) - h0FacZ[S(i, j)];
        double hN = G3(g->hFacW, i, j, k) - h0FacZ[S(i, j + 1)];
        del2u[S(i, j)] = del2u[S(i, j)]
            - G3(g->recip_hFacW, i, j, k) * G2(g->recip_rAw, i, j)
              * (hS * G2(g->dxV, i, j) * G2(g->recip_dyU, i, j) + hN * G2(g->dxV, i, j + 1) * G2(g->recip_dyU, i, j + 1))
              * uFld[S(i, j)] * p->sideDragFactor * G3(g->maskW, i, j, k);
      }
  free(fZon); free(fMer);
}

/* MOM_V_DEL2V, pkg/mom_fluxform/mom_v_del2v.F */
static void mom_v_del2v(const og_grid *g, const og_params *p, int bi, int bj, int k,
                        const double *vFld, const double *hFacZ, const double *h0FacZ, double *del2v) {
  SETUP
  double *fZon = (double *)calloc(px * py, sizeof(double));
  double *fMer = (double *)calloc(px * py, sizeof(double));
  for (int j = 2 - OLy; j <= sNy + OLy - 1; j++)
    for (int i = 2 - OLx; i <= sNx + OLx; i++)
      fZon[S(i, j)] = g->drF[k - 1] * hFacZ[S(i, j)] * G2(g->dyU, i, j) * G2(g->recip_dxV, i, j)
                      * (vFld[S(i, j)] - vFld[S(i - 1, j)]);
  for (int j = 1 - OLy; j <= sNy + OLy - 1; j++)
    for (int i = 2 - OLx; i <= sNx + OLx - 1; i++)
      fMer[S(i, j)] = g->drF[k - 1] * G3(g->hFacC, i, j, k) * G2(g->dxF, i, j) * G2(g->recip_dyF, i, j)
                      * (vFld[S(i, j + 1)] - vFld[S(i, j)]);
  for (int j = 2 - OLy; j <= sNy + OLy - 1; j++)
    for (int i = 2 - OLx; i <= sNx + OLx - 1; i++)
      del2v[S(i, j)] = g->recip_drF[k - 1] * G3(g->recip_hFacS, i, j, k) * G2(g->recip_rAs, i, j)
                       * (fZon[S(i + 1, j)] - fZon[S(i, j)] + fMer[S(i, j)] - fMer[S(i, j - 1)])
                       * G3(g->maskS, i, j, k);
  if (p->no_slip_sides)
    for (int j = 2 - OLy; j <= sNy + OLy - 1; j++)
      for (int i = 2 - OLx; i <= sNx + OLx - 1; i++) {
        double hW = G3(g->hFacS, i, j, k) - h0FacZ[S(i, j)];
        double hE = G3(g->hFacS, i, j, k) - h0FacZ[S(i + 1, j)];
        del2v[S(i, j)] = del2v[S(i, j)]
            - G3(g->recip_hFacS, i, j, k) * G2(g->recip_rAs, i, j)
              * (hW * G2(g->dyU, i, j) * G2(g->recip_dxV, i, j) + hE * G2(g->dyU, i + 1, j) * G2(g->recip_dxV, i + 1, j))
              * vFld[S(i, j)] * p->sideDragFactor * G3(g->maskS, i, j, k);
      }
  free(fZon); free(fMer);
}

/* MOM_U_SIDEDRAG / MOM_V_SIDEDRAG, pkg/mom_common/mom_{u,v}_sidedrag.F,
 * branch sideDragFactor > 0 with constant viscAh_Z / viscA4_Z (no variable
 * viscosity).  The sideDragFactor <= 0 branch (viscAhGrid-dependent) is not
 * restated; the product rejects it. */
static void mom_sidedrag(const og_grid *g, const og_params *p, int bi, int bj, int k, int isV,
                         const double *fld, const double *del2, const double *hFacZ, double *drag) {
  SETUP
  for (int j = 2 - OLy; j <= sNy + OLy - 1; j++)
    for (int i = 2 - OLx; i <= sNx + OLx - 1; i++) {
      if (!isV) {
        double hS = G3(g->hFacW, i, j, k) - hFacZ[S(i, j)];
        double hN = G3(g->hFacW, i, j, k) - hFacZ[S(i, j + 1)];
        drag[S(i, j)] = -G3(g->recip_hFacW, i, j, k) * g->recip_drF[k - 1] * G2(g->recip_rAw, i, j)
            * (hS * G2(g->dxV, i, j) * G2(g->recip_dyU, i, j)
                 * (p->viscAhZ * fld[S(i, j)] - p->viscA4Z * del2[S(i, j)])
             + hN * G2(g->dxV, i, j + 1) * G2(g->recip_dyU, i, j + 1)
                 * (p->viscAhZ * fld[S(i, j)] - p->viscA4Z * del2[S(i, j)]))
            * g->drF[k - 1] * p->sideDragFactor;
      } else {
        const double cf = g->cosFacV[(j + OLy - 1) + offc];
        double hW = G3(g->hFacS, i, j, k) - hFacZ[S(i, j)];
        double hE = G3(g->hFacS, i, j, k) - hFacZ[S(i + 1, j)];
        drag[S(i, j)] = -G3(g->recip_hFacS, i, j, k) * g->recip_drF[k - 1] * G2(g->recip_rAs, i, j)
            * (hW * G2(g->dyU, i, j) * G2(g->recip_dxV, i, j)
                 * (p->viscAhZ * fld[S(i, j)] * cf - p->viscA4Z * del2[S(i, j)] * cf)
             + hE * G2(g->dyU, i + 1, j) * G2(g->recip_dxV, i + 1, j)
                 * (p->viscAhZ * fld[S(i, j)] * cf - p->viscA4Z * del2[S(i, j)] * cf))
            * g->drF[k - 1] * p->sideDragFactor;
      }
    }
}

/* MOM_U_BOTDRAG_COEFF / MOM_V_BOTDRAG_COEFF (z coordinates, inp_KE = .TRUE.),
 * pkg/mom_common/mom_{u,v}_botdrag_coeff.F */
static void mom_botdrag_coeff(const og_grid *g, const og_params *p, int bi, int bj, int k, int isV,
                              const double *uFld, const double *vFld, const double *kappaR,
                              const double *KE, double *cDrag) {
  SETUP
  const double viscFac = p->no_slip_bottom ? 2. : 0.;
  const int kBottom = Nr, kDown = (k + 1 < Nr) ? k + 1 : Nr, kLowF = k + 1;
  const double dragFac = 1.;
  const double recDrC = (k == kBottom) ? g->recip_drF[k - 1] : g->recip_drC[kLowF - 1];
  const double *mask = isV ? g->maskS : g->maskW;
  const double *rhF = isV ? g->recip_hFacS : g->recip_hFacW;
  const int di = isV ? 0 : 1, dj = isV ? 1 : 0;
  /* loop ranges: U: j full, i from 2-OLx ; V: i full, j from 2-OLy */
  const int iLo = 1 - OLx + di, jLo = 1 - OLy + dj;
  for (int j = jLo; j <= sNy + OLy; j++)
    for (int i = iLo; i <= sNx + OLx; i++) cDrag[S(i, j)] = p->bottomDragLinear * dragFac;
  const int iHi = sNx + OLx - 1, jHi = sNy + OLy - 1;
  if (p->no_slip_bottom && p->bottomVisc_pCell) {
    for (int j = jLo; j <= jHi; j++)
      for (int i = iLo; i <= iHi; i++)
        cDrag[S(i, j)] = cDrag[S(i, j)] + K3(kappaR, i, j, kLowF) * recDrC * viscFac * G3(rhF, i, j, k);
  } else if (p->no_slip_bottom) {
    for (int j = jLo; j <= jHi; j++)
      for (int i = iLo; i <= iHi; i++)
        cDrag[S(i, j)] = cDrag[S(i, j)] + K3(kappaR, i, j, kLowF) * recDrC * viscFac;
  }
  if (p->selectBotDragQuadr == 0) {
    for (int j = jLo; j <= jHi; j++)
      for (int i = iLo; i <= iHi; i++) {
        double ks = KE[S(i, j)] + KE[S(i - di, j - dj)];
        if (ks > 0.) cDrag[S(i, j)] = cDrag[S(i, j)] + p->bottomDragQuadratic * sqrt(ks) * dragFac;
      }
  } else if (p->selectBotDragQuadr == 1 || p->selectBotDragQuadr == 2) {
    for (int j = jLo; j <= jHi; j++)
      for (int i = iLo; i <= iHi; i++) {
        double uSq;
        if (!isV) {
          double a = (vFld[S(i - 1, j)] * vFld[S(i - 1, j)] * G3(g->hFacS, i - 1, j, k)
                    + vFld[S(i, j)] * vFld[S(i, j)] * G3(g->hFacS, i, j, k))
                   + (vFld[S(i - 1, j + 1)] * vFld[S(i - 1, j + 1)] * G3(g->hFacS, i - 1, j + 1, k)
                    + vFld[S(i, j + 1)] * vFld[S(i, j + 1)] * G3(g->hFacS, i, j + 1, k));
          if (p->selectBotDragQuadr == 1)
            uSq = uFld[S(i, j)] * uFld[S(i, j)] + a * G3(g->recip_hFacW, i, j, k) * 0.25;
          else {
            double h = (G3(g->hFacS, i - 1, j, k) + G3(g->hFacS, i, j, k))
                     + (G3(g->hFacS, i - 1, j + 1, k) + G3(g->hFacS, i, j + 1, k));
            if (h > 0.) uSq = uFld[S(i, j)] * uFld[S(i, j)] + a / h;
            else uSq = uFld[S(i, j)] * uFld[S(i, j)];
          }
        } else {
          double a = (uFld[S(i, j - 1)] * uFld[S(i, j - 1)] * G3(g->hFacW, i, j - 1, k)
                    + uFld[S(i, j)] * uFld[S(i, j)] * G3(g->hFacW, i, j, k))
                   + (uFld[S(i + 1, j - 1)] * uFld[S(i + 1, j - 1)] * G3(g->hFacW, i + 1, j - 1, k)
                    + uFld[S(i + 1, j)] * uFld[S(i + 1, j)] * G3(g->hFacW, i + 1, j, k));
          if (p->selectBotDragQuadr == 1)
            uSq = vFld[S(i, j)] * vFld[S(i, j)] + a * G3(g->recip_hFacS, i, j, k) * 0.25;
          else {
            double h = (G3(g->hFacW, i, j - 1, k) + G3(g->hFacW, i, j, k))
                     + (G3(g->hFacW, i + 1, j - 1, k) + G3(g->hFacW, i + 1, j, k));
            if (h > 0.) uSq = vFld[S(i, j)] * vFld[S(i, j)] + a / h;
            else uSq = vFld[S(i, j)] * vFld[S(i, j)];
          }
        }
        if (uSq > 0.) cDrag[S(i, j)] = cDrag[S(i, j)] + p->bottomDragQuadratic * sqrt(uSq) * dragFac;
      }
  }
  for (int j = jLo; j <= sNy + OLy; j++)
    for (int i = iLo; i <= sNx + OLx; i++) {
      if (k == kBottom) cDrag[S(i, j)] = cDrag[S(i, j)] * G3(mask, i, j, k);
      else cDrag[S(i, j)] = cDrag[S(i, j)] * G3(mask, i, j, k) * (1. - G3(mask, i, j, kDown));
    }
}

void og_mom_fluxform(const og_grid *g, const og_params *p, int bi, int bj, int k,
                     int iMin, int iMax, int jMin, int jMax,
                     const double *kappaRU, const double *kappaRV,
                     double *fVerUkm, double *fVerVkm, double *fVerUkp, double *fVerVkp,
                     double *guDiss, double *gvDiss,
                     const double *uVel, const double *vVel, const double *wVel,
                     double *gU, double *gV) {
  SETUP
  const size_t ns = px * py;
  double *buf = (double *)calloc(ns * 22, sizeof(double));
  double *vF = buf, *v4F = buf + ns, *uCf = buf + 2 * ns, *vCf = buf + 3 * ns, *mT = buf + 4 * ns,
         *fZon = buf + 5 * ns, *fMer = buf + 6 * ns, *fVrUp = buf + 7 * ns, *fVrDw = buf + 8 * ns,
         *rTransU = buf + 9 * ns, *rTransV = buf + 10 * ns, *hFacZ = buf + 11 * ns,
         *h0FacZ = buf + 12 * ns, *r_hFacZ = buf + 13 * ns, *xA = buf + 14 * ns, *yA = buf + 15 * ns,
         *uTrans = buf + 16 * ns, *vTrans = buf + 17 * ns, *uFld = buf + 18 * ns, *vFld = buf + 19 * ns,
         *KE = buf + 20 * ns, *cDrag = buf + 21 * ns;
  /* mom_fluxform.F:203-232: zero temporaries and dissipation outputs */
  FORALL { guDiss[S(i, j)] = 0.; gvDiss[S(i, j)] = 0.; }

  /* :236-255 term switches */
  const double uDudxFac = p->afFacMom, AhDudxFac = p->vfFacMom, vDudyFac = p->afFacMom,
               AhDudyFac = p->vfFacMom, rVelDudrFac = p->afFacMom;
  double ArDudrFac = p->vfFacMom;
  const double mtFacU = p->mtFacMom, fuFac = p->cfFacMom;
  const double uDvdxFac = p->afFacMom, AhDvdxFac = p->vfFacMom, vDvdyFac = p->afFacMom,
               AhDvdyFac = p->vfFacMom, rVelDvdrFac = p->afFacMom;
  double ArDvdrFac = p->vfFacMom;
  const double mtFacV = p->mtFacMom, fvFac = p->cfFacMom;
  const int metricTerms = p->selectMetricTerms >= 1;
  if (p->implicitViscosity) { ArDudrFac = 0.; ArDvdrFac = 0.; }
  /* :272-279 */
  const int bottomDragTerms = (p->selectImplicitDrag == 0 &&
      (p->no_slip_bottom || p->selectBotDragQuadr >= 0 || p->bottomDragLinear != 0.));

  mom_calc_hfacz(g, bi, bj, k, hFacZ, r_hFacZ);
  /* :287-327 */
  FORALL {
    xA[S(i, j)] = G2(g->dyG, i, j) * g->drF[k - 1] * G3(g->hFacW, i, j, k);
    yA[S(i, j)] = G2(g->dxG, i, j) * g->drF[k - 1] * G3(g->hFacS, i, j, k);
    h0FacZ[S(i, j)] = hFacZ[S(i, j)];
    uFld[S(i, j)] = G3(uVel, i, j, k);
    vFld[S(i, j)] = G3(vVel, i, j, k);
    uTrans[S(i, j)] = uFld[S(i, j)] * xA[S(i, j)];
    vTrans[S(i, j)] = vFld[S(i, j)] * yA[S(i, j)];
  }
  mom_calc_ke(g, bi, bj, k, 2, uFld, vFld, KE); /* :329, KEscheme argument is the literal 2 */

  /* :384-417 first call: surface vertical advective flux */
  if (p->momAdvection && k == 1) {
    mom_calc_rtrans(g, bi, bj, k, wVel, rTransU, rTransV);
    mom_adv_w(g, p, bi, bj, k, 0, uVel, wVel, rTransU, fVerUkm);
    mom_adv_w(g, p, bi, bj, k, 1, vVel, wVel, rTransV, fVerVkm);
  }
  /* :420-424 */
  if (p->momAdvection) mom_calc_rtrans(g, bi, bj, k + 1, wVel, rTransU, rTransV);

  /* ---- zonal momentum ---------------------------------------------------- */
  if (p->momAdvection) {
    /* MOM_U_ADV_UU / _VU / _WU */
    for (int j = 1 - OLy; j <= sNy + OLy - 1; j++)
      for (int i = 1 - OLx; i <= sNx + OLx - 1; i++)
        fZon[S(i, j)] = 0.25 * (uTrans[S(i, j)] + uTrans[S(i + 1, j)]) * (uFld[S(i, j)] + uFld[S(i + 1, j)]);
    for (int j = 2 - OLy; j <= sNy + OLy; j++)
      for (int i = 2 - OLx; i <= sNx + OLx; i++)
        fMer[S(i, j)] = 0.25 * (vTrans[S(i, j)] + vTrans[S(i - 1, j)]) * (uFld[S(i, j)] + uFld[S(i, j - 1)]);
    mom_adv_w(g, p, bi, bj, k + 1, 0, uVel, wVel, rTransU, fVerUkp);
    /* :502-517 */
    for (int j = jMin; j <= jMax; j++)
      for (int i = iMin; i <= iMax; i++)
        G3(gU, i, j, k) = -G3(g->recip_hFacW, i, j, k) * g->recip_drF[k - 1] * G2(g->recip_rAw, i, j)
            * ((fZon[S(i, j)] - fZon[S(i - 1, j)]) * uDudxFac
             + (fMer[S(i, j + 1)] - fMer[S(i, j)]) * vDudyFac
             + (fVerUkp[S(i, j)] - fVerUkm[S(i, j)]) * p->rkSign * rVelDudrFac);
  } else {
    FORALL G3(gU, i, j, k) = 0.;
  }
  if (p->momViscosity) {
    if (p->useBiharmonicVisc) mom_u_del2u(g, p, bi, bj, k, uFld, hFacZ, h0FacZ, v4F);
    /* MOM_U_XVISCFLUX */
    for (int j = 1 - OLy; j <= sNy + OLy - 1; j++)
      for (int i = 1 - OLx; i <= sNx + OLx - 1; i++)
        fZon[S(i, j)] = G2(g->dyF, i, j) * g->drF[k - 1] * G3(g->hFacC, i, j, k)
            * (-p->viscAhD * (uFld[S(i + 1, j)] - uFld[S(i, j)]) * g->cosFacU[(j + OLy - 1) + offc]
               + p->viscA4D * (v4F[S(i + 1, j)] - v4F[S(i, j)]) * g->cosFacU[(j + OLy - 1) + offc])
            * G2(g->recip_dxF, i, j);
    /* MOM_U_YVISCFLUX */
    for (int j = 2 - OLy; j <= sNy + OLy; j++)
      for (int i = 1 - OLx; i <= sNx + OLx; i++)
        fMer[S(i, j)] = G2(g->dxV, i, j) * g->drF[k - 1] * hFacZ[S(i, j)]
            * (-p->viscAhZ * (uFld[S(i, j)] - uFld[S(i, j - 1)])
               + p->viscA4Z * (v4F[S(i, j)] - v4F[S(i, j - 1)]))
            * G2(g->recip_dyU, i, j);
    if (!p->implicitViscosity) {
      mom_rviscflux(g, p, bi, bj, k, 0, uVel, kappaRU, fVrUp);
      mom_rviscflux(g, p, bi, bj, k + 1, 0, uVel, kappaRU, fVrDw);
    }
    for (int j = jMin; j <= jMax; j++)
      for (int i = iMin; i <= iMax; i++)
        guDiss[S(i, j)] = -G3(g->recip_hFacW, i, j, k) * g->recip_drF[k - 1] * G2(g->recip_rAw, i, j)
            * ((fZon[S(i, j)] - fZon[S(i - 1, j)]) * AhDudxFac
             + (fMer[S(i, j + 1)] - fMer[S(i, j)]) * AhDudyFac
             + (fVrDw[S(i, j)] - fVrUp[S(i, j)]) * p->rkSign * ArDudrFac);
    if (p->no_slip_sides) {
      mom_sidedrag(g, p, bi, bj, k, 0, uFld, v4F, h0FacZ, vF);
      for (int j = jMin; j <= jMax; j++)
        for (int i = iMin; i <= iMax; i++) guDiss[S(i, j)] = guDiss[S(i, j)] + vF[S(i, j)];
    }
    if (bottomDragTerms) {
      mom_botdrag_coeff(g, p, bi, bj, k, 0, uFld, vFld, kappaRU, KE, cDrag);
      for (int j = jMin; j <= jMax; j++)
        for (int i = iMin; i <= iMax; i++)
          guDiss[S(i, j)] = guDiss[S(i, j)]
              - cDrag[S(i, j)] * uFld[S(i, j)] * G3(g->recip_hFacW, i, j, k) * g->recip_drF[k - 1];
    }
  }
  /* metric terms, MOM_U_METRIC_SPHERE */
  if (p->usingSphericalPolarGrid && metricTerms) {
    for (int j = 1 - OLy; j <= sNy + OLy - 1; j++)
      for (int i = 2 - OLx; i <= sNx + OLx; i++)
        mT[S(i, j)] = uFld[S(i, j)] * p->recip_rSphere
            * 0.25 * (vFld[S(i, j)] + vFld[S(i - 1, j)] + vFld[S(i, j + 1)] + vFld[S(i - 1, j + 1)])
            * G2(g->tanPhiAtU, i, j);
    for (int j = jMin; j <= jMax; j++)
      for (int i = iMin; i <= iMax; i++) G3(gU, i, j, k) = G3(gU, i, j, k) + mtFacU * mT[S(i, j)];
  }

  /* ---- meridional momentum ---------------------------------------------- */
  memset(v4F, 0, ns * sizeof(double)); /* v4F is reused; zero where del2v is not written */
  if (p->momAdvection) {
    for (int j = 2 - OLy; j <= sNy + OLy; j++)
      for (int i = 2 - OLx; i <= sNx + OLx; i++)
        fZon[S(i, j)] = 0.25 * (uTrans[S(i, j)] + uTrans[S(i, j - 1)]) * (vFld[S(i, j)] + vFld[S(i - 1, j)]);
    for (int j = 1 - OLy; j <= sNy + OLy - 1; j++)
      for (int i = 1 - OLx; i <= sNx + OLx - 1; i++)
        fMer[S(i, j)] = 0.25 * (vTrans[S(i, j)] + vTrans[S(i, j + 1)]) * (vFld[S(i, j)] + vFld[S(i, j + 1)]);
    mom_adv_w(g, p, bi, bj, k + 1, 1, vVel, wVel, rTransV, fVerVkp);
    for (int j = jMin; j <= jMax; j++)
      for (int i = iMin; i <= iMax; i++)
        G3(gV, i, j, k) = -G3(g->recip_hFacS, i, j, k) * g->recip_drF[k - 1] * G2(g->recip_rAs, i, j)
            * ((fZon[S(i + 1, j)] - fZon[S(i, j)]) * uDvdxFac
             + (fMer[S(i, j)] - fMer[S(i, j - 1)]) * vDvdyFac
             + (fVerVkp[S(i, j)] - fVerVkm[S(i, j)]) * p->rkSign * rVelDvdrFac);
  } else {
    FORALL G3(gV, i, j, k) = 0.;
  }
  if (p->momViscosity) {
    if (p->useBiharmonicVisc) mom_v_del2v(g, p, bi, bj, k, vFld, hFacZ, h0FacZ, v4F);
    /* MOM_V_XVISCFLUX */
    for (int j = 1 - OLy; j <= sNy + OLy; j++)
      for (int i = 2 - OLx; i <= sNx + OLx; i++)
        fZon[S(i, j)] = G2(g->dyU, i, j) * g->drF[k - 1] * hFacZ[S(i, j)]
            * (-p->viscAhZ * (vFld[S(i, j)] - vFld[S(i - 1, j)]) * g->cosFacV[(j + OLy - 1) + offc]
               + p->viscA4Z * (v4F[S(i, j)] - v4F[S(i - 1, j)]) * g->cosFacV[(j + OLy - 1) + offc])
            * G2(g->recip_dxV, i, j);
    /* MOM_V_YVISCFLUX */
    for (int j = 1 - OLy; j <= sNy + OLy - 1; j++)
      for (int i = 1 - OLx; i <= sNx + OLx - 1; i++)
        fMer[S(i, j)] = G2(g->dxF, i, j) * g->drF[k - 1] * G3(g->hFacC, i, j, k)
            * (-p->viscAhD * (vFld[S(i, j + 1)] - vFld[S(i, j)])
               + p->viscA4D * (v4F[S(i, j + 1)] - v4F[S(i, j)]))
            * G2(g->recip_dyF, i, j);
    if (!p->implicitViscosity) {
      mom_rviscflux(g, p, bi, bj, k, 1, vVel, kappaRV, fVrUp);
      mom_rviscflux(g, p, bi, bj, k + 1, 1, vVel, kappaRV, fVrDw);
    }
    for (int j = jMin; j <= jMax; j++)
      for (int i = iMin; i <= iMax; i++)
        gvDiss[S(i, j)] = -G3(g->recip_hFacS, i, j, k) * g->recip_drF[k - 1] * G2(g->recip_rAs, i, j)
            * ((fZon[S(i + 1, j)] - fZon[S(i, j)]) * AhDvdxFac
             + (fMer[S(i, j)] - fMer[S(i, j - 1)]) * AhDvdyFac
             + (fVrDw[S(i, j)] - fVrUp[S(i, j)]) * p->rkSign * ArDvdrFac);
    if (p->no_slip_sides) {
      mom_sidedrag(g, p, bi, bj, k, 1, vFld, v4F, h0FacZ, vF);
      for (int j = jMin; j <= jMax; j++)
        for (int i = iMin; i <= iMax; i++) gvDiss[S(i, j)] = gvDiss[S(i, j)] + vF[S(i, j)];
    }
    if (bottomDragTerms) {
      mom_botdrag_coeff(g, p, bi, bj, k, 1, uFld, vFld, kappaRV, KE, cDrag);
      for (int j = jMin; j <= jMax; j++)
        for (int i = iMin; i <= iMax; i++)
          gvDiss[S(i, j)] = gvDiss[S(i, j)]
              - cDrag[S(i, j)] * vFld[S(i, j)] * G3(g->recip_hFacS, i, j, k) * g->recip_drF[k - 1];
    }
  }
  /* MOM_V_METRIC_SPHERE */
  if (p->usingSphericalPolarGrid && metricTerms) {
    for (int j = 2 - OLy; j <= sNy + OLy; j++)
      for (int i = 1 - OLx; i <= sNx + OLx - 1; i++) {
        double ub = 0.25 * (uFld[S(i, j)] + uFld[S(i + 1, j)] + uFld[S(i, j - 1)] + uFld[S(i + 1, j - 1)]);
        mT[S(i, j)] = -p->recip_rSphere * ub * ub * G2(g->tanPhiAtV, i, j);
      }
    for (int j = jMin; j <= jMax; j++)
      for (int i = iMin; i <= iMax; i++) G3(gV, i, j, k) = G3(gV, i, j, k) + mtFacV * mT[S(i, j)];
  }

  /* ---- Coriolis, MOM_U_CORIOLIS / MOM_V_CORIOLIS ------------------------- */
  if (!p->useCDscheme) {
    for (int j = 1 - OLy; j <= sNy + OLy - 1; j++)
      for (int i = 2 - OLx; i <= sNx + OLx; i++) {
        if (p->selectCoriScheme >= 2)
          uCf[S(i, j)] = 0.5 * (G2(g->fCori, i, j) * 0.5 * (vFld[S(i, j)] + vFld[S(i, j + 1)])
                              + G2(g->fCori, i - 1, j) * 0.5 * (vFld[S(i - 1, j)] + vFld[S(i - 1, j + 1)]));
        else
          uCf[S(i, j)] = 0.5 * (G2(g->fCori, i, j) + G2(g->fCori, i - 1, j))
              * 0.25 * (vFld[S(i, j)] + vFld[S(i, j + 1)] + vFld[S(i - 1, j)] + vFld[S(i - 1, j + 1)]);
        if (p->selectCoriScheme == 1 || p->selectCoriScheme == 3)
          uCf[S(i, j)] = uCf[S(i, j)] * 4. / fmax(1., G3(g->maskS, i, j, k) + G3(g->maskS, i, j + 1, k)
                                                   + G3(g->maskS, i - 1, j, k) + G3(g->maskS, i - 1, j + 1, k));
      }
    for (int j = 2 - OLy; j <= sNy + OLy; j++)
      for (int i = 1 - OLx; i <= sNx + OLx - 1; i++) {
        if (p->selectCoriScheme >= 2)
          vCf[S(i, j)] = -0.5 * (G2(g->fCori, i, j) * 0.5 * (uFld[S(i, j)] + uFld[S(i + 1, j)])
                               + G2(g->fCori, i, j - 1) * 0.5 * (uFld[S(i, j - 1)] + uFld[S(i + 1, j - 1)]));
        else
          vCf[S(i, j)] = -0.5 * (G2(g->fCori, i, j) + G2(g->fCori, i, j - 1))
              * 0.25 * (uFld[S(i, j)] + uFld[S(i + 1, j)] + uFld[S(i, j - 1)] + uFld[S(i + 1, j - 1)]);
        if (p->selectCoriScheme == 1 || p->selectCoriScheme == 3)
          vCf[S(i, j)] = vCf[S(i, j)] * 4. / fmax(1., G3(g->maskW, i, j, k) + G3(g->maskW, i + 1, j, k)
                                                   + G3(g->maskW, i, j - 1, k) + G3(g->maskW, i + 1, j - 1, k));
      }
    for (int j = jMin; j <= jMax; j++)
      for (int i = iMin; i <= iMax; i++) {
        G3(gU, i, j, k) = G3(gU, i, j, k) + fuFac * uCf[S(i, j)];
        G3(gV, i, j, k) = G3(gV, i, j, k) + fvFac * vCf[S(i, j)];
      }
  }
  /* :1044-1051 */
  for (int j = jMin; j <= jMax; j++)
    for (int i = iMin; i <= iMax; i++) {
      G3(gU, i, j, k) = G3(gU, i, j, k) * G3(g->maskW, i, j, k);
      guDiss[S(i, j)] = guDiss[S(i, j)] * G3(g->maskW, i, j, k);
      G3(gV, i, j, k) = G3(gV, i, j, k) * G3(g->maskS, i, j, k);
      gvDiss[S(i, j)] = gvDiss[S(i, j)] * G3(g->maskS, i, j, k);
    }
  free(buf);
}

/* ======================================================================================
 * MOM_VECINV (pkg/mom_vecinv/mom_vecinv.F:10-1009) and its leaves.  Same scope limits as
 * above plus: variable
 * viscosity, strain-tension viscosity, Leith-QG, GGL90-Langmuir, NH Coriolis / metric
 * terms and momImplVertAdv are not restated (the product rejects them).
 * Cubed sphere: csCorners is a bit mask of the tile's facet corners (1 SW, 2 SE, 4 NE,
 * 8 NW; 0 on a non-cube topology), myFace the facet number (mom_calc_relvort3.F:79-97).
 * ====================================================================================== */

/* FILL_CS_CORNER_TR_RL (eesupp/src/fill_cs_corner_tr_rl.F), withSigns = .FALSE. */
static void fill_cs_corner_tr(const og_dims *d, int fill4dir, int csCorners, double *f) {
  const int sNx = d->sNx, sNy = d->sNy, OLx = d->OLx, OLy = d->OLy;
  const size_t px = PXd;
  if (!csCorners) return;
  for (int j = 1; j <= OLy; j++)
    for (int i = 1; i <= OLx; i++) {
      if (fill4dir == 1) {
        if (csCorners & 1) f[S(1 - i, 1 - j)] = f[S(1 - j, i)];
        if (csCorners & 2) f[S(sNx + i, 1 - j)] = f[S(sNx + j, i)];
        if (csCorners & 8) f[S(1 - i, sNy + j)] = f[S(1 - j, sNy + 1 - i)];
        if (csCorners & 4) f[S(sNx + i, sNy + j)] = f[S(sNx + j, sNy + 1 - i)];
      } else {
        if (csCorners & 1) f[S(1 - i, 1 - j)] = f[S(j, 1 - i)];
        if (csCorners & 2) f[S(sNx + i, 1 - j)] = f[S(sNx + 1 - j, 1 - i)];
        if (csCorners & 8) f[S(1 - i, sNy + j)] = f[S(j, sNy + i)];
        if (csCorners & 4) f[S(sNx + i, sNy + j)] = f[S(sNx + 1 - j, sNy + i)];
      }
    }
}

/* MOM_CALC_HDIV, hDivScheme = 2 (pkg/mom_common/mom_calc_hdiv.F:56-72) */
static void mom_calc_hdiv2(const og_grid *g, int bi, int bj, int k, const double *uFld, const double *vFld,
                           double *hDiv) {
  SETUP
  for (int j = 1 - OLy; j <= sNy + OLy - 1; j++)
    for (int i = 1 - OLx; i <= sNx + OLx - 1; i++)
      hDiv[S(i, j)] = ((uFld[S(i + 1, j)] * G2(g->dyG, i + 1, j) * G3(g->hFacW, i + 1, j, k)
                      - uFld[S(i, j)] * G2(g->dyG, i, j) * G3(g->hFacW, i, j, k))
                     + (vFld[S(i, j + 1)] * G2(g->dxG, i, j + 1) * G3(g->hFacS, i, j + 1, k)
                      - vFld[S(i, j)] * G2(g->dxG, i, j) * G3(g->hFacS, i, j, k)))
                    * G2(g->recip_rA, i, j) * G3(g->recip_hFacC, i, j, k);
}

/* MOM_CALC_RELVORT3 (pkg/mom_common/mom_calc_relvort3.F:64-304), CALC_CS_CORNER_EXTENDED undefined */
static void mom_calc_relvort3(const og_grid *g, int bi, int bj, int k, const double *uFld, const double *vFld,
                              double *vort3, int csCorners, int myFace) {
  SETUP
  (void)k;
#define VDY(i, j) (vFld[S(i, j)] * G2(g->dyC, i, j))
#define UDX(i, j) (uFld[S(i, j)] * G2(g->dxC, i, j))
  for (int j = 2 - OLy; j <= sNy + OLy; j++)
    for (int i = 2 - OLx; i <= sNx + OLx; i++)
      vort3[S(i, j)] = G2(g->recip_rAz, i, j) * ((VDY(i, j) - VDY(i - 1, j)) - (UDX(i, j) - UDX(i, j - 1)));
  if (csCorners & 1) {
    const int i = 1, j = 1;
    vort3[S(i, j)] = +G2(g->recip_rAz, i, j) * ((VDY(i, j) - UDX(i, j)) + UDX(i, j - 1));
  }
  if (csCorners & 2) {
    const int i = sNx + 1, j = 1;
    if (myFace == 2)
      vort3[S(i, j)] = +G2(g->recip_rAz, i, j) * ((-UDX(i, j) - VDY(i - 1, j)) + UDX(i, j - 1));
    else if (myFace == 4)
      vort3[S(i, j)] = +G2(g->recip_rAz, i, j) * ((-VDY(i - 1, j) + UDX(i, j - 1)) - UDX(i, j));
    else
      vort3[S(i, j)] = +G2(g->recip_rAz, i, j) * ((+UDX(i, j - 1) - UDX(i, j)) - VDY(i - 1, j));
  }
  if (csCorners & 8) {
    const int i = 1, j = sNy + 1;
    if (myFace == 1)
      vort3[S(i, j)] = +G2(g->recip_rAz, i, j) * ((+UDX(i, j - 1) + VDY(i, j)) - UDX(i, j));
    else if (myFace == 3)
      vort3[S(i, j)] = +G2(g->recip_rAz, i, j) * ((-UDX(i, j) + UDX(i, j - 1)) + VDY(i, j));
    else
      vort3[S(i, j)] = +G2(g->recip_rAz, i, j) * ((+VDY(i, j) - UDX(i, j)) + UDX(i, j - 1));
  }
  if (csCorners & 4) {
    const int i = sNx + 1, j = sNy + 1;
    if (myFace % 2 == 1)
      vort3[S(i, j)] = +G2(g->recip_rAz, i, j) * ((-UDX(i, j) - VDY(i - 1, j)) + UDX(i, j - 1));
    else
      vort3[S(i, j)] = +G2(g->recip_rAz, i, j) * ((+UDX(i, j - 1) - UDX(i, j)) - VDY(i - 1, j));
  }
#undef VDY
#undef UDX
}

/* MOM_VI_DEL2UV (pkg/mom_vecinv/mom_vi_del2uv.F:78-124); hDiv's facet corners are refilled in place */
static void mom_vi_del2uv(const og_grid *g, int bi, int bj, int k, double *hDiv, const double *vort3,
                          const double *hFacZ, double *del2u, double *del2v, int csCorners) {
  SETUP
  fill_cs_corner_tr(d, 1, csCorners, hDiv);
  for (int j = 2 - OLy; j <= sNy + OLy - 1; j++)
    for (int i = 2 - OLx; i <= sNx + OLx - 1; i++)
      del2u[S(i, j)] = ((hDiv[S(i, j)] - hDiv[S(i - 1, j)]) * G2(g->recip_dxC, i, j)
                        - G3(g->recip_hFacW, i, j, k)
                            * (hFacZ[S(i, j + 1)] * vort3[S(i, j + 1)] - hFacZ[S(i, j)] * vort3[S(i, j)])
                            * G2(g->recip_dyG, i, j))
                       * G3(g->maskW, i, j, k);
  fill_cs_corner_tr(d, 2, csCorners, hDiv);
  for (int j = 2 - OLy; j <= sNy + OLy - 1; j++)
    for (int i = 2 - OLx; i <= sNx + OLx - 1; i++)
      del2v[S(i, j)] = ((hDiv[S(i, j)] - hDiv[S(i, j - 1)]) * G2(g->recip_dyC, i, j)
                        + G3(g->recip_hFacS, i, j, k)
                            * (hFacZ[S(i + 1, j)] * vort3[S(i + 1, j)] - hFacZ[S(i, j)] * vort3[S(i, j)])
                            * G2(g->recip_dxG, i, j))
                       * G3(g->maskS, i, j, k);
}

/* MOM_VI_HDISSIP (pkg/mom_vecinv/mom_vi_hdissip.F:60-271), constant coefficients,
 * MOM_VI_ORIGINAL_VISCA4 and ISOTROPIC_COS_SCALING undefined */
static void mom_vi_hdissip(const og_grid *g, const og_params *p, int bi, int bj, int k, const double *hDiv,
                           const double *vort3, const double *dStar, const double *zStar, const double *hFacZ,
                           int harmonic, int biharmonic, double *uDissip, double *vDissip) {
  SETUP
  for (int j = 2 - OLy; j <= sNy + OLy - 1; j++)
    for (int i = 2 - OLx; i <= sNx + OLx - 1; i++) {
      const double cU = g->cosFacU[(j + OLy - 1) + offc], cV = g->cosFacV[(j + OLy - 1) + offc];
      if (harmonic) {
        const double Dim = hDiv[S(i, j - 1)], Dij = hDiv[S(i, j)], Dmj = hDiv[S(i - 1, j)];
        const double Zip = hFacZ[S(i, j + 1)] * vort3[S(i, j + 1)], Zij = hFacZ[S(i, j)] * vort3[S(i, j)],
                     Zpj = hFacZ[S(i + 1, j)] * vort3[S(i + 1, j)];
        const double uD2 = p->viscAhD * cU * (Dij - Dmj) * G2(g->recip_dxC, i, j)
                         - p->viscAhZ * G3(g->recip_hFacW, i, j, k) * (Zip - Zij) * G2(g->recip_dyG, i, j);
        const double vD2 = p->viscAhZ * G3(g->recip_hFacS, i, j, k) * cV * (Zpj - Zij) * G2(g->recip_dxG, i, j)
                         + p->viscAhD * (Dij - Dim) * G2(g->recip_dyC, i, j);
        uDissip[S(i, j)] = uD2 * G3(g->maskW, i, j, k);
        vDissip[S(i, j)] = vD2 * G3(g->maskS, i, j, k);
      } else {
        uDissip[S(i, j)] = 0.;
        vDissip[S(i, j)] = 0.;
      }
    }
  if (biharmonic) {
    for (int j = 2 - OLy; j <= sNy + OLy - 1; j++)
      for (int i = 2 - OLx; i <= sNx + OLx - 1; i++) {
        const double cU = g->cosFacU[(j + OLy - 1) + offc], cV = g->cosFacV[(j + OLy - 1) + offc];
        const double Dim = dStar[S(i, j - 1)], Dij = dStar[S(i, j)], Dmj = dStar[S(i - 1, j)];
        const double Zip = hFacZ[S(i, j + 1)] * zStar[S(i, j + 1)], Zij = hFacZ[S(i, j)] * zStar[S(i, j)],
                     Zpj = hFacZ[S(i + 1, j)] * zStar[S(i + 1, j)];
        double uD4 = p->viscA4D * cU * (Dij - Dmj) * G2(g->recip_dxC, i, j)
                   - p->viscA4Z * G3(g->recip_hFacW, i, j, k) * (Zip - Zij) * G2(g->recip_dyG, i, j);
        double vD4 = p->viscA4Z * G3(g->recip_hFacS, i, j, k) * cV * (Zpj - Zij) * G2(g->recip_dxG, i, j)
                   + p->viscA4D * (Dij - Dim) * G2(g->recip_dyC, i, j);
        uD4 = -uD4 * G3(g->maskW, i, j, k);
        vD4 = -vD4 * G3(g->maskS, i, j, k);
        uDissip[S(i, j)] = uDissip[S(i, j)] + uD4;
        vDissip[S(i, j)] = vDissip[S(i, j)] + vD4;
      }
  }
}

/* MOM_VI_CORIOLIS (pkg/mom_vecinv/mom_vi_coriolis.F:48-190) */
static void mom_vi_coriolis(const og_grid *g, const og_params *p, int bi, int bj, int k, const double *uFld,
                            const double *vFld, double *uCf, double *vCf) {
  SETUP
  const double epsil = 1e-9;
  const int sch = p->selectCoriScheme;
#define VDXH(i, j) (vFld[S(i, j)] * G2(g->dxG, i, j) * G3(g->hFacS, i, j, k))
#define UDYH(i, j) (uFld[S(i, j)] * G2(g->dyG, i, j) * G3(g->hFacW, i, j, k))
  for (int j = 1 - OLy; j <= sNy + OLy - 1; j++)
    for (int i = 2 - OLx; i <= sNx + OLx; i++) {
      const double f0 = G2(g->fCoriG, i, j), f1 = G2(g->fCoriG, i, j + 1);
      if (sch == 0) {
        const double vBarXY = 0.25 * ((vFld[S(i, j)] * G2(g->dxG, i, j) + vFld[S(i - 1, j)] * G2(g->dxG, i - 1, j))
                                    + (vFld[S(i, j + 1)] * G2(g->dxG, i, j + 1) + vFld[S(i - 1, j + 1)] * G2(g->dxG, i - 1, j + 1)));
        uCf[S(i, j)] = +0.5 * (f0 + f1) * vBarXY * G2(g->recip_dxC, i, j) * G3(g->maskW, i, j, k);
      } else if (sch == 1) {
        const double vBarXY = ((VDXH(i, j) + VDXH(i - 1, j)) + (VDXH(i, j + 1) + VDXH(i - 1, j + 1)))
            / fmax(epsil, (G3(g->hFacS, i, j, k) + G3(g->hFacS, i - 1, j, k))
                        + (G3(g->hFacS, i, j + 1, k) + G3(g->hFacS, i - 1, j + 1, k)));
        uCf[S(i, j)] = +0.5 * (f0 + f1) * vBarXY * G2(g->recip_dxC, i, j) * G3(g->maskW, i, j, k);
      } else if (sch == 2) {
        const double vBarXY = 0.25 * ((VDXH(i, j) + VDXH(i - 1, j)) + (VDXH(i, j + 1) + VDXH(i - 1, j + 1)));
        uCf[S(i, j)] = +0.5 * (f0 + f1) * vBarXY * G2(g->recip_dxC, i, j) * G3(g->recip_hFacW, i, j, k);
      } else {
        const double vBarXm = 0.5 * (VDXH(i, j) + VDXH(i - 1, j)), vBarXp = 0.5 * (VDXH(i, j + 1) + VDXH(i - 1, j + 1));
        uCf[S(i, j)] = +0.5 * (vBarXm * f0 + vBarXp * f1) * G2(g->recip_dxC, i, j) * G3(g->recip_hFacW, i, j, k);
      }
    }
  for (int j = 2 - OLy; j <= sNy + OLy; j++)
    for (int i = 1 - OLx; i <= sNx + OLx - 1; i++) {
      const double f0 = G2(g->fCoriG, i, j), f1 = G2(g->fCoriG, i + 1, j);
      if (sch == 0) {
        const double uBarXY = 0.25 * ((uFld[S(i, j)] * G2(g->dyG, i, j) + uFld[S(i, j - 1)] * G2(g->dyG, i, j - 1))
                                    + (uFld[S(i + 1, j)] * G2(g->dyG, i + 1, j) + uFld[S(i + 1, j - 1)] * G2(g->dyG, i + 1, j - 1)));
        vCf[S(i, j)] = -0.5 * (f0 + f1) * uBarXY * G2(g->recip_dyC, i, j) * G3(g->maskS, i, j, k);
      } else if (sch == 1) {
        const double uBarXY = ((UDYH(i, j) + UDYH(i, j - 1)) + (UDYH(i + 1, j) + UDYH(i + 1, j - 1)))
            / fmax(epsil, (G3(g->hFacW, i, j, k) + G3(g->hFacW, i, j - 1, k))
                        + (G3(g->hFacW, i + 1, j, k) + G3(g->hFacW, i + 1, j - 1, k)));
        vCf[S(i, j)] = -0.5 * (f0 + f1) * uBarXY * G2(g->recip_dyC, i, j) * G3(g->maskS, i, j, k);
      } else if (sch == 2) {
        const double uBarXY = 0.25 * ((UDYH(i, j) + UDYH(i, j - 1)) + (UDYH(i + 1, j) + UDYH(i + 1, j - 1)));
        vCf[S(i, j)] = -0.5 * (f0 + f1) * uBarXY * G2(g->recip_dyC, i, j) * G3(g->recip_hFacS, i, j, k);
      } else {
        const double uBarYm = 0.5 * (UDYH(i, j) + UDYH(i, j - 1)), uBarYp = 0.5 * (UDYH(i + 1, j) + UDYH(i + 1, j - 1));
        vCf[S(i, j)] = -0.5 * (uBarYm * f0 + uBarYp * f1) * G2(g->recip_dyC, i, j) * G3(g->recip_hFacS, i, j, k);
      }
    }
}

/* MOM_VI_U_CORIOLIS (pkg/mom_vecinv/mom_vi_u_coriolis.F:54-197), upwindVort3 = .FALSE. */
static void mom_vi_u_coriolis(const og_grid *g, const og_params *p, int bi, int bj, int k, const double *vFld,
                              const double *omega3, const double *hFacZ, const double *r_hFacZ, double *uCf) {
  SETUP
  const double epsil = 1e-9, oneThird = 1. / 3.;
  const int sch = p->selectVortScheme;
  for (int j = 1 - OLy; j <= sNy + OLy - 1; j++)
    for (int i = 2 - OLx; i <= sNx + OLx - (sch == 3 ? 1 : 0); i++) {
      if (sch == 0) {
        const double vBarXY = 0.25 * ((VDXH(i, j) + VDXH(i - 1, j)) + (VDXH(i, j + 1) + VDXH(i - 1, j + 1)));
        const double vort3u = 0.5 * (omega3[S(i, j)] * r_hFacZ[S(i, j)] + omega3[S(i, j + 1)] * r_hFacZ[S(i, j + 1)]);
        uCf[S(i, j)] = +vort3u * vBarXY * G2(g->recip_dxC, i, j) * G3(g->maskW, i, j, k);
      } else if (sch == 1) {
        const double vBarXY = 0.5 * ((vFld[S(i, j)] * G2(g->dxG, i, j) * hFacZ[S(i, j)]
                                    + vFld[S(i - 1, j)] * G2(g->dxG, i - 1, j) * hFacZ[S(i, j)])
                                   + (vFld[S(i, j + 1)] * G2(g->dxG, i, j + 1) * hFacZ[S(i, j + 1)]
                                    + vFld[S(i - 1, j + 1)] * G2(g->dxG, i - 1, j + 1) * hFacZ[S(i, j + 1)]))
                              / fmax(epsil, hFacZ[S(i, j)] + hFacZ[S(i, j + 1)]);
        const double vort3u = 0.5 * (omega3[S(i, j)] + omega3[S(i, j + 1)]);
        uCf[S(i, j)] = +vort3u * vBarXY * G2(g->recip_dxC, i, j) * G3(g->maskW, i, j, k);
      } else if (sch == 2) {
        const double vBarXm = 0.5 * (VDXH(i, j) + VDXH(i - 1, j)), vBarXp = 0.5 * (VDXH(i, j + 1) + VDXH(i - 1, j + 1));
        const double vort3u = (vBarXm * r_hFacZ[S(i, j)] * omega3[S(i, j)]
                             + vBarXp * r_hFacZ[S(i, j + 1)] * omega3[S(i, j + 1)]) * 0.5;
        uCf[S(i, j)] = +vort3u * G2(g->recip_dxC, i, j) * G3(g->maskW, i, j, k);
      } else {
#define RZ(i, j) (r_hFacZ[S(i, j)] * omega3[S(i, j)])
        const double vort3mj = (RZ(i, j) + (RZ(i, j + 1) + RZ(i - 1, j))) * oneThird * VDXH(i - 1, j);
        const double vort3ij = (RZ(i, j) + (RZ(i, j + 1) + RZ(i + 1, j))) * oneThird * VDXH(i, j);
        const double vort3mp = (RZ(i, j + 1) + (RZ(i, j) + RZ(i - 1, j + 1))) * oneThird * VDXH(i - 1, j + 1);
        const double vort3ip = (RZ(i, j + 1) + (RZ(i, j) + RZ(i + 1, j + 1))) * oneThird * VDXH(i, j + 1);
        uCf[S(i, j)] = +((vort3mj + vort3ij) + (vort3mp + vort3ip)) * 0.25 * G2(g->recip_dxC, i, j) * G3(g->maskW, i, j, k);
      }
    }
  if (p->useJamartMomAdv)
    for (int j = 1 - OLy; j <= sNy + OLy - 1; j++)
      for (int i = 2 - OLx; i <= sNx + OLx - 1; i++)
        uCf[S(i, j)] = uCf[S(i, j)] * 4. * G3(g->hFacW, i, j, k)
            / fmax(epsil, (G3(g->hFacS, i, j, k) + G3(g->hFacS, i - 1, j, k))
                        + (G3(g->hFacS, i, j + 1, k) + G3(g->hFacS, i - 1, j + 1, k)));
}

/* MOM_VI_V_CORIOLIS (pkg/mom_vecinv/mom_vi_v_coriolis.F:54-197), upwindVort3 = .FALSE. */
static void mom_vi_v_coriolis(const og_grid *g, const og_params *p, int bi, int bj, int k, const double *uFld,
                              const double *omega3, const double *hFacZ, const double *r_hFacZ, double *vCf) {
  SETUP
  const double epsil = 1e-9, oneThird = 1. / 3.;
  const int sch = p->selectVortScheme;
  for (int j = 2 - OLy; j <= sNy + OLy - (sch == 3 ? 1 : 0); j++)
    for (int i = 1 - OLx; i <= sNx + OLx - 1; i++) {
      if (sch == 0) {
        const double uBarXY = 0.25 * ((UDYH(i, j) + UDYH(i, j - 1)) + (UDYH(i + 1, j) + UDYH(i + 1, j - 1)));
        const double vort3v = 0.5 * (omega3[S(i, j)] * r_hFacZ[S(i, j)] + omega3[S(i + 1, j)] * r_hFacZ[S(i + 1, j)]);
        vCf[S(i, j)] = -vort3v * uBarXY * G2(g->recip_dyC, i, j) * G3(g->maskS, i, j, k);
      } else if (sch == 1) {
        const double uBarXY = 0.5 * ((uFld[S(i, j)] * G2(g->dyG, i, j) * hFacZ[S(i, j)]
                                    + uFld[S(i, j - 1)] * G2(g->dyG, i, j - 1) * hFacZ[S(i, j)])
                                   + (uFld[S(i + 1, j)] * G2(g->dyG, i + 1, j) * hFacZ[S(i + 1, j)]
                                    + uFld[S(i + 1, j - 1)] * G2(g->dyG, i + 1, j - 1) * hFacZ[S(i + 1, j)]))
                              / fmax(epsil, hFacZ[S(i, j)] + hFacZ[S(i + 1, j)]);
        const double vort3v = 0.5 * (omega3[S(i, j)] + omega3[S(i + 1, j)]);
        vCf[S(i, j)] = -vort3v * uBarXY * G2(g->recip_dyC, i, j) * G3(g->maskS, i, j, k);
      } else if (sch == 2) {
        const double uBarYm = 0.5 * (UDYH(i, j) + UDYH(i, j - 1)), uBarYp = 0.5 * (UDYH(i + 1, j) + UDYH(i + 1, j - 1));
        const double vort3v = (uBarYm * r_hFacZ[S(i, j)] * omega3[S(i, j)]
                             + uBarYp * r_hFacZ[S(i + 1, j)] * omega3[S(i + 1, j)]) * 0.5;
        vCf[S(i, j)] = -vort3v * G2(g->recip_dyC, i, j) * G3(g->maskS, i, j, k);
      } else {
        const double vort3im = (RZ(i, j) + (RZ(i + 1, j) + RZ(i, j - 1))) * oneThird * UDYH(i, j - 1);
        const double vort3ij = (RZ(i, j) + (RZ(i + 1, j) + RZ(i, j + 1))) * oneThird * UDYH(i, j);
        const double vort3pm = (RZ(i + 1, j) + (RZ(i, j) + RZ(i + 1, j - 1))) * oneThird * UDYH(i + 1, j - 1);
        const double vort3pj = (RZ(i + 1, j) + (RZ(i, j) + RZ(i + 1, j + 1))) * oneThird * UDYH(i + 1, j);
        vCf[S(i, j)] = -((vort3im + vort3ij) + (vort3pm + vort3pj)) * 0.25 * G2(g->recip_dyC, i, j) * G3(g->maskS, i, j, k);
      }
    }
  if (p->useJamartMomAdv)
    for (int j = 2 - OLy; j <= sNy + OLy - 1; j++)
      for (int i = 1 - OLx; i <= sNx + OLx - 1; i++)
        vCf[S(i, j)] = vCf[S(i, j)] * 4. * G3(g->hFacS, i, j, k)
            / fmax(epsil, (G3(g->hFacW, i, j, k) + G3(g->hFacW, i, j - 1, k))
                        + (G3(g->hFacW, i + 1, j, k) + G3(g->hFacW, i + 1, j - 1, k)));
}
#undef RZ
#undef VDXH
#undef UDYH

/* MOM_VI_U_CORIOLIS_C4 / MOM_VI_V_CORIOLIS_C4 (pkg/mom_vecinv/mom_vi_{u,v}_coriolis_c4.F:60-215): 4th-order
 * (fourthVort3 = .TRUE.) or upwind interpolation of the vorticity; selectVortScheme 0 and 2 only; written on
 * U: i = 1..sNx+1, j = 1..sNy, V: i = 1..sNx, j = 1..sNy+1 -- elsewhere the output array keeps what it held. */
static int mom_vi_coriolis_c4(const og_grid *g, const og_params *p, int bi, int bj, int k, int isV, const double *fld,
                              const double *omega3, const double *r_hFacZ, double *cf, int csCorners) {
  SETUP
  const double oneSixth = 1. / 6., oneTwelve = 1. / 12.;
  const int sch = p->selectVortScheme, upw = p->upwindVorticity;
  if (sch != 0 && sch != 2) return 1;
  double *v3 = (double *)calloc(px * py, sizeof(double));
  FORALL v3[S(i, j)] = r_hFacZ[S(i, j)] * omega3[S(i, j)];
  if (csCorners && p->highOrderVorticity) {
    if (!isV) {
      if (csCorners & 1) v3[S(1, 0)] = (v3[S(1, 0)] + v3[S(2, 1)]) * 0.5;
      if (csCorners & 2) v3[S(sNx + 1, 0)] = (v3[S(sNx + 1, 0)] + v3[S(sNx, 1)]) * 0.5;
      if (csCorners & 8) v3[S(1, sNy + 2)] = (v3[S(1, sNy + 2)] + v3[S(2, sNy + 1)]) * 0.5;
      if (csCorners & 4) v3[S(sNx + 1, sNy + 2)] = (v3[S(sNx + 1, sNy + 2)] + v3[S(sNx, sNy + 1)]) * 0.5;
    } else {
      if (csCorners & 1) v3[S(0, 1)] = (v3[S(0, 1)] + v3[S(1, 2)]) * 0.5;
      if (csCorners & 2) v3[S(sNx + 2, 1)] = (v3[S(sNx + 2, 1)] + v3[S(sNx + 1, 2)]) * 0.5;
      if (csCorners & 8) v3[S(0, sNy + 1)] = (v3[S(0, sNy + 1)] + v3[S(1, sNy)]) * 0.5;
      if (csCorners & 4) v3[S(sNx + 2, sNy + 1)] = (v3[S(sNx + 2, sNy + 1)] + v3[S(sNx + 1, sNy)]) * 0.5;
    }
  }
  const int di = isV ? 1 : 0, dj = isV ? 0 : 1;          /* direction along which the vorticity is interpolated */
  const int iHi = isV ? sNx : sNx + 1, jHi = isV ? sNy + 1 : sNy;
  for (int j = 1; j <= jHi; j++)
    for (int i = 1; i <= iHi; i++) {
      double bm, bp;     /* transports on the two sides (vBarXm/p or uBarYm/p) */
      if (!isV) {
        bm = G3(fld, i, j, k) * G2(g->dxG, i, j) * G3(g->hFacS, i, j, k) + G3(fld, i - 1, j, k) * G2(g->dxG, i - 1, j) * G3(g->hFacS, i - 1, j, k);
        bp = G3(fld, i, j + 1, k) * G2(g->dxG, i, j + 1) * G3(g->hFacS, i, j + 1, k)
           + G3(fld, i - 1, j + 1, k) * G2(g->dxG, i - 1, j + 1) * G3(g->hFacS, i - 1, j + 1, k);
      } else {
        bm = G3(fld, i, j, k) * G2(g->dyG, i, j) * G3(g->hFacW, i, j, k) + G3(fld, i, j - 1, k) * G2(g->dyG, i, j - 1) * G3(g->hFacW, i, j - 1, k);
        bp = G3(fld, i + 1, j, k) * G2(g->dyG, i + 1, j) * G3(g->hFacW, i + 1, j, k)
           + G3(fld, i + 1, j - 1, k) * G2(g->dyG, i + 1, j - 1) * G3(g->hFacW, i + 1, j - 1, k);
      }
      const double z0 = v3[S(i, j)], z1 = v3[S(i + di, j + dj)], zm = v3[S(i - di, j - dj)], z2 = v3[S(i + 2 * di, j + 2 * dj)];
      const double rd = isV ? G2(g->recip_dyC, i, j) : G2(g->recip_dxC, i, j);
      const double mk = isV ? G3(g->maskS, i, j, k) : G3(g->maskW, i, j, k);
      const double sgn = isV ? -1. : 1.;
      if (sch == 0) {
        const double bXY = 0.25 * (bm + bp);
        double vort;
        if (upw) vort = bXY > 0. ? z0 : z1;
        else {
          const double Rjp = z2 - z1, Rjm = z0 - zm;
          vort = 0.5 * ((z0 + z1) - oneTwelve * (Rjp - Rjm));
        }
        cf[S(i, j)] = isV ? -vort * bXY * rd * mk : vort * bXY * rd * mk;
      } else {
        const double bM = 0.5 * bm, bP = 0.5 * bp;
        double vort;
        if (upw) vort = (bM + bP) > 0. ? bM * z0 : bP * z1;
        else {
          double Rjp = z2 - z1, Rjm = z0 - zm;
          const double Rj = z1 - z0;
          Rjp = z1 - oneSixth * (Rjp - Rj);
          Rjm = z0 - oneSixth * (Rj - Rjm);
          vort = 0.5 * (bM * Rjm + bP * Rjp);
        }
        cf[S(i, j)] = isV ? -vort * rd * mk : vort * rd * mk;
      }
      (void)sgn;
    }
  free(v3);
  return 0;
}

/* MOM_VI_U_VERTSHEAR / MOM_VI_V_VERTSHEAR (pkg/mom_vecinv/mom_vi_{u,v}_vertshear.F:41-133) */
static void mom_vi_vertshear(const og_grid *g, const og_params *p, int bi, int bj, int k, int isV,
                             const double *fld, const double *wVel, double *shear) {
  SETUP
  const int rAdvAreaWeight = !(p->selectKEscheme == 1 || p->selectKEscheme == 3);
  const int Kp1 = (k + 1 < Nr) ? k + 1 : Nr, Km1 = (k - 1 > 1) ? k - 1 : 1;
  const double mask_Kp1 = (k == Nr) ? 0. : 1., mask_Km1 = (k == 1) ? 0. : 1.;
  const int di = isV ? 0 : 1, dj = isV ? 1 : 0;
  const double *rrA = isV ? g->recip_rAs : g->recip_rAw;
  const double *rh = isV ? g->recip_hFacS : g->recip_hFacW;
  for (int j = 1 - OLy + dj; j <= sNy + OLy; j++)
    for (int i = 1 - OLx + di; i <= sNx + OLx; i++) {
      double wBm, wBp;
      if (rAdvAreaWeight) {
        wBm = 0.5 * (G3(wVel, i, j, k) * G2(g->rA, i, j) * G3(g->maskC, i, j, Km1)
                   + G3(wVel, i - di, j - dj, k) * G2(g->rA, i - di, j - dj) * G3(g->maskC, i - di, j - dj, Km1))
              * mask_Km1 * G2(rrA, i, j);
        wBp = 0.5 * (G3(wVel, i, j, Kp1) * G2(g->rA, i, j) + G3(wVel, i - di, j - dj, Kp1) * G2(g->rA, i - di, j - dj))
              * mask_Kp1 * G2(rrA, i, j);
      } else {
        wBm = 0.5 * (G3(wVel, i, j, k) * G3(g->maskC, i, j, Km1)
                   + G3(wVel, i - di, j - dj, k) * G3(g->maskC, i - di, j - dj, Km1)) * mask_Km1;
        wBp = 0.5 * (G3(wVel, i, j, Kp1) + G3(wVel, i - di, j - dj, Kp1)) * mask_Kp1;
      }
      const double fZm = (G3(fld, i, j, k) - mask_Km1 * G3(fld, i, j, Km1)) * p->rkSign;
      const double fZp = (mask_Kp1 * G3(fld, i, j, Kp1) - G3(fld, i, j, k)) * p->rkSign;
      if (p->upwindShear)
        shear[S(i, j)] = -0.5 * ((wBp * fZp + wBm * fZm) + (fabs(wBp) * fZp - fabs(wBm) * fZm))
                         * G3(rh, i, j, k) * g->recip_drF[k - 1];
      else
        shear[S(i, j)] = -0.5 * (wBp * fZp + wBm * fZm) * G3(rh, i, j, k) * g->recip_drF[k - 1];
    }
}

int og_mom_vecinv(const og_grid *g, const og_params *p, int bi, int bj, int k,
                  int iMin, int iMax, int jMin, int jMax,
                  const double *kappaRU, const double *kappaRV,
                  const double *fVerUkm, const double *fVerVkm, double *fVerUkp, double *fVerVkp,
                  double *guDiss, double *gvDiss,
                  const double *uVel, const double *vVel, const double *wVel,
                  double *gU, double *gV, int csCorners, int myFace) {
  SETUP
  if (p->momImplVertAdv) return 1;
  const int c4 = p->highOrderVorticity || p->upwindVorticity;
  if (c4 && p->selectVortScheme != 0 && p->selectVortScheme != 2) return 3;
  if (p->selectVortScheme < 0 || p->selectVortScheme > 3 || p->selectCoriScheme < 0 || p->selectCoriScheme > 3) return 2;
  const size_t ns = px * py;
  double *buf = (double *)calloc(ns * 19, sizeof(double));
  double *vF = buf, *vrF = buf + ns, *uCf = buf + 2 * ns, *vCf = buf + 3 * ns, *del2u = buf + 4 * ns,
         *del2v = buf + 5 * ns, *dStar = buf + 6 * ns, *zStar = buf + 7 * ns, *hFacZ = buf + 8 * ns,
         *h0FacZ = buf + 9 * ns, *r_hFacZ = buf + 10 * ns, *uFld = buf + 11 * ns, *vFld = buf + 12 * ns,
         *cDrag = buf + 13 * ns, *KE = buf + 14 * ns, *omega3 = buf + 15 * ns, *vort3 = buf + 16 * ns,
         *vort3BC = buf + 17 * ns, *hDiv = buf + 18 * ns;
  /* mom_vecinv.F:196-250 */
  FORALL { guDiss[S(i, j)] = 0.; gvDiss[S(i, j)] = 0.; }
  const double ArDudrFac = p->vfFacMom * 1., ArDvdrFac = p->vfFacMom * 1.;
  const double sideMaskFac = p->no_slip_sides ? p->sideDragFactor : 0.;
  const int bottomDragTerms = (p->selectImplicitDrag == 0 &&
      (p->no_slip_bottom || p->selectBotDragQuadr >= 0 || p->bottomDragLinear != 0.));
  const int harmonic = (p->viscAhD != 0. || p->viscAhZ != 0.);     /* useHarmonicVisc, ini_parms / set_parms */
  const int biharmonic = p->useBiharmonicVisc;

  mom_calc_hfacz(g, bi, bj, k, hFacZ, r_hFacZ);                    /* :266 */
  FORALL { uFld[S(i, j)] = G3(uVel, i, j, k); vFld[S(i, j)] = G3(vVel, i, j, k); }
  mom_calc_ke(g, bi, bj, k, p->selectKEscheme, uFld, vFld, KE);    /* :287 */
  mom_calc_relvort3(g, bi, bj, k, uFld, vFld, vort3, csCorners, myFace);
  FORALL {                                                         /* :292-300 */
    vort3BC[S(i, j)] = vort3[S(i, j)];
    if (hFacZ[S(i, j)] == 0.) { vort3BC[S(i, j)] = sideMaskFac * vort3BC[S(i, j)]; vort3[S(i, j)] = 0.; }
  }

  if (p->momViscosity) {                                           /* :308-663 */
    FORALL h0FacZ[S(i, j)] = hFacZ[S(i, j)];
    mom_calc_hdiv2(g, bi, bj, k, uFld, vFld, hDiv);
    if (biharmonic) {
      mom_vi_del2uv(g, bi, bj, k, hDiv, vort3, hFacZ, del2u, del2v, csCorners);
      mom_calc_hdiv2(g, bi, bj, k, del2u, del2v, dStar);
      mom_calc_relvort3(g, bi, bj, k, del2u, del2v, zStar, csCorners, myFace);
    }
    mom_vi_hdissip(g, p, bi, bj, k, hDiv, vort3, dStar, zStar, hFacZ, harmonic, biharmonic, guDiss, gvDiss);
    /* ---- U */
    if (!p->implicitViscosity) {
      mom_rviscflux(g, p, bi, bj, k + 1, 0, uVel, kappaRU, vrF);
      for (int j = jMin; j <= jMax; j++)
        for (int i = iMin; i <= iMax; i++) fVerUkp[S(i, j)] = ArDudrFac * vrF[S(i, j)];
      for (int j = jMin; j <= jMax; j++)
        for (int i = iMin; i <= iMax; i++)
          guDiss[S(i, j)] = guDiss[S(i, j)] - G3(g->recip_hFacW, i, j, k) * g->recip_drF[k - 1] * G2(g->recip_rAw, i, j)
                                                  * (fVerUkp[S(i, j)] - fVerUkm[S(i, j)]) * p->rkSign;
    }
    if (p->no_slip_sides) {
      mom_sidedrag(g, p, bi, bj, k, 0, uFld, del2u, h0FacZ, vF);
      for (int j = jMin; j <= jMax; j++)
        for (int i = iMin; i <= iMax; i++) guDiss[S(i, j)] = guDiss[S(i, j)] + vF[S(i, j)];
    }
    if (bottomDragTerms) {
      mom_botdrag_coeff(g, p, bi, bj, k, 0, uFld, vFld, kappaRU, KE, cDrag);
      for (int j = jMin; j <= jMax; j++)
        for (int i = iMin; i <= iMax; i++) {
          vF[S(i, j)] = -cDrag[S(i, j)] * uFld[S(i, j)] * G3(g->recip_hFacW, i, j, k) * g->recip_drF[k - 1];
          guDiss[S(i, j)] = guDiss[S(i, j)] + vF[S(i, j)];
        }
    }
    /* ---- V */
    if (!p->implicitViscosity) {
      mom_rviscflux(g, p, bi, bj, k + 1, 1, vVel, kappaRV, vrF);
      for (int j = jMin; j <= jMax; j++)
        for (int i = iMin; i <= iMax; i++) fVerVkp[S(i, j)] = ArDvdrFac * vrF[S(i, j)];
      for (int j = jMin; j <= jMax; j++)
        for (int i = iMin; i <= iMax; i++)
          gvDiss[S(i, j)] = gvDiss[S(i, j)] - G3(g->recip_hFacS, i, j, k) * g->recip_drF[k - 1] * G2(g->recip_rAs, i, j)
                                                  * (fVerVkp[S(i, j)] - fVerVkm[S(i, j)]) * p->rkSign;
    }
    if (p->no_slip_sides) {
      mom_sidedrag(g, p, bi, bj, k, 1, vFld, del2v, h0FacZ, vF);
      for (int j = jMin; j <= jMax; j++)
        for (int i = iMin; i <= iMax; i++) gvDiss[S(i, j)] = gvDiss[S(i, j)] + vF[S(i, j)];
    }
    if (bottomDragTerms) {
      mom_botdrag_coeff(g, p, bi, bj, k, 1, uFld, vFld, kappaRV, KE, cDrag);
      for (int j = jMin; j <= jMax; j++)
        for (int i = iMin; i <= iMax; i++) {
          vF[S(i, j)] = -cDrag[S(i, j)] * vFld[S(i, j)] * G3(g->recip_hFacS, i, j, k) * g->recip_drF[k - 1];
          gvDiss[S(i, j)] = gvDiss[S(i, j)] + vF[S(i, j)];
        }
    }
  }

  /* :672-673 MOM_CALC_ABSVORT3 */
  if (p->useAbsVorticity) {
    const double nonLinFac = p->momAdvection ? 1. : 0., useCoriolisFac = p->useCoriolis ? 1. : 0.;
    FORALL omega3[S(i, j)] = G2(g->fCoriG, i, j) * useCoriolisFac + vort3[S(i, j)] * nonLinFac;
  }
  /* :683-738 Coriolis */
  if (p->useCoriolis && !(p->useCDscheme || (p->useAbsVorticity && p->momAdvection))) {
    if (p->useAbsVorticity) {
      mom_vi_u_coriolis(g, p, bi, bj, k, vFld, omega3, hFacZ, r_hFacZ, uCf);
      mom_vi_v_coriolis(g, p, bi, bj, k, uFld, omega3, hFacZ, r_hFacZ, vCf);
    } else {
      mom_vi_coriolis(g, p, bi, bj, k, uFld, vFld, uCf, vCf);
    }
    for (int j = jMin; j <= jMax; j++)
      for (int i = iMin; i <= iMax; i++) { G3(gU, i, j, k) = uCf[S(i, j)]; G3(gV, i, j, k) = vCf[S(i, j)]; }
  } else {
    for (int j = jMin; j <= jMax; j++)
      for (int i = iMin; i <= iMax; i++) { G3(gU, i, j, k) = 0.; G3(gV, i, j, k) = 0.; }
  }
  /* :745-884 advection */
  if (p->momAdvection) {
    const double *w3 = p->useAbsVorticity ? omega3 : vort3;
    if (c4) mom_vi_coriolis_c4(g, p, bi, bj, k, 0, vVel, w3, r_hFacZ, uCf, csCorners);      /* :746-757 */
    else mom_vi_u_coriolis(g, p, bi, bj, k, vFld, w3, hFacZ, r_hFacZ, uCf);
    for (int j = jMin; j <= jMax; j++)
      for (int i = iMin; i <= iMax; i++) G3(gU, i, j, k) = G3(gU, i, j, k) + uCf[S(i, j)];
    if (c4) mom_vi_coriolis_c4(g, p, bi, bj, k, 1, uVel, w3, r_hFacZ, vCf, csCorners);
    else mom_vi_v_coriolis(g, p, bi, bj, k, uFld, w3, hFacZ, r_hFacZ, vCf);
    for (int j = jMin; j <= jMax; j++)
      for (int i = iMin; i <= iMax; i++) G3(gV, i, j, k) = G3(gV, i, j, k) + vCf[S(i, j)];
    mom_vi_vertshear(g, p, bi, bj, k, 0, uVel, wVel, uCf);
    for (int j = jMin; j <= jMax; j++)
      for (int i = iMin; i <= iMax; i++) G3(gU, i, j, k) = G3(gU, i, j, k) + uCf[S(i, j)];
    mom_vi_vertshear(g, p, bi, bj, k, 1, vVel, wVel, vCf);
    for (int j = jMin; j <= jMax; j++)
      for (int i = iMin; i <= iMax; i++) G3(gV, i, j, k) = G3(gV, i, j, k) + vCf[S(i, j)];
    /* MOM_VI_{U,V}_GRAD_KE */
    for (int j = 1 - OLy; j <= sNy + OLy; j++)
      for (int i = 2 - OLx; i <= sNx + OLx; i++)
        uCf[S(i, j)] = -G2(g->recip_dxC, i, j) * (KE[S(i, j)] - KE[S(i - 1, j)]) * G3(g->maskW, i, j, k);
    for (int j = jMin; j <= jMax; j++)
      for (int i = iMin; i <= iMax; i++) G3(gU, i, j, k) = G3(gU, i, j, k) + uCf[S(i, j)];
    for (int j = 2 - OLy; j <= sNy + OLy; j++)
      for (int i = 1 - OLx; i <= sNx + OLx; i++)
        vCf[S(i, j)] = -G2(g->recip_dyC, i, j) * (KE[S(i, j)] - KE[S(i, j - 1)]) * G3(g->maskS, i, j, k);
    for (int j = jMin; j <= jMax; j++)
      for (int i = iMin; i <= iMax; i++) G3(gV, i, j, k) = G3(gV, i, j, k) + vCf[S(i, j)];
  }
  /* :922-927 */
  for (int j = jMin; j <= jMax; j++)
    for (int i = iMin; i <= iMax; i++) {
      G3(gU, i, j, k) = G3(gU, i, j, k) * G3(g->maskW, i, j, k);
      G3(gV, i, j, k) = G3(gV, i, j, k) * G3(g->maskS, i, j, k);
    }
  free(buf);
  return 0;
}
