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
