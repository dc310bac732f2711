/* phys_oracle.c -- CPU restatement of the physics glue that verification/
 * tutorial_baroclinic_gyre (config 2) needs around the hot path, so that the
 * oracle's GAD_CALC_RHS and the Nr > 1 branches of MOM_FLUXFORM can be PINNED
 * against that experiment's golden output (results/output.txt).
 * TEST INFRASTRUCTURE ONLY.
 *
 * Restated, each for one tile:
 *   FIND_RHO_2D, LINEAR branch            model/src/find_rho.F (FIND_RHO_2D, 'LINEAR')
 *   density + GRAD_SIGMA(sigmaR) + CALC_IVDC sequence of DO_OCEANIC_PHYS
 *                                         model/src/do_oceanic_phys.F, grad_sigma.F, calc_ivdc.F
 *   FORCING_SURF_RELAX (theta)            model/src/forcing_surf_relax.F
 *   APPLY_FORCING_T (surface term)        model/src/apply_forcing.F (APPLY_FORCING_T, k = kSurface)
 *   CALC_3D_DIFFUSIVITY (no KPP/GM/BL79)  model/src/calc_3d_diffusivity.F
 *   GAD_IMPLICIT_R (diffusion only) + SOLVE_TRIDIAGONAL (default branch)
 *                                         pkg/generic_advdiff/gad_implicit_r.F, model/src/solve_tridiagonal.F
 *   CALC_PHI_HYD ('OCEANIC', integr_GeoPot = 2, uniformFreeSurfLev) + CALC_GRAD_PHI_HYD
 *                                         model/src/calc_phi_hyd.F, calc_grad_phi_hyd.F
 *   INTEGR_CONTINUITY, exactConserv part + UPDATE_ETAH (implicDiv2DFlow = 1)
 *                                         model/src/integr_continuity.F, update_etah.F
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "mitgcm_oracle.h"
#include "glue_oracle.h"

#define S(i, j) ((size_t)((i) + OLx - 1) + (size_t)px * (size_t)((j) + OLy - 1))
#define G2(a, i, j) (a)[S(i, j) + off2]
#define G3(a, i, j, k) (a)[S(i, j) + (size_t)px * py * (size_t)((k)-1) + off3]
#define L3(a, i, j, k) (a)[S(i, j) + (size_t)px * py * (size_t)((k)-1)]
#define FORALL for (int j = 1 - OLy; j <= sNy + OLy; j++) for (int i = 1 - OLx; i <= sNx + OLx; i++)
#define SETUP                                                                          \
  const og_dims *d = &g->d;                                                            \
  const int sNx = d->sNx, sNy = d->sNy, OLx = d->OLx, OLy = d->OLy, Nr = d->Nr;        \
  const size_t px = (size_t)(sNx + 2 * OLx), py = (size_t)(sNy + 2 * OLy);             \
  const size_t tile = (size_t)(bi - 1) + (size_t)d->nSx * (size_t)(bj - 1);            \
  const size_t off2 = px * py * tile, off3 = px * py * (size_t)Nr * tile;              \
  (void)Nr; (void)off2; (void)off3;

/* FIND_RHO_2D, equationOfState = 'LINEAR' (find_rho.F): tFld, sFld, rhoLoc are slabs. */
static void find_rho_2d_linear(const og_dims *d, int iMin, int iMax, int jMin, int jMax,
                               const double *tFld, const double *sFld, double refTemp, double refSalt,
                               const og_eos *e, double *rhoLoc) {
  const int OLx = d->OLx, OLy = d->OLy;
  const size_t px = (size_t)(d->sNx + 2 * OLx);
  const double dRho = e->rhoNil - e->rhoConst;
  for (int j = jMin; j <= jMax; j++)
    for (int i = iMin; i <= iMax; i++)
      rhoLoc[S(i, j)] = e->rhoNil * (e->sBeta * (sFld[S(i, j)] - refSalt) - e->tAlpha * (tFld[S(i, j)] - refTemp)) + dRho;
}

/* DO_OCEANIC_PHYS, density part for one tile: rhoInSitu(k) = FIND_RHO_2D(theta(k), salt(k), kRef = k)
 * for k = 1..Nr; then for k = Nr..2: rhoKp1 = rhoInSitu(k), rhoKm1 = FIND_RHO_2D(theta(k-1),
 * salt(k-1), kRef = k), sigmaR(k) = maskC(k) maskC(k-1) recip_drC(k) rkSign (rhoKp1 - rhoKm1)
 * (grad_sigma.F), IVDConvCount(k) = 1 where -sigmaR*gravitySign > 0 (calc_ivdc.F); level 1 stays 0.
 * theta, salt, rhoInSitu, IVDConvCount are tile3d arrays. */
void og_density_ivdc(const og_grid *g, const og_params *p, const og_eos *e, int bi, int bj,
                     const double *theta, const double *salt, const double *tRef, const double *sRef,
                     double *rhoInSitu, double *IVDConvCount) {
  SETUP
  const size_t ns = px * py;
  const int iMin = 1 - OLx, iMax = sNx + OLx, jMin = 1 - OLy, jMax = sNy + OLy;
  double *rhoKm1 = (double *)calloc(ns, sizeof(double));
  for (int k = 1; k <= Nr; k++) {
    FORALL { G3(rhoInSitu, i, j, k) = 0.; G3(IVDConvCount, i, j, k) = 0.; }
    find_rho_2d_linear(d, iMin, iMax, jMin, jMax, &G3(theta, 1 - OLx, 1 - OLy, k), &G3(salt, 1 - OLx, 1 - OLy, k),
                       tRef[k - 1], sRef[k - 1], e, &G3(rhoInSitu, 1 - OLx, 1 - OLy, k));
  }
  const double gravitySign = -1.; /* z coordinates, ini_vertical_grid.F:54 */
  for (int k = Nr; k >= 2; k--) {
    find_rho_2d_linear(d, iMin, iMax, jMin, jMax, &G3(theta, 1 - OLx, 1 - OLy, k - 1),
                       &G3(salt, 1 - OLx, 1 - OLy, k - 1), tRef[k - 1], sRef[k - 1], e, rhoKm1);
    FORALL {
      const double sigmaR = G3(g->maskC, i, j, k) * G3(g->maskC, i, j, k - 1) * g->recip_drC[k - 1] * p->rkSign
                            * (G3(rhoInSitu, i, j, k) - rhoKm1[S(i, j)]);
      G3(IVDConvCount, i, j, k) = (-sigmaR * gravitySign > 0.) ? 1. : 0.;
    }
  }
  free(rhoKm1);
}

/* FORCING_SURF_RELAX (theta only) followed by the surfaceForcingT line of EXTERNAL_FORCING_SURF
 * with Qnet = Qsw = 0: surfaceForcingT = -lambda (theta(ks) - SST) drF(ks) hFacC(ks), ks = 1. */
void og_forcing_surf_relax_T(const og_grid *g, int bi, int bj, const double *theta, const double *SST,
                             const double *lambdaThetaClimRelax, double recip_Cp, double mass2rUnit,
                             double *surfaceForcingT) {
  SETUP
  const int ks = 1;
  FORALL {
    G2(surfaceForcingT, i, j) = -G2(lambdaThetaClimRelax, i, j) * (G3(theta, i, j, ks) - G2(SST, i, j))
                                * g->drF[ks - 1] * G3(g->hFacC, i, j, ks);
    G2(surfaceForcingT, i, j) = G2(surfaceForcingT, i, j) - (0. - 0.) * recip_Cp * mass2rUnit;
  }
}

/* APPLY_FORCING_T at k = kSurface = 1: gtForc (slab) += surfaceForcingT recip_drF(k) recip_hFacC(k)
 * on 0..sNx+1, 0..sNy+1. */
void og_apply_forcing_T(const og_grid *g, int bi, int bj, int k, const double *surfaceForcingT, double *gtForc) {
  SETUP
  if (k != 1) return;
  for (int j = 0; j <= sNy + 1; j++)
    for (int i = 0; i <= sNx + 1; i++)
      gtForc[S(i, j)] = gtForc[S(i, j)] + G2(surfaceForcingT, i, j) * g->recip_drF[k - 1] * G3(g->recip_hFacC, i, j, k);
}

/* CALC_3D_DIFFUSIVITY for temperature without KPP / GM / 3-D diffKr:
 * KappaRTr(k) = IVDConvCount(k)*ivdc_kappa + KbryanLewis79(k) + diffKrNrT(k); kappaRk is a
 * per-tile (slab, Nr) local array. */
void og_calc_3d_diffusivity(const og_grid *g, int bi, int bj, const double *IVDConvCount, double ivdc_kappa,
                            const double *KbryanLewis79, const double *diffKrNrT, double *kappaRk) {
  SETUP
  for (int k = 1; k <= Nr; k++)
    FORALL L3(kappaRk, i, j, k) = G3(IVDConvCount, i, j, k) * ivdc_kappa + KbryanLewis79[k - 1];
  for (int k = 1; k <= Nr; k++)
    FORALL L3(kappaRk, i, j, k) = L3(kappaRk, i, j, k) + diffKrNrT[k - 1];
}

/* SOLVE_TRIDIAGONAL, default branch (neither SOLVE_DIAGONAL_LOWMEMORY nor _KINNER):
 * whole slab; a3d, b3d, c3d, y3d are (slab, Nr). Returns errCode. */
static int solve_tridiagonal(const og_dims *d, const double *a3d, const double *b3d, const double *c3d, double *y3d) {
  const int sNx = d->sNx, sNy = d->sNy, OLx = d->OLx, OLy = d->OLy, Nr = d->Nr;
  const size_t px = (size_t)(sNx + 2 * OLx), py = (size_t)(sNy + 2 * OLy);
  const size_t n = px * py * (size_t)Nr;
  int errCode = 0;
  double *c_prime = (double *)calloc(3 * n, sizeof(double));
  double *y_prime = c_prime + n, *y_m1 = c_prime + 2 * n;
  memcpy(y_m1, y3d, n * sizeof(double));
  for (int k = 1; k <= Nr; k++) {
    if (k == 1) {
      FORALL {
        if (L3(b3d, i, j, 1) != 0.) {
          const double recVar = 1. / L3(b3d, i, j, 1);
          L3(c_prime, i, j, 1) = L3(c3d, i, j, 1) * recVar;
          L3(y_prime, i, j, 1) = L3(y_m1, i, j, 1) * recVar;
        } else { L3(c_prime, i, j, 1) = 0.; L3(y_prime, i, j, 1) = 0.; errCode = 1; }
      }
    } else {
      FORALL {
        const double tmpVar = L3(b3d, i, j, k) - L3(a3d, i, j, k) * L3(c_prime, i, j, k - 1);
        if (tmpVar != 0.) {
          const double recVar = 1. / tmpVar;
          L3(c_prime, i, j, k) = L3(c3d, i, j, k) * recVar;
          L3(y_prime, i, j, k) = (L3(y_m1, i, j, k) - L3(a3d, i, j, k) * L3(y_prime, i, j, k - 1)) * recVar;
        } else { L3(c_prime, i, j, k) = 0.; L3(y_prime, i, j, k) = 0.; errCode = 1; }
      }
    }
  }
  for (int k = Nr; k >= 1; k--) {
    if (k == Nr) { FORALL L3(y3d, i, j, k) = L3(y_prime, i, j, k); }
    else { FORALL L3(y3d, i, j, k) = L3(y_prime, i, j, k) - L3(c_prime, i, j, k) * L3(y3d, i, j, k + 1); }
  }
  free(c_prime);
  return errCode;
}

/* GAD_IMPLICIT_R with implicitDiffusion = T, implicitAdvection = F (gad_implicit_r.F):
 * b5d(k) = -dT(k) maskC(k-1) recip_hFac(k) recip_drF(k) kappaRX(k) recip_drC(k),
 * d5d(k) = -dT(k) maskC(k+1) recip_hFac(k) recip_drF(k) kappaRX(k+1) recip_drC(k+1),
 * c5d = 1 - (b5d + d5d), on iMin..iMax x jMin..jMax (identity elsewhere); SOLVE_TRIDIAGONAL.
 * kappaRX, recip_hFac, gTracer are per-tile (slab, Nr). Returns the solver's errCode. */
int og_gad_implicit_r(const og_grid *g, int bi, int bj, int iMin, int iMax, int jMin, int jMax,
                      const double *deltaTLev, const double *kappaRX, const double *recip_hFac, double *gTracer) {
  SETUP
  if (Nr <= 1) return 0;
  const size_t n = px * py * (size_t)Nr;
  double *b5d = (double *)calloc(3 * n, sizeof(double));
  double *c5d = b5d + n, *d5d = b5d + 2 * n;
  for (size_t q = 0; q < n; q++) c5d[q] = 1.;
  for (int k = 2; k <= Nr; k++)
    for (int j = jMin; j <= jMax; j++)
      for (int i = iMin; i <= iMax; i++)
        L3(b5d, i, j, k) = -deltaTLev[k - 1] * G3(g->maskC, i, j, k - 1) * L3(recip_hFac, i, j, k) * g->recip_drF[k - 1]
                           * L3(kappaRX, i, j, k) * g->recip_drC[k - 1];
  for (int k = 1; k <= Nr - 1; k++)
    for (int j = jMin; j <= jMax; j++)
      for (int i = iMin; i <= iMax; i++)
        L3(d5d, i, j, k) = -deltaTLev[k - 1] * G3(g->maskC, i, j, k + 1) * L3(recip_hFac, i, j, k) * g->recip_drF[k - 1]
                           * L3(kappaRX, i, j, k + 1) * g->recip_drC[k];
  for (int k = 1; k <= Nr; k++)
    for (int j = jMin; j <= jMax; j++)
      for (int i = iMin; i <= iMax; i++)
        L3(c5d, i, j, k) = 1. - (L3(b5d, i, j, k) + L3(d5d, i, j, k));
  const int err = solve_tridiagonal(d, b5d, c5d, d5d, gTracer);
  free(b5d);
  return err;
}

/* MOM_U_IMPLICIT_R / MOM_V_IMPLICIT_R (pkg/mom_common/mom_{u,v}_implicit_r.F:100-300) with implicitViscosity = T,
 * momImplVertAdv = F, selectImplicitDrag = 0: b5d(k) = -deltaTMom recip_hFac(k) recip_drF(k) kappaR(k) recip_drC(k)
 * where mask(k-1) = 1, d5d(k) likewise with kappaR(k+1) recip_drC(k+1) where mask(k+1) = 1, c5d = 1 - (b5d + d5d),
 * on U: i = 1..sNx+1, j = 1..sNy; V: i = 1..sNx, j = 1..sNy+1 (identity elsewhere); SOLVE_TRIDIAGONAL on gU / gV.
 * kappaR is the per-tile (slab, Nr+1) array; gFld is tile3d.  Returns the solver's errCode. */
int og_mom_implicit_r(const og_grid *g, const og_params *p, int bi, int bj, int isV, const double *kappaR, double *gFld) {
  SETUP
  const size_t n = px * py * (size_t)Nr;
  const int iMin = 1, iMax = isV ? sNx : sNx + 1, jMin = 1, jMax = isV ? sNy + 1 : sNy;
  const double *mask = isV ? g->maskS : g->maskW, *rh = isV ? g->recip_hFacS : g->recip_hFacW;
  double *b5d = (double *)calloc(3 * n, sizeof(double));
  double *c5d = b5d + n, *d5d = b5d + 2 * n;
  for (size_t q = 0; q < n; q++) c5d[q] = 1.;
  int err = 0;
  if (Nr > 1) {
    for (int k = 2; k <= Nr; k++)
      for (int j = jMin; j <= jMax; j++)
        for (int i = iMin; i <= iMax; i++)
          if (G3(mask, i, j, k - 1) == 1.)
            L3(b5d, i, j, k) = -p->deltaTMom * G3(rh, i, j, k) * g->recip_drF[k - 1] * L3(kappaR, i, j, k) * g->recip_drC[k - 1];
    for (int k = 1; k <= Nr - 1; k++)
      for (int j = jMin; j <= jMax; j++)
        for (int i = iMin; i <= iMax; i++)
          if (G3(mask, i, j, k + 1) == 1.)
            L3(d5d, i, j, k) = -p->deltaTMom * G3(rh, i, j, k) * g->recip_drF[k - 1] * L3(kappaR, i, j, k + 1) * g->recip_drC[k];
    for (int k = 1; k <= Nr; k++)
      for (int j = jMin; j <= jMax; j++)
        for (int i = iMin; i <= iMax; i++)
          L3(c5d, i, j, k) = 1. - (L3(b5d, i, j, k) + L3(d5d, i, j, k));
    err = solve_tridiagonal(d, b5d, c5d, d5d, gFld + off3);
  }
  free(b5d);
  return err;
}

/* CALC_PHI_HYD, buoyancyRelation 'OCEANIC', integr_GeoPot = 2 (finite volume), uniformFreeSurfLev,
 * linear free surface, followed by CALC_GRAD_PHI_HYD (phi0surf = 0): one tile, one level.
 * rhoInSitu is tile3d; rF has Nr+1, rC Nr entries; phiHydF/phiHydC/dPhiHydX/dPhiHydY are slabs. */
void og_calc_phi_hyd(const og_grid *g, int bi, int bj, int iMin, int iMax, int jMin, int jMax, int k,
                     const double *rhoInSitu, const double *rF, const double *rC, double gravity,
                     double recip_rhoConst, const double *phi0surf,
                     double *phiHydF, double *phiHydC, double *dPhiHydX, double *dPhiHydY) {
  SETUP
  const size_t ns = px * py;
  double *varLoc = (double *)calloc(ns, sizeof(double));
  if (k == 1) FORALL phiHydF[S(i, j)] = 0.;
  double dRlocM = 0.5 * g->drC[k - 1] * 1.;
  if (k == 1) dRlocM = (rF[k - 1] - rC[k - 1]) * 1.;
  double dRlocP;
  if (k == Nr) dRlocP = (rC[k - 1] - rF[k]) * 1.;
  else dRlocP = 0.5 * g->drC[k] * 1.;
  for (int j = jMin; j <= jMax; j++)
    for (int i = iMin; i <= iMax; i++) {
      const double alphaRho = G3(rhoInSitu, i, j, k);
      phiHydC[S(i, j)] = phiHydF[S(i, j)] + dRlocM * gravity * alphaRho * recip_rhoConst;
      phiHydF[S(i, j)] = phiHydC[S(i, j)] + dRlocP * gravity * alphaRho * recip_rhoConst;
    }
  for (int j = jMin; j <= jMax; j++)
    for (int i = iMin; i <= iMax; i++) varLoc[S(i, j)] = phiHydC[S(i, j)] + G2(phi0surf, i, j);
  FORALL { dPhiHydX[S(i, j)] = 0.; dPhiHydY[S(i, j)] = 0.; }
  for (int j = jMin; j <= jMax; j++)
    for (int i = iMin + 1; i <= iMax; i++)
      dPhiHydX[S(i, j)] = G2(g->recip_dxC, i, j) * 1. * (varLoc[S(i, j)] - varLoc[S(i - 1, j)]) * 1.;
  for (int j = jMin + 1; j <= jMax; j++)
    for (int i = iMin; i <= iMax; i++)
      dPhiHydY[S(i, j)] = G2(g->recip_dyC, i, j) * 1. * (varLoc[S(i, j)] - varLoc[S(i, j - 1)]) * 1.;
  free(varLoc);
}

/* INTEGR_CONTINUITY, exactConserv part for one tile at myIter != nIter0, no fresh-water flux
 * (integr_continuity.F): hDivFlow = sum_k maskC (d uTrans + d vTrans); dEtaHdt = -hDivFlow recip_rA;
 * etaN = etaH + implicDiv2DFlow dEtaHdt deltaTFreeSurf on the interior.  (The caller then runs
 * INTEGRATE_FOR_W, exchanges etaN and copies it to etaH = UPDATE_ETAH.) */
void og_integr_continuity_ec(const og_grid *g, const og_params *p, int bi, int bj, const double *uVel,
                             const double *vVel, const double *etaH, double *dEtaHdt, double *etaN,
                             int updateEtaN) {
  SETUP
  const size_t ns = px * py;
  double *hDivFlow = (double *)calloc(ns * 3, sizeof(double));
  double *uTrans = hDivFlow + ns, *vTrans = hDivFlow + 2 * ns;
  for (int k = 1; k <= Nr; k++) {
    for (int j = 1; j <= sNy + 1; j++)
      for (int i = 1; i <= sNx + 1; i++) {
        uTrans[S(i, j)] = G3(uVel, i, j, k) * G2(g->dyG, i, j) * g->drF[k - 1] * G3(g->hFacW, i, j, k);
        vTrans[S(i, j)] = G3(vVel, i, j, k) * G2(g->dxG, i, j) * g->drF[k - 1] * G3(g->hFacS, i, j, k);
      }
    for (int j = 1; j <= sNy; j++)
      for (int i = 1; i <= sNx; i++)
        hDivFlow[S(i, j)] = hDivFlow[S(i, j)]
            + G3(g->maskC, i, j, k) * (uTrans[S(i + 1, j)] - uTrans[S(i, j)] + vTrans[S(i, j + 1)] - vTrans[S(i, j)]);
  }
  for (int j = 1; j <= sNy; j++)
    for (int i = 1; i <= sNx; i++) {
      G2(dEtaHdt, i, j) = -hDivFlow[S(i, j)] * G2(g->recip_rA, i, j) * 1. - 0. * 0.;
      if (updateEtaN)
        G2(etaN, i, j) = G2(etaH, i, j) + p->implicDiv2DFlow * G2(dEtaHdt, i, j) * p->deltaTFreeSurf;
    }
  free(hDivFlow);
}
