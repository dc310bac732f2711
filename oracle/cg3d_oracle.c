/* cg3d_oracle.c -- CPU restatement of INI_CG3D (model/src/ini_cg3d.F:60-330) and CG3D (model/src/cg3d.F:13-545),
 * the 3-D preconditioned conjugate-gradient solver of the non-hydrostatic pressure.
 * TEST INFRASTRUCTURE ONLY (see mitgcm_oracle.h).  PARITY UNPINNED beyond the operator normalisation: the known
 * answer `CG3D normalisation factor` of verification/tutorial_deep_convection (5.0000000000000003E-02) is
 * reproduced (tests/test_oracle_golden.py); the solver lines of that experiment need the whole non-hydrostatic
 * step, which is out of reach.  Scope: select_rStar = 0 (no surface term), selectNHfreeSurf = 0, no OBCS
 * (maskInC = 1), deepFac2F = 1, kSurfC = 1 wherever the column is wet.
 * All arrays are tile3d in the halo'd layout; cg3d_q, cg3d_r, cg3d_s (0:sN+1 in the reference) use the same
 * layout here (the ring 0 / sN+1 is what is read). */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "mitgcm_oracle.h"

#define PX (d->sNx + 2 * d->OLx)
#define PY (d->sNy + 2 * d->OLy)
#define I3(i, j, k, bi, bj) \
  ((size_t)((i) + d->OLx - 1) + (size_t)PX * ((size_t)((j) + d->OLy - 1) + (size_t)PY * ((size_t)((k)-1) + (size_t)Nr * ((size_t)((bi)-1) + (size_t)d->nSx * ((bj)-1)))))
#define I2(i, j, bi, bj) \
  ((size_t)((i) + d->OLx - 1) + (size_t)PX * ((size_t)((j) + d->OLy - 1) + (size_t)PY * ((size_t)((bi)-1) + (size_t)d->nSx * ((bj)-1))))
#define TILES for (int bj = 1; bj <= d->nSy; bj++) for (int bi = 1; bi <= d->nSx; bi++)
#define INTERIOR for (int j = 1; j <= sNy; j++) for (int i = 1; i <= sNx; i++)
#define RING for (int j = 0; j <= sNy + 1; j++) for (int i = 0; i <= sNx + 1; i++)

/* INI_CG3D.  nh_Fac_rVel2w2 = nh_Fac*rVel2wUnit(k)**2 + igwFac*dBdrRef(k)*deltaTMom*dTtracerLev(k) is taken
 * k-independent (= 1/nh_Am2 for the z-coordinate non-hydrostatic ocean). */
void og_ini_cg3d(const og_grid *g, const og_params *p, double vertFac, double cg3dTargetResidual,
                 double cg3dTargetResWunit, og_cg3d_op *op) {
  const og_dims *d = &g->d;
  const int sNx = d->sNx, sNy = d->sNy, Nr = d->Nr;
  const size_t n3 = (size_t)PX * PY * Nr * d->nSx * d->nSy;
  double *f[] = {op->aW3d, op->aS3d, op->aV3d, op->aC3d, op->zMC, op->zML, op->zMU};
  for (int n = 0; n < 7; n++) memset(f[n], 0, n3 * sizeof(double));
  const double implicitNHPress = 1.;
  double myNorm = 0.;
  TILES {
    for (int k = 1; k <= Nr; k++) {
      for (int j = 1; j <= sNy; j++)
        for (int i = 1; i <= sNx + 1; i++) {
          const double faceArea = g->dyG[I2(i, j, bi, bj)] * g->drF[k - 1] * g->hFacW[I3(i, j, k, bi, bj)] * 1. * 1.;
          op->aW3d[I3(i, j, k, bi, bj)] = faceArea * g->recip_dxC[I2(i, j, bi, bj)] * implicitNHPress * p->implicDiv2DFlow;
          myNorm = fmax(fabs(op->aW3d[I3(i, j, k, bi, bj)]), myNorm);
        }
      for (int j = 1; j <= sNy + 1; j++)
        for (int i = 1; i <= sNx; i++) {
          const double faceArea = g->dxG[I2(i, j, bi, bj)] * g->drF[k - 1] * g->hFacS[I3(i, j, k, bi, bj)] * 1. * 1.;
          op->aS3d[I3(i, j, k, bi, bj)] = faceArea * g->recip_dyC[I2(i, j, bi, bj)] * implicitNHPress * p->implicDiv2DFlow;
          myNorm = fmax(fabs(op->aS3d[I3(i, j, k, bi, bj)]), myNorm);
        }
    }
    for (int k = 2; k <= Nr; k++) {
      double tmpFac = vertFac;
      if (tmpFac > 0.) tmpFac = 1. / tmpFac;
      INTERIOR {
        const double faceArea = g->rA[I2(i, j, bi, bj)] * g->maskC[I3(i, j, k, bi, bj)] * g->maskC[I3(i, j, k - 1, bi, bj)] * 1. * 1.;
        op->aV3d[I3(i, j, k, bi, bj)] = faceArea * g->recip_drC[k - 1] * tmpFac * implicitNHPress * p->implicDiv2DFlow;
        myNorm = fmax(fabs(op->aV3d[I3(i, j, k, bi, bj)]), myNorm);
      }
    }
  }
  myNorm = myNorm != 0. ? 1. / myNorm : 1.;
  op->cg3dNorm = myNorm;
  op->cg3dNormaliseRHS = cg3dTargetResWunit <= 0.;
  const double tol = op->cg3dNormaliseRHS ? cg3dTargetResidual : myNorm * cg3dTargetResWunit * p->globalArea / p->deltaTMom;
  op->cg3dTolerance_sq = tol * tol;
  TILES {
    for (int k = 1; k <= Nr; k++) INTERIOR {
      const double aW = op->aW3d[I3(i, j, k, bi, bj)], aE = op->aW3d[I3(i + 1, j, k, bi, bj)];
      const double aN = op->aS3d[I3(i, j + 1, k, bi, bj)], aS = op->aS3d[I3(i, j, k, bi, bj)];
      const double aU = op->aV3d[I3(i, j, k, bi, bj)], aL = k != Nr ? op->aV3d[I3(i, j, k + 1, bi, bj)] : 0.;
      op->aC3d[I3(i, j, k, bi, bj)] = -aW - aE - aN - aS - aU - aL;
    }
    INTERIOR {       /* free-surface term in the surface level (kSurfC = 1 where the column is wet, else Nr+1) */
      if (g->maskC[I3(i, j, 1, bi, bj)] != 0.)
        op->aC3d[I3(i, j, 1, bi, bj)] = op->aC3d[I3(i, j, 1, bi, bj)]
            - p->freeSurfFac * g->recip_Bo[I2(i, j, bi, bj)] * g->rA[I2(i, j, bi, bj)] * 1. / p->deltaTMom / p->deltaTFreeSurf;
    }
    for (int k = 1; k <= Nr; k++) INTERIOR {
      op->aW3d[I3(i, j, k, bi, bj)] *= myNorm; op->aS3d[I3(i, j, k, bi, bj)] *= myNorm;
      op->aV3d[I3(i, j, k, bi, bj)] *= myNorm; op->aC3d[I3(i, j, k, bi, bj)] *= myNorm;
    }
  }
  og_exch_uv_xyz(d, op->aW3d, op->aS3d, Nr);
  og_exch_xyz(d, op->aV3d, Nr);
  og_exch_xyz(d, op->aC3d, Nr);
  TILES {
    for (int k = 1; k <= Nr; k++) INTERIOR {
      const size_t q = I3(i, j, k, bi, bj);
      if (op->aC3d[q] != 0.) {
        op->zMC[q] = op->aC3d[q];
        op->zML[q] = op->aV3d[q];
        op->zMU[q] = k != Nr ? op->aV3d[I3(i, j, k + 1, bi, bj)] : 0.;
      } else { op->zMC[q] = 1.; op->zMU[q] = 0.; op->zML[q] = 0.; }
    }
    INTERIOR {
      const size_t q = I3(i, j, 1, bi, bj);
      op->zMC[q] = 1. / op->zMC[q];
      op->zMU[q] = op->zMU[q] * op->zMC[q];
    }
    for (int k = 2; k <= Nr; k++) INTERIOR {
      const size_t q = I3(i, j, k, bi, bj);
      op->zMC[q] = 1. / (op->zMC[q] - op->zML[q] * op->zMU[I3(i, j, k - 1, bi, bj)]);
      op->zMU[q] = op->zMU[q] * op->zMC[q];
    }
    for (int k = 1; k <= Nr; k++) INTERIOR {
      const size_t q = I3(i, j, k, bi, bj);
      if (op->aC3d[q] == 0.) { op->zMC[q] = 1.; op->zML[q] = 0.; op->zMU[q] = 0.; }
    }
  }
  og_exch_xyz(d, op->zMC, Nr);
  og_exch_xyz(d, op->zML, Nr);
  og_exch_xyz(d, op->zMU, Nr);
}

/* CG3D.  maskC is needed for the RHS scaling (cg3d.F:128-130). */
void og_cg3d(const og_dims *d, const og_cg3d_op *op, const double *maskC, double *cg3d_b, double *cg3d_x,
             double *firstResidual, double *lastResidual, int *numIters, double *sumRHSout, double *rhsMaxOut) {
  const int sNx = d->sNx, sNy = d->sNy, Nr = d->Nr;
  const size_t n3 = (size_t)PX * PY * Nr * d->nSx * d->nSy;
  double *q3 = (double *)calloc(3 * n3, sizeof(double)), *r3 = q3 + n3, *s3 = q3 + 2 * n3;
  double *tile = (double *)calloc((size_t)d->nSx * d->nSy, sizeof(double));
#define IT(bi, bj) (((bi)-1) + d->nSx * ((bj)-1))
  double eta_qrNM1 = 1., rhsMax = 0., rhsNorm = 1.;
  TILES for (int k = 1; k <= Nr; k++) INTERIOR {
    const size_t q = I3(i, j, k, bi, bj);
    cg3d_b[q] = cg3d_b[q] * op->cg3dNorm * maskC[q];
    rhsMax = fmax(fabs(cg3d_b[q]), rhsMax);
  }
  if (op->cg3dNormaliseRHS) {
    if (rhsMax != 0.) rhsNorm = 1. / rhsMax;
    TILES for (int k = 1; k <= Nr; k++) INTERIOR {
      const size_t q = I3(i, j, k, bi, bj);
      cg3d_b[q] = cg3d_b[q] * rhsNorm;
      cg3d_x[q] = cg3d_x[q] * rhsNorm;
    }
  }
  og_exch_xyz(d, cg3d_x, Nr);
  double sumRHS, err_sq;
  double *sumT = (double *)calloc((size_t)d->nSx * d->nSy, sizeof(double));
  TILES {
    tile[IT(bi, bj)] = 0.; sumT[IT(bi, bj)] = 0.;
    for (int k = 1; k <= Nr; k++) {
      const int km1 = k - 1 > 1 ? k - 1 : 1, kp1 = k + 1 < Nr ? k + 1 : Nr;
      const double maskM1 = k == 1 ? 0. : 1., maskP1 = k == Nr ? 0. : 1.;
      INTERIOR {
        const size_t q = I3(i, j, k, bi, bj);
        r3[q] = cg3d_b[q]
            - (0. + op->aW3d[q] * cg3d_x[I3(i - 1, j, k, bi, bj)] + op->aW3d[I3(i + 1, j, k, bi, bj)] * cg3d_x[I3(i + 1, j, k, bi, bj)]
               + op->aS3d[q] * cg3d_x[I3(i, j - 1, k, bi, bj)] + op->aS3d[I3(i, j + 1, k, bi, bj)] * cg3d_x[I3(i, j + 1, k, bi, bj)]
               + op->aV3d[q] * cg3d_x[I3(i, j, km1, bi, bj)] * maskM1
               + op->aV3d[I3(i, j, kp1, bi, bj)] * cg3d_x[I3(i, j, kp1, bi, bj)] * maskP1 + op->aC3d[q] * cg3d_x[q]);
        tile[IT(bi, bj)] += r3[q] * r3[q];
        sumT[IT(bi, bj)] += cg3d_b[q];
      }
    }
  }
  og_exch_xyz(d, r3, Nr);                       /* EXCH_S3D_RL( cg3d_r, Nr ): the ring is all that is read */
  sumRHS = og_global_sum_tile(d, sumT);
  err_sq = og_global_sum_tile(d, tile);
  int actualIts = 0;
  *firstResidual = sqrt(err_sq);
  if (sumRHSout) *sumRHSout = sumRHS;
  if (rhsMaxOut) *rhsMaxOut = rhsMax;
  if (!(err_sq < op->cg3dTolerance_sq)) {
    for (int it3d = 1; it3d <= *numIters; it3d++) {
      TILES {
        double eta = 0.;
        RING q3[I3(i, j, 1, bi, bj)] = op->zMC[I3(i, j, 1, bi, bj)] * r3[I3(i, j, 1, bi, bj)];
        for (int k = 2; k <= Nr; k++) RING {
          const size_t q = I3(i, j, k, bi, bj);
          q3[q] = op->zMC[q] * (r3[q] - op->zML[q] * q3[I3(i, j, k - 1, bi, bj)]);
        }
        INTERIOR eta += q3[I3(i, j, Nr, bi, bj)] * r3[I3(i, j, Nr, bi, bj)];
        for (int k = Nr - 1; k >= 1; k--) {
          RING {
            const size_t q = I3(i, j, k, bi, bj);
            q3[q] = q3[q] - op->zMU[q] * q3[I3(i, j, k + 1, bi, bj)];
          }
          INTERIOR eta += q3[I3(i, j, k, bi, bj)] * r3[I3(i, j, k, bi, bj)];
        }
        tile[IT(bi, bj)] = eta;
      }
      const double eta_qrN = og_global_sum_tile(d, tile);
      const double cgBeta = eta_qrN / eta_qrNM1;
      eta_qrNM1 = eta_qrN;
      TILES for (int k = 1; k <= Nr; k++) RING {
        const size_t q = I3(i, j, k, bi, bj);
        s3[q] = q3[q] + cgBeta * s3[q];
      }
      TILES {
        double al = 0.;
        for (int k = 1; k <= Nr; k++) INTERIOR {
          const size_t q = I3(i, j, k, bi, bj);
          double v = op->aW3d[q] * s3[I3(i - 1, j, k, bi, bj)] + op->aW3d[I3(i + 1, j, k, bi, bj)] * s3[I3(i + 1, j, k, bi, bj)]
                   + op->aS3d[q] * s3[I3(i, j - 1, k, bi, bj)] + op->aS3d[I3(i, j + 1, k, bi, bj)] * s3[I3(i, j + 1, k, bi, bj)];
          if (k > 1) v = v + op->aV3d[q] * s3[I3(i, j, k - 1, bi, bj)];
          if (k < Nr) v = v + op->aV3d[I3(i, j, k + 1, bi, bj)] * s3[I3(i, j, k + 1, bi, bj)];
          v = v + op->aC3d[q] * s3[q];
          q3[q] = v;
          al += s3[q] * v;
        }
        tile[IT(bi, bj)] = al;
      }
      double alpha = og_global_sum_tile(d, tile);
      alpha = eta_qrN / alpha;
      TILES {
        double e = 0.;
        for (int k = 1; k <= Nr; k++) INTERIOR {
          const size_t q = I3(i, j, k, bi, bj);
          cg3d_x[q] = cg3d_x[q] + alpha * s3[q];
          r3[q] = r3[q] - alpha * q3[q];
          e += r3[q] * r3[q];
        }
        tile[IT(bi, bj)] = e;
      }
      actualIts = it3d;
      err_sq = og_global_sum_tile(d, tile);
      if (err_sq < op->cg3dTolerance_sq) break;
      og_exch_xyz(d, r3, Nr);
    }
  }
  if (op->cg3dNormaliseRHS)
    TILES for (int k = 1; k <= Nr; k++) INTERIOR cg3d_x[I3(i, j, k, bi, bj)] = cg3d_x[I3(i, j, k, bi, bj)] / rhsNorm;
  *lastResidual = sqrt(err_sq);
  *numIters = actualIts;
  free(q3); free(tile); free(sumT);
}
