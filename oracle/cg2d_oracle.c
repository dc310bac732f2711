/* cg2d_oracle.c -- CPU restatement of CG2D / CG2D_SR / INI_CG2D and the eesupp
 * exchange + global-sum primitives they call.  TEST INFRASTRUCTURE ONLY (see
 * mitgcm_oracle.h).  Follows, loop for loop:
 *   model/src/cg2d.F:100-388, model/src/cg2d_sr.F:111-454,
 *   model/src/ini_cg2d.F:76-234, eesupp/src/exch1_rx.template:170-201,
 *   eesupp/src/exch_s3d_rx.template:8-78, eesupp/src/global_sum_tile.F:185-216.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "mitgcm_oracle.h"

#define PX (d->sNx + 2 * d->OLx)
#define PY (d->sNy + 2 * d->OLy)
/* full-halo 2-D field, Fortran indices */
#define I2(i, j, bi, bj) \
  ((size_t)((i) + d->OLx - 1) + (size_t)PX * ((size_t)((j) + d->OLy - 1) + (size_t)PY * ((size_t)((bi)-1) + (size_t)d->nSx * ((bj)-1))))
/* (0:sNx+1,0:sNy+1,nSx,nSy) work arrays of cg2d.F:91-92 */
#define RX (d->sNx + 2)
#define RY (d->sNy + 2)
#define IR(i, j, bi, bj) \
  ((size_t)(i) + (size_t)RX * ((size_t)(j) + (size_t)RY * ((size_t)((bi)-1) + (size_t)d->nSx * ((bj)-1))))
/* (1:sNx,1:sNy,nSx,nSy) work arrays of cg2d.F:89-90 */
#define IQ(i, j, bi, bj) \
  ((size_t)((i)-1) + (size_t)d->sNx * ((size_t)((j)-1) + (size_t)d->sNy * ((size_t)((bi)-1) + (size_t)d->nSx * ((bj)-1))))
#define IT(bi, bj) (((bi)-1) + d->nSx * ((bj)-1))

static int wrap(int b, int n) { return b < 1 ? b + n : (b > n ? b - n : b); }

/* ---- pkg/exch2 hook -------------------------------------------------------------------------
 * When a tile graph is installed (og_set_exch2_maps), EXCH_XY(Z)_RL and EXCH_S3D_RL follow it
 * instead of the periodic nSx x nSy tiling.  The maps are (dst, src) flat-index lists over a
 * (tiles, PY, PX) array (full width) and a (tiles, sNy+2, sNx+2) array (S3D); oracle/exch2_oracle.py
 * derives them by running its literal restatement of EXCH2_RX1_CUBE on an index-valued field. */
static int e2_nFull = 0, e2_nS3d = 0;
static const long long *e2_dFull = 0, *e2_sFull = 0, *e2_dS3d = 0, *e2_sS3d = 0;
void og_set_exch2_maps(int nFull, const long long *dFull, const long long *sFull,
                       int nS3d, const long long *dS3d, const long long *sS3d) {
  e2_nFull = nFull; e2_dFull = dFull; e2_sFull = sFull;
  e2_nS3d = nS3d; e2_dS3d = dS3d; e2_sS3d = sS3d;
}

/* EXCH1_RX forward mode with EXCH_UPDATE_CORNERS on a single periodic process:
 * X edges are put and received first, then Y edges over the full X range
 * (including the freshly filled X halos), exch1_rx.template:170-201,
 * exch_rx_send_put_y.template:147-151. */
void og_exch_xyz(const og_dims *d, double *a, int nz) {
  const int sNx = d->sNx, sNy = d->sNy, OLx = d->OLx, OLy = d->OLy;
  const size_t px = PX, py = PY;
  if (e2_nFull > 0) {   /* tile graph: all sources are interior cells, so the gather is order-free */
    const size_t slab = px * py;
    for (int k = 0; k < nz; k++)
      for (int q = 0; q < e2_nFull; q++) {
        const size_t dt = (size_t)e2_dFull[q] / slab, dc = (size_t)e2_dFull[q] % slab;
        const size_t st = (size_t)e2_sFull[q] / slab, sc = (size_t)e2_sFull[q] % slab;
        a[dc + slab * ((size_t)k + (size_t)nz * dt)] = a[sc + slab * ((size_t)k + (size_t)nz * st)];
      }
    return;
  }
#define A(i, j, k, bi, bj) \
  a[(size_t)((i) + OLx - 1) + px * ((size_t)((j) + OLy - 1) + py * ((size_t)(k) + (size_t)nz * ((size_t)((bi)-1) + (size_t)d->nSx * ((bj)-1))))]
  for (int bj = 1; bj <= d->nSy; bj++)
    for (int bi = 1; bi <= d->nSx; bi++) {
      int bw = wrap(bi - 1, d->nSx), be = wrap(bi + 1, d->nSx);
      for (int k = 0; k < nz; k++)
        for (int j = 1; j <= sNy; j++)
          for (int o = 1; o <= OLx; o++) {
            A(1 - o, j, k, bi, bj) = A(sNx + 1 - o, j, k, bw, bj);
            A(sNx + o, j, k, bi, bj) = A(o, j, k, be, bj);
          }
    }
  for (int bj = 1; bj <= d->nSy; bj++)
    for (int bi = 1; bi <= d->nSx; bi++) {
      int bs = wrap(bj - 1, d->nSy), bn = wrap(bj + 1, d->nSy);
      for (int k = 0; k < nz; k++)
        for (int o = 1; o <= OLy; o++)
          for (int i = 1 - OLx; i <= sNx + OLx; i++) {
            A(i, 1 - o, k, bi, bj) = A(i, sNy + 1 - o, k, bi, bs);
            A(i, sNy + o, k, bi, bj) = A(i, o, k, bi, bn);
          }
    }
#undef A
}

void og_exch_uv_xyz(const og_dims *d, double *u, double *v, int nz) {
  og_exch_xyz(d, u, nz);
  og_exch_xyz(d, v, nz);
}

/* EXCH_S3D_RL(phi,1): EXCH1_RX with overlap 1 and EXCH_IGNORE_CORNERS. */
void og_exch_s3d(const og_dims *d, double *a) {
  const int sNx = d->sNx, sNy = d->sNy;
  if (e2_nS3d > 0) {
    for (int q = 0; q < e2_nS3d; q++) a[e2_dS3d[q]] = a[e2_sS3d[q]];
    return;
  }
  for (int bj = 1; bj <= d->nSy; bj++)
    for (int bi = 1; bi <= d->nSx; bi++) {
      int bw = wrap(bi - 1, d->nSx), be = wrap(bi + 1, d->nSx);
      int bs = wrap(bj - 1, d->nSy), bn = wrap(bj + 1, d->nSy);
      for (int j = 1; j <= sNy; j++) {
        a[IR(0, j, bi, bj)] = a[IR(sNx, j, bw, bj)];
        a[IR(sNx + 1, j, bi, bj)] = a[IR(1, j, be, bj)];
      }
      for (int i = 1; i <= sNx; i++) {
        a[IR(i, 0, bi, bj)] = a[IR(i, sNy, bi, bs)];
        a[IR(i, sNy + 1, bi, bj)] = a[IR(i, 1, bi, bn)];
      }
    }
}

double og_global_sum_tile(const og_dims *d, const double *tile) {
  double s = 0.;
  for (int bj = 1; bj <= d->nSy; bj++)
    for (int bi = 1; bi <= d->nSx; bi++) s = s + tile[IT(bi, bj)];
  return s;
}

/* INI_CG2D, ini_cg2d.F:76-234 (kSurfC = 1, deepFac2F = 1, no OBCS). */
void og_ini_cg2d(const og_grid *g, const og_params *p, og_cg2d_op *op) {
  const og_dims *d = &g->d;
  const int sNx = d->sNx, sNy = d->sNy, Nr = d->Nr;
  const size_t n2 = (size_t)PX * PY * d->nSx * d->nSy;
  memset(op->aW2d, 0, n2 * sizeof(double));
  memset(op->aS2d, 0, n2 * sizeof(double));
  memset(op->aC2d, 0, n2 * sizeof(double));
  memset(op->pW, 0, n2 * sizeof(double));
  memset(op->pS, 0, n2 * sizeof(double));
  memset(op->pC, 0, n2 * sizeof(double));
#define I3(i, j, k, bi, bj) \
  ((size_t)((i) + d->OLx - 1) + (size_t)PX * ((size_t)((j) + d->OLy - 1) + (size_t)PY * ((size_t)((k)-1) + (size_t)Nr * ((size_t)((bi)-1) + (size_t)d->nSx * ((bj)-1)))))
  double myNorm = 0.;
  for (int bj = 1; bj <= d->nSy; bj++)
    for (int bi = 1; bi <= d->nSx; bi++) {
      for (int k = 1; k <= Nr; k++)
        for (int j = 1; j <= sNy; j++)
          for (int i = 1; i <= sNx; i++) {
            double faceArea = g->dyG[I2(i, j, bi, bj)] * g->drF[k - 1] * g->hFacW[I3(i, j, k, bi, bj)];
            op->aW2d[I2(i, j, bi, bj)] = op->aW2d[I2(i, j, bi, bj)] +
                p->implicSurfPress * p->implicDiv2DFlow * faceArea * g->recip_dxC[I2(i, j, bi, bj)];
            faceArea = g->dxG[I2(i, j, bi, bj)] * g->drF[k - 1] * g->hFacS[I3(i, j, k, bi, bj)];
            op->aS2d[I2(i, j, bi, bj)] = op->aS2d[I2(i, j, bi, bj)] +
                p->implicSurfPress * p->implicDiv2DFlow * faceArea * g->recip_dyC[I2(i, j, bi, bj)];
          }
      for (int j = 1; j <= sNy; j++)
        for (int i = 1; i <= sNx; i++) {
          myNorm = fmax(fabs(op->aW2d[I2(i, j, bi, bj)]), myNorm);
          myNorm = fmax(fabs(op->aS2d[I2(i, j, bi, bj)]), myNorm);
        }
    }
  if (myNorm != 0.) myNorm = 1. / myNorm; else myNorm = 1.;
  for (int bj = 1; bj <= d->nSy; bj++)
    for (int bi = 1; bi <= d->nSx; bi++)
      for (int j = 1; j <= sNy; j++)
        for (int i = 1; i <= sNx; i++) {
          op->aW2d[I2(i, j, bi, bj)] = op->aW2d[I2(i, j, bi, bj)] * myNorm;
          op->aS2d[I2(i, j, bi, bj)] = op->aS2d[I2(i, j, bi, bj)] * myNorm;
        }
  og_exch_uv_xyz(d, op->aW2d, op->aS2d, 1);
  op->cg2dNorm = myNorm;
  op->cg2dNormaliseRHS = (p->cg2dTargetResWunit <= 0.);
  double cg2dTolerance;
  if (op->cg2dNormaliseRHS) cg2dTolerance = p->cg2dTargetResidual;
  else cg2dTolerance = op->cg2dNorm * p->cg2dTargetResWunit * p->globalArea / p->deltaTMom;
  op->cg2dTolerance_sq = cg2dTolerance * cg2dTolerance;

  for (int bj = 1; bj <= d->nSy; bj++)
    for (int bi = 1; bi <= d->nSx; bi++) {
      for (int j = 0; j <= sNy; j++)
        for (int i = 0; i <= sNx; i++) {
          op->aC2d[I2(i, j, bi, bj)] = -(
              op->aW2d[I2(i, j, bi, bj)] + op->aW2d[I2(i + 1, j, bi, bj)]
            + op->aS2d[I2(i, j, bi, bj)] + op->aS2d[I2(i, j + 1, bi, bj)]
            + p->freeSurfFac * myNorm * g->recip_Bo[I2(i, j, bi, bj)]
                * g->rA[I2(i, j, bi, bj)] / p->deltaTMom / p->deltaTFreeSurf);
        }
      for (int j = 1; j <= sNy; j++)
        for (int i = 1; i <= sNx; i++) {
          double aC = op->aC2d[I2(i, j, bi, bj)];
          double aCs = op->aC2d[I2(i, j - 1, bi, bj)];
          double aCw = op->aC2d[I2(i - 1, j, bi, bj)];
          if (aC == 0.) op->pC[I2(i, j, bi, bj)] = 1.;
          else op->pC[I2(i, j, bi, bj)] = 1. / aC;
          if (aC + aCw == 0.) op->pW[I2(i, j, bi, bj)] = 0.;
          else {
            double t = p->cg2dpcOffDFac * (aCw + aC);
            op->pW[I2(i, j, bi, bj)] = -op->aW2d[I2(i, j, bi, bj)] / (t * t);
          }
          if (aC + aCs == 0.) op->pS[I2(i, j, bi, bj)] = 0.;
          else {
            double t = p->cg2dpcOffDFac * (aCs + aC);
            op->pS[I2(i, j, bi, bj)] = -op->aS2d[I2(i, j, bi, bj)] / (t * t);
          }
        }
    }
  og_exch_xyz(d, op->pC, 1);
  og_exch_uv_xyz(d, op->pW, op->pS, 1);
#undef I3
}

/* UPDATE_CG2D, update_cg2d.F:57-192 (non-linear free surface / r*: the column thickness changes every step).
 * hFacW, hFacS of the grid are the CURRENT ones; cg2dNorm and the tolerance stay as INI_CG2D left them
 * (update_cg2d.F:41-42).  No deepAtmosphere, no OBCS, selectImplicitDrag < 2.  updatePreCond as decided by the
 * caller (update_cg2d.F:54-60: cg2dPreCondFreq). */
void og_update_cg2d(const og_grid *g, const og_params *p, og_cg2d_op *op, int updatePreCond) {
  const og_dims *d = &g->d;
  const int sNx = d->sNx, sNy = d->sNy, Nr = d->Nr;
#define I3(i, j, k, bi, bj) \
  ((size_t)((i) + d->OLx - 1) + (size_t)PX * ((size_t)((j) + d->OLy - 1) + (size_t)PY * ((size_t)((k)-1) + (size_t)Nr * ((size_t)((bi)-1) + (size_t)d->nSx * ((bj)-1)))))
  for (int bj = 1; bj <= d->nSy; bj++)
    for (int bi = 1; bi <= d->nSx; bi++) {
      for (int j = 1 - d->OLy; j <= sNy + d->OLy; j++)
        for (int i = 1 - d->OLx; i <= sNx + d->OLx; i++) {
          op->aW2d[I2(i, j, bi, bj)] = 0.;
          op->aS2d[I2(i, j, bi, bj)] = 0.;
        }
      for (int k = 1; k <= Nr; k++)
        for (int j = 1; j <= sNy + 1; j++)
          for (int i = 1; i <= sNx + 1; i++) {
            double faceArea = g->dyG[I2(i, j, bi, bj)] * g->drF[k - 1] * g->hFacW[I3(i, j, k, bi, bj)];
            op->aW2d[I2(i, j, bi, bj)] = op->aW2d[I2(i, j, bi, bj)] + faceArea * g->recip_dxC[I2(i, j, bi, bj)];
            faceArea = g->dxG[I2(i, j, bi, bj)] * g->drF[k - 1] * g->hFacS[I3(i, j, k, bi, bj)];
            op->aS2d[I2(i, j, bi, bj)] = op->aS2d[I2(i, j, bi, bj)] + faceArea * g->recip_dyC[I2(i, j, bi, bj)];
          }
      for (int j = 1; j <= sNy + 1; j++)
        for (int i = 1; i <= sNx + 1; i++) {
          op->aW2d[I2(i, j, bi, bj)] = op->aW2d[I2(i, j, bi, bj)] * op->cg2dNorm * p->implicSurfPress * p->implicDiv2DFlow;
          op->aS2d[I2(i, j, bi, bj)] = op->aS2d[I2(i, j, bi, bj)] * op->cg2dNorm * p->implicSurfPress * p->implicDiv2DFlow;
        }
      for (int j = 1; j <= sNy; j++)
        for (int i = 1; i <= sNx; i++)
          op->aC2d[I2(i, j, bi, bj)] = -(
              op->aW2d[I2(i, j, bi, bj)] + op->aW2d[I2(i + 1, j, bi, bj)]
            + op->aS2d[I2(i, j, bi, bj)] + op->aS2d[I2(i, j + 1, bi, bj)]
            + p->freeSurfFac * op->cg2dNorm * g->recip_Bo[I2(i, j, bi, bj)]
                * g->rA[I2(i, j, bi, bj)] / p->deltaTMom / p->deltaTFreeSurf);
    }
  if (!updatePreCond) return;
  og_exch_xyz(d, op->aC2d, 1);
  for (int bj = 1; bj <= d->nSy; bj++)
    for (int bi = 1; bi <= d->nSx; bi++)
      for (int j = 1; j <= sNy + 1; j++)
        for (int i = 1; i <= sNx + 1; i++) {
          const double aC = op->aC2d[I2(i, j, bi, bj)];
          if (aC == 0.) op->pC[I2(i, j, bi, bj)] = 1.;
          else op->pC[I2(i, j, bi, bj)] = 1. / aC;
          const double pW_tmp = aC + op->aC2d[I2(i - 1, j, bi, bj)];
          if (pW_tmp == 0.) op->pW[I2(i, j, bi, bj)] = 0.;
          else {
            const double t = p->cg2dpcOffDFac * pW_tmp;
            op->pW[I2(i, j, bi, bj)] = -op->aW2d[I2(i, j, bi, bj)] / (t * t);
          }
          const double pS_tmp = aC + op->aC2d[I2(i, j - 1, bi, bj)];
          if (pS_tmp == 0.) op->pS[I2(i, j, bi, bj)] = 0.;
          else {
            const double t = p->cg2dpcOffDFac * pS_tmp;
            op->pS[I2(i, j, bi, bj)] = -op->aS2d[I2(i, j, bi, bj)] / (t * t);
          }
        }
#undef I3
}

#define TILES for (int bj = 1; bj <= d->nSy; bj++) for (int bi = 1; bi <= d->nSx; bi++)
/* Tiles are independent inside a sweep (per-tile partial sums), exactly the reference's
 * thread / MPI-rank parallelism over tiles; results do not depend on the thread count. */
#define TILES_PAR _Pragma("omp parallel for collapse(2) schedule(static)") TILES
#define INTERIOR for (int j = 1; j <= sNy; j++) for (int i = 1; i <= sNx; i++)

/* shared prologue: cg2d.F:100-200 == cg2d_sr.F:111-216 */
static double prologue(const og_dims *d, const og_cg2d_op *op, double *b, double *x,
                       double *r, double *xmin, int nIterMin, double *rhsNormOut,
                       double *sumRHS, double *rhsMaxOut) {
  const int sNx = d->sNx, sNy = d->sNy;
  const int nT = d->nSx * d->nSy;
  double *errTile = (double *)calloc(nT, sizeof(double));
  double *sumRHStile = (double *)calloc(nT, sizeof(double));
  double rhsMax = 0.;
  TILES INTERIOR {
    b[I2(i, j, bi, bj)] = b[I2(i, j, bi, bj)] * op->cg2dNorm;
    rhsMax = fmax(fabs(b[I2(i, j, bi, bj)]), rhsMax);
  }
  double rhsNorm = 1.;
  if (op->cg2dNormaliseRHS) {
    if (rhsMax != 0.) rhsNorm = 1. / rhsMax;
    TILES INTERIOR {
      b[I2(i, j, bi, bj)] = b[I2(i, j, bi, bj)] * rhsNorm;
      x[I2(i, j, bi, bj)] = x[I2(i, j, bi, bj)] * rhsNorm;
    }
  }
  og_exch_xyz(d, x, 1);
  TILES {
    if (nIterMin >= 0) INTERIOR xmin[IQ(i, j, bi, bj)] = x[I2(i, j, bi, bj)];
    INTERIOR {
      r[IR(i, j, bi, bj)] = b[I2(i, j, bi, bj)] -
          (op->aW2d[I2(i, j, bi, bj)] * x[I2(i - 1, j, bi, bj)]
         + op->aW2d[I2(i + 1, j, bi, bj)] * x[I2(i + 1, j, bi, bj)]
         + op->aS2d[I2(i, j, bi, bj)] * x[I2(i, j - 1, bi, bj)]
         + op->aS2d[I2(i, j + 1, bi, bj)] * x[I2(i, j + 1, bi, bj)]
         + op->aC2d[I2(i, j, bi, bj)] * x[I2(i, j, bi, bj)]);
      errTile[IT(bi, bj)] = errTile[IT(bi, bj)] + r[IR(i, j, bi, bj)] * r[IR(i, j, bi, bj)];
      sumRHStile[IT(bi, bj)] = sumRHStile[IT(bi, bj)] + b[I2(i, j, bi, bj)];
    }
  }
  og_exch_s3d(d, r);
  double err_sq = og_global_sum_tile(d, errTile);
  *sumRHS = og_global_sum_tile(d, sumRHStile);
  *rhsNormOut = rhsNorm;
  *rhsMaxOut = rhsMax;
  free(errTile);
  free(sumRHStile);
  return err_sq;
}

static void epilogue(const og_dims *d, const og_cg2d_op *op, double *x, const double *xmin,
                     int nIterMin, double err_sq, double minResidualSq, double rhsNorm) {
  const int sNx = d->sNx, sNy = d->sNy;
  if (nIterMin >= 0 && err_sq > minResidualSq) TILES INTERIOR x[I2(i, j, bi, bj)] = xmin[IQ(i, j, bi, bj)];
  if (op->cg2dNormaliseRHS) TILES INTERIOR x[I2(i, j, bi, bj)] = x[I2(i, j, bi, bj)] / rhsNorm;
}

void og_cg2d(const og_dims *d, const og_cg2d_op *op, double *cg2d_b, double *cg2d_x,
             double *firstResidual, double *minResidualSq, double *lastResidual,
             int *numIters, int *nIterMin, double *sumRHS, double *rhsMax, double *resHist) {
  const int sNx = d->sNx, sNy = d->sNy;
  const int nT = d->nSx * d->nSy;
  const size_t nR = (size_t)RX * RY * nT, nQ = (size_t)sNx * sNy * nT;
  double *r = (double *)calloc(nR, sizeof(double));
  double *s = (double *)calloc(nR, sizeof(double));
  double *q = (double *)calloc(nQ, sizeof(double));
  double *xmin = (double *)calloc(nQ, sizeof(double));
  double *tileA = (double *)calloc(nT, sizeof(double));
  double rhsNorm;
  *minResidualSq = -1.;
  double eta_qrNM1 = 1.;
  double err_sq = prologue(d, op, cg2d_b, cg2d_x, r, xmin, *nIterMin, &rhsNorm, sumRHS, rhsMax);
  int actualIts = 0;
  *firstResidual = sqrt(err_sq);
  if (*nIterMin >= 0) { *nIterMin = 0; *minResidualSq = err_sq; }
  if (!(err_sq < op->cg2dTolerance_sq)) {
    for (int it2d = 1; it2d <= *numIters; it2d++) {
      /* q = M r ; eta = <q,r>   cg2d.F:217-243 */
      TILES_PAR {
        tileA[IT(bi, bj)] = 0.;
        INTERIOR {
          q[IQ(i, j, bi, bj)] =
              op->pC[I2(i, j, bi, bj)] * r[IR(i, j, bi, bj)]
            + op->pW[I2(i, j, bi, bj)] * r[IR(i - 1, j, bi, bj)]
            + op->pW[I2(i + 1, j, bi, bj)] * r[IR(i + 1, j, bi, bj)]
            + op->pS[I2(i, j, bi, bj)] * r[IR(i, j - 1, bi, bj)]
            + op->pS[I2(i, j + 1, bi, bj)] * r[IR(i, j + 1, bi, bj)];
          tileA[IT(bi, bj)] = tileA[IT(bi, bj)] + q[IQ(i, j, bi, bj)] * r[IR(i, j, bi, bj)];
        }
      }
      double eta_qrN = og_global_sum_tile(d, tileA);
      double cgBeta = eta_qrN / eta_qrNM1;
      eta_qrNM1 = eta_qrN;
      TILES_PAR INTERIOR s[IR(i, j, bi, bj)] = q[IQ(i, j, bi, bj)] + cgBeta * s[IR(i, j, bi, bj)];
      og_exch_s3d(d, s);
      /* q = A s ; alpha = <s,q>   cg2d.F:274-301 */
      TILES_PAR {
        tileA[IT(bi, bj)] = 0.;
        INTERIOR {
          q[IQ(i, j, bi, bj)] =
              op->aW2d[I2(i, j, bi, bj)] * s[IR(i - 1, j, bi, bj)]
            + op->aW2d[I2(i + 1, j, bi, bj)] * s[IR(i + 1, j, bi, bj)]
            + op->aS2d[I2(i, j, bi, bj)] * s[IR(i, j - 1, bi, bj)]
            + op->aS2d[I2(i, j + 1, bi, bj)] * s[IR(i, j + 1, bi, bj)]
            + op->aC2d[I2(i, j, bi, bj)] * s[IR(i, j, bi, bj)];
          tileA[IT(bi, bj)] = tileA[IT(bi, bj)] + s[IR(i, j, bi, bj)] * q[IQ(i, j, bi, bj)];
        }
      }
      double alpha = og_global_sum_tile(d, tileA);
      alpha = eta_qrN / alpha;
      /* x += alpha s ; r -= alpha q ; err = <r,r>   cg2d.F:305-327 */
      TILES_PAR {
        tileA[IT(bi, bj)] = 0.;
        INTERIOR {
          cg2d_x[I2(i, j, bi, bj)] = cg2d_x[I2(i, j, bi, bj)] + alpha * s[IR(i, j, bi, bj)];
          r[IR(i, j, bi, bj)] = r[IR(i, j, bi, bj)] - alpha * q[IQ(i, j, bi, bj)];
          tileA[IT(bi, bj)] = tileA[IT(bi, bj)] + r[IR(i, j, bi, bj)] * r[IR(i, j, bi, bj)];
        }
      }
      actualIts = it2d;
      err_sq = og_global_sum_tile(d, tileA);
      if (resHist) resHist[it2d - 1] = sqrt(err_sq);
      if (err_sq < op->cg2dTolerance_sq) break;
      if (err_sq < *minResidualSq) {
        *minResidualSq = err_sq;
        *nIterMin = it2d;
        TILES_PAR INTERIOR xmin[IQ(i, j, bi, bj)] = cg2d_x[I2(i, j, bi, bj)];
      }
      og_exch_s3d(d, r);
    }
  }
  epilogue(d, op, cg2d_x, xmin, *nIterMin, err_sq, *minResidualSq, rhsNorm);
  *lastResidual = sqrt(err_sq);
  *numIters = actualIts;
  free(r); free(s); free(q); free(xmin); free(tileA);
}

void og_cg2d_sr(const og_dims *d, const og_cg2d_op *op, double *cg2d_b, double *cg2d_x,
                double *firstResidual, double *minResidualSq, double *lastResidual,
                int *numIters, int *nIterMin, double *sumRHS, double *rhsMax, double *resHist) {
  const int sNx = d->sNx, sNy = d->sNy;
  const int nT = d->nSx * d->nSy;
  const size_t nR = (size_t)RX * RY * nT, nQ = (size_t)sNx * sNy * nT;
  double *r = (double *)calloc(nR, sizeof(double));
  double *s = (double *)calloc(nR, sizeof(double));
  double *y = (double *)calloc(nR, sizeof(double));
  double *q = (double *)calloc(nQ, sizeof(double));
  double *v = (double *)calloc(nQ, sizeof(double));
  double *xmin = (double *)calloc(nQ, sizeof(double));
  double *t1 = (double *)calloc(nT, sizeof(double));
  double *t2 = (double *)calloc(nT, sizeof(double));
  double *t3 = (double *)calloc(nT, sizeof(double));
  double rhsNorm;
  *minResidualSq = -1.;
  double eta_qrNM1 = 1.;
  double err_sq = prologue(d, op, cg2d_b, cg2d_x, r, xmin, *nIterMin, &rhsNorm, sumRHS, rhsMax);
  int it2d = 0;
  *firstResidual = sqrt(err_sq);
  if (*nIterMin >= 0) { *nIterMin = 0; *minResidualSq = err_sq; }
  if (!(err_sq < op->cg2dTolerance_sq)) {
    /* start-up iteration, cg2d_sr.F:220-291 */
    TILES_PAR {
      t1[IT(bi, bj)] = 0.;
      INTERIOR {
        y[IR(i, j, bi, bj)] =
            op->pC[I2(i, j, bi, bj)] * r[IR(i, j, bi, bj)]
          + op->pW[I2(i, j, bi, bj)] * r[IR(i - 1, j, bi, bj)]
          + op->pW[I2(i + 1, j, bi, bj)] * r[IR(i + 1, j, bi, bj)]
          + op->pS[I2(i, j, bi, bj)] * r[IR(i, j - 1, bi, bj)]
          + op->pS[I2(i, j + 1, bi, bj)] * r[IR(i, j + 1, bi, bj)];
        s[IR(i, j, bi, bj)] = y[IR(i, j, bi, bj)];
        t1[IT(bi, bj)] = t1[IT(bi, bj)] + y[IR(i, j, bi, bj)] * r[IR(i, j, bi, bj)];
      }
    }
    og_exch_s3d(d, s);
    double eta_qrN = og_global_sum_tile(d, t1);
    eta_qrNM1 = eta_qrN;
    TILES_PAR {
      t1[IT(bi, bj)] = 0.;
      INTERIOR {
        q[IQ(i, j, bi, bj)] =
            op->aW2d[I2(i, j, bi, bj)] * s[IR(i - 1, j, bi, bj)]
          + op->aW2d[I2(i + 1, j, bi, bj)] * s[IR(i + 1, j, bi, bj)]
          + op->aS2d[I2(i, j, bi, bj)] * s[IR(i, j - 1, bi, bj)]
          + op->aS2d[I2(i, j + 1, bi, bj)] * s[IR(i, j + 1, bi, bj)]
          + op->aC2d[I2(i, j, bi, bj)] * s[IR(i, j, bi, bj)];
        t1[IT(bi, bj)] = t1[IT(bi, bj)] + s[IR(i, j, bi, bj)] * q[IQ(i, j, bi, bj)];
      }
    }
    double alpha = og_global_sum_tile(d, t1);
    double sigma = eta_qrN / alpha;
    TILES INTERIOR {
      cg2d_x[I2(i, j, bi, bj)] = cg2d_x[I2(i, j, bi, bj)] + sigma * s[IR(i, j, bi, bj)];
      r[IR(i, j, bi, bj)] = r[IR(i, j, bi, bj)] - sigma * q[IQ(i, j, bi, bj)];
    }
    og_exch_s3d(d, r);
    int converged = 0;
    /* main loop, cg2d_sr.F:294-410; Fortran DO leaves it2d = numIters on exhaustion */
    for (it2d = 1; it2d <= *numIters - 1; it2d++) {
      TILES_PAR INTERIOR {
        y[IR(i, j, bi, bj)] =
            op->pC[I2(i, j, bi, bj)] * r[IR(i, j, bi, bj)]
          + op->pW[I2(i, j, bi, bj)] * r[IR(i - 1, j, bi, bj)]
          + op->pW[I2(i + 1, j, bi, bj)] * r[IR(i + 1, j, bi, bj)]
          + op->pS[I2(i, j, bi, bj)] * r[IR(i, j - 1, bi, bj)]
          + op->pS[I2(i, j + 1, bi, bj)] * r[IR(i, j + 1, bi, bj)];
      }
      og_exch_s3d(d, y);
      TILES_PAR {
        t1[IT(bi, bj)] = 0.; t2[IT(bi, bj)] = 0.; t3[IT(bi, bj)] = 0.;
        INTERIOR {
          v[IQ(i, j, bi, bj)] =
              op->aW2d[I2(i, j, bi, bj)] * y[IR(i - 1, j, bi, bj)]
            + op->aW2d[I2(i + 1, j, bi, bj)] * y[IR(i + 1, j, bi, bj)]
            + op->aS2d[I2(i, j, bi, bj)] * y[IR(i, j - 1, bi, bj)]
            + op->aS2d[I2(i, j + 1, bi, bj)] * y[IR(i, j + 1, bi, bj)]
            + op->aC2d[I2(i, j, bi, bj)] * y[IR(i, j, bi, bj)];
          t1[IT(bi, bj)] = t1[IT(bi, bj)] + y[IR(i, j, bi, bj)] * r[IR(i, j, bi, bj)];
          t2[IT(bi, bj)] = t2[IT(bi, bj)] + y[IR(i, j, bi, bj)] * v[IQ(i, j, bi, bj)];
          t3[IT(bi, bj)] = t3[IT(bi, bj)] + r[IR(i, j, bi, bj)] * r[IR(i, j, bi, bj)];
        }
      }
      /* GLOBAL_SUM_VECTOR_RL(3,...) : same tile order per component */
      eta_qrN = og_global_sum_tile(d, t1);
      double delta = og_global_sum_tile(d, t2);
      err_sq = og_global_sum_tile(d, t3);
      if (resHist) resHist[it2d - 1] = sqrt(err_sq);
      if (err_sq < op->cg2dTolerance_sq) { converged = 1; break; }
      if (err_sq < *minResidualSq) {
        *minResidualSq = err_sq;
        *nIterMin = it2d;
        TILES_PAR INTERIOR xmin[IQ(i, j, bi, bj)] = cg2d_x[I2(i, j, bi, bj)];
      }
      double cgBeta = eta_qrN / eta_qrNM1;
      eta_qrNM1 = eta_qrN;
      alpha = delta - (cgBeta * cgBeta) * alpha;
      sigma = eta_qrN / alpha;
      TILES_PAR INTERIOR {
        s[IR(i, j, bi, bj)] = y[IR(i, j, bi, bj)] + cgBeta * s[IR(i, j, bi, bj)];
        cg2d_x[I2(i, j, bi, bj)] = cg2d_x[I2(i, j, bi, bj)] + sigma * s[IR(i, j, bi, bj)];
        q[IQ(i, j, bi, bj)] = v[IQ(i, j, bi, bj)] + cgBeta * q[IQ(i, j, bi, bj)];
        r[IR(i, j, bi, bj)] = r[IR(i, j, bi, bj)] - sigma * q[IQ(i, j, bi, bj)];
      }
      og_exch_s3d(d, r);
    }
    if (!converged) {
      /* cg2d_sr.F:411-423 (loop fell through; it2d == MAX(numIters,1)) */
      TILES_PAR {
        t1[IT(bi, bj)] = 0.;
        INTERIOR t1[IT(bi, bj)] = t1[IT(bi, bj)] + r[IR(i, j, bi, bj)] * r[IR(i, j, bi, bj)];
      }
      err_sq = og_global_sum_tile(d, t1);
    }
  }
  epilogue(d, op, cg2d_x, xmin, *nIterMin, err_sq, *minResidualSq, rhsNorm);
  *lastResidual = sqrt(err_sq);
  *numIters = it2d;
  free(r); free(s); free(y); free(q); free(v); free(xmin); free(t1); free(t2); free(t3);
}
