/* mitgcm_oracle.h -- CPU restatement ("oracle") of MITgcm's CG2D / GAD_CALC_RHS /
 * MOM_FLUXFORM hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load it, and only as the checker or the timed CPU baseline.
 *
 * The reference (Fortran 77, /root/reference) cannot be compiled in this image
 * (no Fortran compiler), so this is a loop-for-loop C restatement in the same
 * operation order and the same array layout.  It is PINNED against the
 * reference's own golden output:
 *   verification/tutorial_barotropic_gyre/results/output.txt
 *     (cg2dNorm, per-step cg2d_init_res / iters / last_res, Sum(rhs),rhsMax,
 *      %MON dynstat_{eta,uvel,vvel}_* for 10 steps) -- see tests/test_oracle_golden.py
 * Compile with -O2 -ffp-contract=off (the goldens are non-FMA -ieee builds).
 *
 * Layout (model/inc/SIZE.h): element (i,j,k,bi,bj) of a
 * Fortran array (1-OLx:sNx+OLx, 1-OLy:sNy+OLy, Nr, nSx, nSy) lives at
 *   (i+OLx-1) + PX*((j+OLy-1) + PY*((k-1) + Nr*((bi-1) + nSx*(bj-1)))).
 * All index arguments below are the Fortran (1-based, halo-negative) values.
 */
#ifndef MITGCM_ORACLE_H
#define MITGCM_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
  int sNx, sNy, OLx, OLy, nSx, nSy, Nr;
} og_dims;

/* Grid / mask arrays (model/inc/GRID.h:311-506).  2-D arrays are tile2d
 * (PX*PY*nSx*nSy), 3-D arrays tile3d (PX*PY*Nr*nSx*nSy), cosFac* are
 * (PY*nSx*nSy), vertical arrays have Nr (drC, recip_drC: Nr+1) entries.
 * deepFac*, rhoFac* are identically 1 for every config in scope (no deep
 * atmosphere / anelastic) and are omitted: multiplying by 1.0 is exact. */
typedef struct {
  og_dims d;
  double *dxC, *dyC, *dxG, *dyG, *dxF, *dyF, *dxV, *dyU;
  double *rA, *rAw, *rAs, *rAz;
  double *recip_dxC, *recip_dyC, *recip_dxG, *recip_dyG, *recip_dxF, *recip_dyF,
         *recip_dxV, *recip_dyU;
  double *recip_rA, *recip_rAw, *recip_rAs, *recip_rAz;
  double *fCori, *fCoriG, *tanPhiAtU, *tanPhiAtV;
  double *cosFacU, *cosFacV;
  double *drF, *drC, *recip_drF, *recip_drC;
  double *hFacC, *hFacW, *hFacS, *recip_hFacC, *recip_hFacW, *recip_hFacS;
  double *maskC, *maskW, *maskS;
  double *recip_Bo;            /* tile2d, ini_linear_phisurf.F:84-85 */
} og_grid;

/* Run-time parameters on the path (model/inc/PARAMS.h). */
typedef struct {
  /* time stepping / free surface */
  double deltaTMom, deltaTFreeSurf, freeSurfFac, implicSurfPress, implicDiv2DFlow;
  double rkSign;                       /* -1 for z coordinates */
  /* cg2d */
  double cg2dpcOffDFac, cg2dTargetResidual, cg2dTargetResWunit, globalArea;
  /* momentum */
  double viscAhD, viscAhZ, viscA4D, viscA4Z;
  double sideDragFactor, bottomDragLinear, bottomDragQuadratic;
  double recip_rSphere;
  double afFacMom, vfFacMom, cfFacMom, mtFacMom;
  int momAdvection, momViscosity, useBiharmonicVisc, implicitViscosity;
  int no_slip_sides, no_slip_bottom, bottomVisc_pCell;
  int selectBotDragQuadr, selectImplicitDrag;
  int useCDscheme, selectCoriScheme, selectMetricTerms, usingSphericalPolarGrid;
  int rigidLid, select_rStar;
  int selectKEscheme;
  /* tracers */
  int implicitDiffusion;
  /* vector-invariant momentum (pkg/mom_vecinv) */
  int useCoriolis, useAbsVorticity, selectVortScheme, useJamartMomAdv, upwindShear;
  int highOrderVorticity, upwindVorticity, momImplVertAdv;
} og_params;

/* ---- eesupp primitives (single process, periodic; nPx=nPy=1) ------------- */
/* EXCH_XY_RL / EXCH_XYZ_RL: full-width halo fill with corners, X phase then Y
 * phase (eesupp/src/exch1_rx.template:170-201). nz = levels per tile. */
void og_exch_xyz(const og_dims *d, double *a, int nz);
/* EXCH_UV_XY(Z)_RS/RL on a non-cube topology = two scalar exchanges
 * (eesupp/src/exch_uv_xyz_rx.template). */
void og_exch_uv_xyz(const og_dims *d, double *u, double *v, int nz);
/* EXCH_S3D_RL(phi,1): width-1 halo, no corners, on a (0:sNx+1,0:sNy+1) array
 * (eesupp/src/exch_s3d_rx.template:8-78). */
void og_exch_s3d(const og_dims *d, double *a);
/* pkg/exch2 tile graph for the three exchanges above (NULL / 0 restores the periodic tiling) */
void og_set_exch2_maps(int nFull, const long long *dFull, const long long *sFull,
                       int nS3d, const long long *dS3d, const long long *sS3d);
/* GLOBAL_SUM_TILE_RL: tiles summed bi fast, bj slow (global_sum_tile.F:185-191) */
double og_global_sum_tile(const og_dims *d, const double *tile);

/* ---- CG2D ---------------------------------------------------------------- */
typedef struct {
  double *aW2d, *aS2d, *aC2d, *pW, *pS, *pC;   /* tile2d, model/inc/CG2D.h:32-42 */
  double cg2dNorm, cg2dTolerance_sq;
  int cg2dNormaliseRHS;
} og_cg2d_op;

/* INI_CG2D (model/src/ini_cg2d.F:76-234) */
void og_ini_cg2d(const og_grid *g, const og_params *p, og_cg2d_op *op);
void og_update_cg2d(const og_grid *g, const og_params *p, og_cg2d_op *op, int updatePreCond);

/* CG2D (model/src/cg2d.F:13-415).  resHist (may be NULL) receives sqrt(err_sq)
 * after each iteration; sumRHS/rhsMax are the values cg2d.F:199-200 prints. */
void og_cg2d(const og_dims *d, const og_cg2d_op *op, double *cg2d_b, double *cg2d_x,
             double *firstResidual, double *minResidualSq, double *lastResidual,
             int *numIters, int *nIterMin, double *sumRHS, double *rhsMax,
             double *resHist);
/* CG2D_SR (model/src/cg2d_sr.F:13-458) */
void og_cg2d_sr(const og_dims *d, const og_cg2d_op *op, double *cg2d_b, double *cg2d_x,
                double *firstResidual, double *minResidualSq, double *lastResidual,
                int *numIters, int *nIterMin, double *sumRHS, double *rhsMax,
                double *resHist);

/* ---- CG3D (model/src/cg3d.F:13-545, model/src/ini_cg3d.F, model/inc/CG3D.h) -- see cg3d_oracle.c ---------- */
typedef struct {
  double *aW3d, *aS3d, *aV3d, *aC3d, *zMC, *zML, *zMU;   /* tile3d */
  double cg3dNorm, cg3dTolerance_sq;
  int cg3dNormaliseRHS;
} og_cg3d_op;
void og_ini_cg3d(const og_grid *g, const og_params *p, double vertFac, double cg3dTargetResidual,
                 double cg3dTargetResWunit, og_cg3d_op *op);
void og_cg3d(const og_dims *d, const og_cg3d_op *op, const double *maskC, double *cg3d_b, double *cg3d_x,
             double *firstResidual, double *lastResidual, int *numIters, double *sumRHS, double *rhsMax);

/* ---- MOM_FLUXFORM (pkg/mom_fluxform/mom_fluxform.F:42-1064) ------------- */
/* One tile, one level.  uVel,vVel,wVel,gU,gV are tile3d (COMMON DYNVARS.h);
 * kappaRU/V are (PX*PY*(Nr+1)) per-tile arrays; the six slabs are PX*PY. */
void og_mom_fluxform(const og_grid *g, const og_params *p, int bi, int bj, int k,
                     int iMin, int iMax, int jMin, int jMax,
                     const double *kappaRU, const double *kappaRV,
                     double *fVerUkm, double *fVerVkm, double *fVerUkp, double *fVerVkp,
                     double *guDiss, double *gvDiss,
                     const double *uVel, const double *vVel, const double *wVel,
                     double *gU, double *gV);

/* ---- MOM_VECINV (pkg/mom_vecinv/mom_vecinv.F:10-1009) ------------------- */
/* Same conventions as og_mom_fluxform.  csCorners: bit mask of the facet corners this tile owns on
 * the cubed sphere (1 SW, 2 SE, 4 NE, 8 NW; 0 = not a cube), myFace the facet number.  Returns
 * non-zero for options that are not restated. */
int og_mom_vecinv(const og_grid *g, const og_params *p, int bi, int bj, int k,
                  int iMin, int iMax, int jMin, int jMax,
                  const double *kappaRU, const double *kappaRV,
                  const double *fVerUkm, const double *fVerVkm, double *fVerUkp, double *fVerVkp,
                  double *guDiss, double *gvDiss,
                  const double *uVel, const double *vVel, const double *wVel,
                  double *gU, double *gV, int csCorners, int myFace);

/* ---- GAD_CALC_RHS (pkg/generic_advdiff/gad_calc_rhs.F:10-795) ------------ */
/* One tile, one level, one tracer.  Slab args are PX*PY; TracerN, TracAB,
 * gTracer are (PX*PY*Nr) per-tile; fVerT is (PX*PY*2). */
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
                     double *fZon, double *fMer, double *fVerT, double *gTracer);

/* GAD_ADVECTION (pkg/generic_advdiff/gad_advection.F): multi-dimensional advection, one tile, all levels.
 * Cubed sphere (3 passes): nCFace = exch2_myFace of the tile, edges = 1 N | 2 S | 4 E | 8 W facet edges it touches;
 * 0, 0 otherwise (2 passes). */
int og_gad_advection(const og_grid *g, const og_params *p, int bi, int bj, int advectionScheme,
                     int vertAdvecScheme, int implicitAdvection, int compressible,
                     const double *deltaTLev, const double *uFld, const double *vFld, const double *wFld,
                     const double *tracer, double *gTracer, int nCFace, int edges);

/* CALC_ADV_FLOW (model/src/calc_adv_flow.F) for one tile, one level */
void og_calc_adv_flow(const og_grid *g, int bi, int bj, int k,
                      const double *uVel, const double *vVel, const double *wVel,
                      double *xA, double *yA, double *maskUp,
                      double *uFld, double *vFld, double *wFld,
                      double *uTrans, double *vTrans, double *rTrans, double *rTransKp1);

#ifdef __cplusplus
}
#endif
#endif
