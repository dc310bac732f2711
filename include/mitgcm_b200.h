/* mitgcm_b200.h -- C ABI of the B200-native CG2D / GAD_CALC_RHS / MOM_FLUXFORM path.
 *
 * Calling convention = what a Fortran-77 MITgcm build binds without any glue
 * (the convention the tree already uses for cloc_/sigreg_, eesupp/src/tim.c,
 * tools/genmake2:792-886): lower-case name with a trailing underscore, every
 * argument by reference, LOGICAL passed as 4-byte int, arrays in the eesupp tile
 * layout (1-OLx:sNx+OLx, 1-OLy:sNy+OLy, [Nr,] nSx, nSy) of model/inc/SIZE.h.
 * No torch types, no C++ types.  Array arguments may be HOST pointers (the
 * reference call sites; the library stages them through the GPU) or DEVICE
 * pointers (resident state; no copies) -- detected per pointer.
 *
 * Error convention: the library never exits.  Every entry point
 * that can fail sets *ierr (0 = ok); hot-path entry points that keep the exact
 * reference argument list record the code for mitgcm_b200_last_error_().
 * Solver non-convergence is NOT an error (cg2d.F:358-369).
 * Threading: master thread only (nTx = nTy = 1); one process <-> one GPU.
 */
#ifndef MITGCM_B200_H
#define MITGCM_B200_H

#ifdef __cplusplus
extern "C" {
#endif

/* ---- field ids for the device mirrors of COMMON-block arrays --------------- */
/* 2-D tile arrays, model/inc/GRID.h:311-506, CG2D.h:32-42, SURFACE.h, FFIELDS.h */
enum {
  MG_DXC = 0, MG_DYC, MG_DXG, MG_DYG, MG_DXF, MG_DYF, MG_DXV, MG_DYU,
  MG_RA, MG_RAW, MG_RAS, MG_RAZ,
  MG_RECIP_DXC, MG_RECIP_DYC, MG_RECIP_DXG, MG_RECIP_DYG, MG_RECIP_DXF, MG_RECIP_DYF,
  MG_RECIP_DXV, MG_RECIP_DYU, MG_RECIP_RA, MG_RECIP_RAW, MG_RECIP_RAS, MG_RECIP_RAZ,
  MG_FCORI, MG_FCORIG, MG_TANPHIATU, MG_TANPHIATV, MG_RECIP_BO, MG_BO_SURF,
  MG_AW2D, MG_AS2D, MG_AC2D, MG_PW, MG_PS, MG_PC,
  MG_ETAN, MG_SURFFORCU, MG_SURFFORCV, MG_SURFFORCT, MG_CG2D_B, MG_CG2D_X,
  MG_ETAH, MG_DETAHDT, MG_SST, MG_LAMBDATHETACLIMRELAX,   /* SURFACE.h, FFIELDS.h */
  MG_N2D,
  /* 3-D tile arrays (Nr levels), GRID.h, DYNVARS.h */
  MG_HFACC = 100, MG_HFACW, MG_HFACS, MG_RECIP_HFACC, MG_RECIP_HFACW, MG_RECIP_HFACS,
  MG_MASKC, MG_MASKW, MG_MASKS,
  MG_UVEL, MG_VVEL, MG_WVEL, MG_THETA, MG_SALT, MG_GU, MG_GV, MG_GUNM1, MG_GVNM1,
  MG_GT, MG_GTNM1, MG_GS, MG_GSNM1, MG_PHIHYD, MG_KAPPART, MG_THETA2, MG_RHOINSITU,
  /* CG3D operators and preconditioner, COMMON /CG3D_R/ (model/inc/CG3D.h:30-48) */
  MG_AW3D, MG_AS3D, MG_AV3D, MG_AC3D, MG_ZMC, MG_ZML, MG_ZMU,
  MG_SALT2, MG_KAPPARS,
  MG_N3D_END,
  /* (Nr+1)-level tile arrays */
  MG_KAPPARU = 200, MG_KAPPARV, MG_N3DP_END,
  /* per-row arrays (1-OLy:sNy+OLy,nSx,nSy) */
  MG_COSFACU = 300, MG_COSFACV, MG_NJ_END,
  /* vertical arrays: drF,recip_drF (Nr); drC,recip_drC (Nr+1) */
  MG_DRF = 400, MG_DRC, MG_RECIP_DRF, MG_RECIP_DRC, MG_TREF, MG_SREF, MG_RF, MG_RC, MG_NK_END
};

/* ---- run-time parameter ids (model/inc/PARAMS.h) ---------------------------- */
enum {
  MP_DELTATMOM = 0, MP_DELTATFREESURF, MP_FREESURFFAC, MP_IMPLICSURFPRESS, MP_IMPLICDIV2DFLOW,
  MP_RKSIGN, MP_CG2DNORM, MP_CG2DTOLERANCE_SQ,
  MP_VISCAHD, MP_VISCAHZ, MP_VISCA4D, MP_VISCA4Z, MP_SIDEDRAGFACTOR, MP_BOTTOMDRAGLINEAR,
  MP_BOTTOMDRAGQUADRATIC, MP_RECIP_RSPHERE, MP_AFFACMOM, MP_VFFACMOM, MP_CFFACMOM, MP_MTFACMOM,
  MP_ABEPS, MP_DELTATTRACER, MP_DIFFKHT, MP_DIFFK4T, MP_GRAVITY, MP_TALPHA, MP_RHONIL, MP_RHOCONST,
  MP_DIFFKRT, MP_VISCAR, MP_SBETA, MP_IVDC_KAPPA, MP_CG3DNORM, MP_CG3DTOLERANCE_SQ,
  MP_DIFFKHS, MP_DIFFK4S, MP_DIFFKRS, MP_CG2DPCOFFDFAC,
  MP_ND,
  MI_CG2DNORMALISERHS = 100, MI_CG2DMAXITERS, MI_CG2DUSEMINRESSOL, MI_PRINTRESIDUALFREQ,
  MI_MOMADVECTION, MI_MOMVISCOSITY, MI_USEBIHARMONICVISC, MI_IMPLICITVISCOSITY,
  MI_NO_SLIP_SIDES, MI_NO_SLIP_BOTTOM, MI_BOTTOMVISC_PCELL, MI_SELECTBOTDRAGQUADR,
  MI_SELECTIMPLICITDRAG, MI_USECDSCHEME, MI_SELECTCORISCHEME, MI_SELECTMETRICTERMS,
  MI_USINGSPHERICALPOLARGRID, MI_RIGIDLID, MI_SELECT_RSTAR, MI_IMPLICITDIFFUSION,
  MI_MOMFORCING, MI_MOMDISSIP_IN_AB, MI_TEMPADVSCHEME, MI_TEMPVERTADVSCHEME, MI_USESRCGSOLVER,
  MI_TEMPSTEPPING, MI_NITER0, MI_PROFILE,
  MI_EXACTCONSERV, MI_BUOYANCYLINEAR, MI_DOTHETACLIMRELAX, MI_GAD_MULTIDIM_COMPRESSIBLE,
  /* pkg/mom_vecinv */
  MI_USECORIOLIS, MI_USEABSVORTICITY, MI_SELECTVORTSCHEME, MI_USEJAMARTMOMADV, MI_UPWINDSHEAR,
  MI_SELECTKESCHEME, MI_HIGHORDERVORTICITY, MI_UPWINDVORTICITY, MI_MOMIMPLVERTADV,
  MI_VECTORINVARIANTMOMENTUM, MI_CG3DNORMALISERHS, MI_MULTIDIMADVECTION,
  MI_SALTSTEPPING, MI_SALTADVSCHEME, MI_SALTVERTADVSCHEME, MI_CG2DPRECONDFREQ,
  MI_NI_END
};

/* ---- life cycle -------------------------------------------------------------- */
/* dims = {sNx,sNy,OLx,OLy,nSx,nSy,Nr,nPx,nPy,myPx,myPy} (model/inc/SIZE.h + the process
 * position INI_PROCS derives, eesupp/src/ini_procs.F:145-260); device = CUDA ordinal or -1
 * for LOCAL_RANK / 0. */
void mitgcm_b200_init_(const int *dims, const int *device, int *ierr);
void mitgcm_b200_finalize_(void);
int  mitgcm_b200_last_error_(void);
const char *mitgcm_b200_last_error_string(void);

void mitgcm_b200_set_param_d_(const int *id, const double *val, int *ierr);
void mitgcm_b200_set_param_i_(const int *id, const int *val, int *ierr);

/* device mirrors: host -> device (creates the mirror on first use), device -> host,
 * and the raw device address (for callers that keep state resident).  The address of a mirror is stable except
 * for MG_THETA / MG_THETA2 and MG_SALT / MG_SALT2: CYCLE_TRACER (model/src/cycle_tracer.F) is a pointer swap,
 * so these four addresses are INVALIDATED by mitgcm_b200_forward_step_ / _step_part_ -- query them again after
 * every step (tests/test_step_gpu.py::test_theta_field_ptr_after_steps). */
void mitgcm_b200_set_field_(const int *id, const double *host, int *ierr);
void mitgcm_b200_get_field_(const int *id, double *host, int *ierr);
double *mitgcm_b200_field_ptr(int id);
void mitgcm_b200_fill_field_(const int *id, const double *value, int *ierr);   /* mirror := value */
/* Page-locks a HOST array the caller will keep passing to the per-call entry points (the COMMON-block arrays cg2d_b,
 * cg2d_x, gU, gV, theta ... have fixed addresses for the whole run), so that their staging copies are direct DMA
 * transfers instead of going through the driver's bounce buffer (cudaHostRegister; undone by mitgcm_b200_finalize_).
 * nDoubles = length of the array.  Optional: unpinned arrays work, only slower. */
void mitgcm_b200_pin_host_(double *host, const long long *nDoubles, int *ierr);
void mitgcm_b200_sync_(void);
/* Timing on the library's own stream (CUDA events), for bench.py: record event `slot`
 * (0..15) / elapsed milliseconds between two recorded slots (synchronises on the later one). */
void mitgcm_b200_event_record_(const int *slot);
void mitgcm_b200_event_elapsed_ms_(const int *slotA, const int *slotB, double *ms);
/* Number of kernels the library has launched since init (bench.py's gpu_launches). */
long long mitgcm_b200_launch_count_(void);
/* Per-phase device time of the last forward step when MI_PROFILE != 0, milliseconds:
 * {thermo, dyn, rhs, cg2d, eta+exch(x), corr, exchanges}. */
void mitgcm_b200_step_timings_(double *ms7);

/* ---- CG2D / CG2D_SR ------------------------------------------------------------
 * Same argument list as SUBROUTINE CG2D (model/src/cg2d.F:13-17) and CG2D_SR
 * (model/src/cg2d_sr.F:13-17); operators come from the mirrors MG_AW2D..MG_PC of
 * COMMON /CG2D_I_RS/ (model/inc/CG2D.h:32-42), cg2dNorm / cg2dTolerance_sq /
 * cg2dNormaliseRHS from the parameters.  Caller: solve_for_pressure.F:292,309. */
void cg2d_b200_(double *cg2d_b, double *cg2d_x, double *firstResidual, double *minResidualSq,
                double *lastResidual, int *numIters, int *nIterMin, const int *myThid);
void cg2d_sr_b200_(double *cg2d_b, double *cg2d_x, double *firstResidual, double *minResidualSq,
                   double *lastResidual, int *numIters, int *nIterMin, const int *myThid);
/* UPDATE_CG2D, same argument list as model/src/update_cg2d.F:7 (caller: forward_step.F with the non-linear free
 * surface, once per step): rebuilds the mirrors MG_AW2D, MG_AS2D, MG_AC2D and -- when myIter = nIter0 or
 * MOD(myIter, cg2dPreCondFreq) = 0 (MI_CG2DPRECONDFREQ, default 1; MI_NITER0) -- MG_PC, MG_PW, MG_PS from the current
 * mirrors MG_HFACW / MG_HFACS (the caller refreshes those after UPDATE_SURF_DR / r*) and the 2-D grid mirrors;
 * MP_CG2DNORM, MP_CG2DPCOFFDFAC (default 0.51), implicSurfPress, implicDiv2DFlow, freeSurfFac and the time steps
 * come from the parameters.  Errors are recorded for mitgcm_b200_last_error_(). */
void update_cg2d_b200_(const double *myTime, const int *myIter, const int *myThid);
/* The values cg2d.F:199-200 prints inside the solver (`cg2d: Sum(rhs),rhsMax`), and the
 * per-iteration residuals of cg2d.F:329-336 (n <= iterations run). */
void mitgcm_b200_cg2d_stats_(double *sumRHS, double *rhsMax);
void mitgcm_b200_cg2d_residuals_(double *resid, const int *n);

/* ---- GAD_CALC_RHS ---------------------------------------------------------------
 * Same argument list as pkg/generic_advdiff/gad_calc_rhs.F:10-21 (callers
 * temp_integrate.F:357, salt_integrate.F:349, ptracers_integrate.F:315).  Grid arrays
 * come from the mirrors. */
void gad_calc_rhs_b200_(const int *bi, const int *bj, const int *iMin, const int *iMax,
                        const int *jMin, const int *jMax, const int *k, const int *kM1,
                        const int *kUp, const int *kDown,
                        const double *xA, const double *yA, const double *maskUp,
                        const double *uFld, const double *vFld, const double *wFld,
                        const double *uTrans, const double *vTrans, const double *rTrans,
                        const double *rTransKp1, const double *diffKh, const double *diffK4,
                        const double *KappaR, const double *diffKr4,
                        const double *TracerN, const double *TracAB, const double *deltaTLev,
                        const int *trIdentity, const int *advectionSchArg, const int *vertAdvecSchArg,
                        const int *calcAdvection, const int *implicitAdvection,
                        const int *applyAB_onTracer, const int *trUseDiffKr4, const int *trUseGMRedi,
                        const int *trUseKPP, const int *trUseSmolHack,
                        double *fZon, double *fMer, double *fVerT, double *gTracer,
                        const double *myTime, const int *myIter, const int *myThid);

/* ---- GAD_ADVECTION ---------------------------------------------------------------
 * Multi-dimensional advection of one tracer on one tile, all levels: same argument list as
 * pkg/generic_advdiff/gad_advection.F:11-17 (callers temp_integrate.F:283, salt_integrate.F:275,
 * ptracers_integrate.F).  uFld, vFld, wFld, gTracer are (slab, Nr) arrays of the tile, tracer the
 * full (.., Nr, nSx, nSy) array.  The GAD_MULTIDIM_COMPRESSIBLE build option (GAD_OPTIONS.h:44) is
 * the run-time parameter MI_GAD_MULTIDIM_COMPRESSIBLE.  With a pkg/exch2 topology (mitgcm_b200_set_exch2_topology_
 * + mitgcm_b200_set_cs_tiles_) the three facet-dependent passes of the cubed sphere run. */
void gad_advection_b200_(const int *implicitAdvection, const int *advectionSchArg, const int *vertAdvecSchArg,
                         const int *trIdentity, const double *deltaTLev, const double *uFld,
                         const double *vFld, const double *wFld, const double *tracer, double *gTracer,
                         const int *bi, const int *bj, const double *myTime, const int *myIter,
                         const int *myThid);

/* ---- CG3D ---------------------------------------------------
 * Same argument list as SUBROUTINE CG3D (model/src/cg3d.F:13-17; caller solve_for_pressure.F:427).  Operators and
 * preconditioner come from the mirrors MG_AW3D .. MG_ZMU of COMMON /CG3D_R/ (model/inc/CG3D.h:30-48, full-halo
 * tile3d as INI_CG3D leaves them), maskC from MG_MASKC, cg3dNorm / cg3dTolerance_sq / cg3dNormaliseRHS from the
 * parameters.  cg3d_b returns normalised, cg3d_x the interior solution, as in the reference.  Single rank,
 * select_rStar = 0. */
void cg3d_b200_(double *cg3d_b, double *cg3d_x, double *firstResidual, double *lastResidual, int *numIters,
                const int *myIter, const int *myThid);
/* the values cg3d.F:243-244 prints inside the solver (`cg3d: Sum(rhs),rhsMax`) */
void mitgcm_b200_cg3d_rhs_stats_(double *sumRHS, double *rhsMax);

/* ---- MOM_FLUXFORM -----------------------------------------------------------------
 * Same argument list as pkg/mom_fluxform/mom_fluxform.F:42-48 (caller dynamics.F:517),
 * followed by the COMMON /DYNVARS_R/ arrays the routine reads and writes
 * (model/inc/DYNVARS.h): the shim passes them explicitly. */
void mom_fluxform_b200_(const int *bi, const int *bj, const int *k, const int *iMin, const int *iMax,
                        const int *jMin, const int *jMax,
                        const double *kappaRU, const double *kappaRV,
                        double *fVerUkm, double *fVerVkm, double *fVerUkp, double *fVerVkp,
                        double *guDiss, double *gvDiss,
                        const double *myTime, const int *myIter, const int *myThid,
                        const double *uVel, const double *vVel, const double *wVel,
                        double *gU, double *gV);

/* ---- MOM_VECINV ------------------------------------------
 * Same argument list as pkg/mom_vecinv/mom_vecinv.F:10-16 (caller dynamics.F:527), followed by the
 * COMMON /DYNVARS_R/ arrays and by what MOM_CALC_RELVORT3 (mom_calc_relvort3.F:79-97) and
 * FILL_CS_CORNER_TR_RL read from W2_EXCH2_TOPOLOGY.h for this tile: csCorners = bit mask of the
 * facet corners it owns (1 SW, 2 SE, 4 NE, 8 NW; 0 when .NOT.useCubedSphereExchange) and
 * myFace = exch2_myFace(W2_myTileList(bi,bj)).  Run-time switches: MI_USECORIOLIS,
 * MI_USEABSVORTICITY, MI_SELECTVORTSCHEME, MI_SELECTKESCHEME, MI_SELECTCORISCHEME,
 * MI_USEJAMARTMOMADV, MI_UPWINDSHEAR and the viscosity / drag parameters of MOM_FLUXFORM. */
void mom_vecinv_b200_(const int *bi, const int *bj, const int *k, const int *iMin, const int *iMax,
                      const int *jMin, const int *jMax,
                      const double *kappaRU, const double *kappaRV,
                      const double *fVerUkm, const double *fVerVkm, double *fVerUkp, double *fVerVkp,
                      double *guDiss, double *gvDiss,
                      const double *myTime, const int *myIter, const int *myThid,
                      const double *uVel, const double *vVel, const double *wVel,
                      double *gU, double *gV, const int *csCorners, const int *myFace);

/* ---- MOM_U_IMPLICIT_R / MOM_V_IMPLICIT_R --------------------------
 * Argument list of pkg/mom_common/mom_{u,v}_implicit_r.F:6-8 (caller dynamics.F:576-579) followed by the COMMON
 * /DYNVARS_R/ array the routine solves in place (u* in gU, v* in gV).  implicitViscosity only: the tridiagonal of
 * the vertical viscosity, SOLVE_TRIDIAGONAL per column; momImplVertAdv / selectImplicitDrag >= 1 are refused. */
void mom_u_implicit_r_b200_(const double *kappaRU, const int *bi, const int *bj, const double *myTime,
                            const int *myIter, const int *myThid, double *gU);
void mom_v_implicit_r_b200_(const double *kappaRV, const int *bi, const int *bj, const double *myTime,
                            const int *myIter, const int *myThid, double *gV);

/* Facet data of the local tiles (tiles bi fast) for the entry points that cannot take them as arguments -- the
 * resident step with MI_VECTORINVARIANTMOMENTUM and gad_advection_b200_ on the cubed sphere: csCorners and
 * myFace as in mom_vecinv_b200_, edges = 1 N | 2 S | 4 E | 8 W facet edges the tile touches
 * (exch2_is{N,S,E,W}edge, gad_advection.F:253-257). */
void mitgcm_b200_set_cs_tiles_(const int *csCorners, const int *myFace, const int *edges, int *ierr);

/* ---- resident time step ----------------------------------
 * One model step on the device mirrors in the order of model/src/forward_step.F (non-staggered):
 * THERMODYNAMICS (TEMP_INTEGRATE: CALC_ADV_FLOW + GAD_CALC_RHS + ADAMS_BASHFORTH2 + TIMESTEP_TRACER
 * + CYCLE_TRACER), DYNAMICS (MOM_FLUXFORM + TIMESTEP), SOLVE_FOR_PRESSURE (CALC_DIV_GHAT + CG2D),
 * MOMENTUM_CORRECTION_STEP, INTEGR_CONTINUITY, DO_FIELDS_BLOCKING_EXCHANGES.  State and forcing
 * are the mirrors MG_UVEL.. MG_ETAN, MG_SURFFORCU/V; nothing crosses PCIe.
 * Optional physics of the wider configurations (verification/tutorial_baroclinic_gyre), each switched by
 * its PARAMS.h flag: MI_BUOYANCYLINEAR (eosType = 'LINEAR': DO_OCEANIC_PHYS density + CALC_IVDC when
 * MP_IVDC_KAPPA != 0, CALC_PHI_HYD + CALC_GRAD_PHI_HYD in DYNAMICS), MI_DOTHETACLIMRELAX (FORCING_SURF_RELAX
 * towards MG_SST with MG_LAMBDATHETACLIMRELAX, APPLY_FORCING_T), MI_IMPLICITDIFFUSION (GAD_IMPLICIT_R +
 * SOLVE_TRIDIAGONAL), MI_EXACTCONSERV (etaH in the solver right-hand side, dEtaHdt / etaN / etaH update
 * of INTEGR_CONTINUITY + UPDATE_ETAH).  The solver scalars of
 * the step are returned like SOLVE_FOR_PRESSURE prints them (solve_for_pressure.F:337-348). */
void mitgcm_b200_forward_step_(const int *myIter, double *cg2d_init_res, int *cg2d_iters,
                               double *cg2d_last_res, int *ierr);
/* Column geometry of the resident step (csrc/colgeom.cu): 1 = the nine mirrors hFacC/W/S, recip_hFacC/W/S, maskC/W/S
 * have the z-level form (1 down to the deepest wet level, one partial value there, 0 below; mask = hFac != 0) and the
 * step kernels rebuild them per column instead of reading the 3-D arrays; 0 = not checked since a mirror last changed
 * (the check runs at the next forward step); -1 = they do not (general kernels).  MITGCM_B200_NO_COLGEOM=1 disables it. */
int mitgcm_b200_col_geom_state_(void);
/* EXCH_XYZ_RL / EXCH_XY_RL on a mirror (eesupp/src/exch_xyz_rx.template, exch_xy_rx.template):
 * full-width halo with corners. */
void mitgcm_b200_exch_(const int *id, int *ierr);

/* ---- pkg/exch2 tile graph (cubed sphere and other multi-facet grids) ---------------------------
 * Hands over the topology the model holds in COMMON /W2_EXCH2_TOPO_I/ and /W2_EXCH2_HALO_SPEC/
 * (pkg/exch2/W2_EXCH2_TOPOLOGY.h:63-122, filled by W2_E2SETUP) plus W2_myTileList (:162), arrays in
 * Fortran order: (W2_maxNeighbours, nTiles), pij (4, W2_maxNeighbours, nTiles).  From then on the
 * width-1 exchange inside CG2D / CG2D_SR follows EXCH2_S3D_RX (exch2_s3d_rx.template) and
 * mitgcm_b200_exch_ follows EXCH2_3D_RX (exch2_3d_rx.template: IGNORE_CORNERS pass, then
 * UPDATE_CORNERS pass) instead of the periodic nSx x nSy tiling.
 * One process: all nTiles = nSx*nSy tiles live on this GPU.  Several processes (nPx*nPy > 1 in mitgcm_b200_init_,
 * peers connected): nTiles is the size of the whole graph, every rank holds nTiles / nRanks of them (W2_myTileList in
 * increasing tile id) and mitgcm_b200_set_exch2_tile_proc_ must have named the owners; halo cells whose source tile
 * lives on another rank are then read from that rank's peer arena over NVLink -- the MPI messages of
 * exch2_send_rx{1,2}.template / exch2_recv_rx{1,2}.template -- and CG2D pushes its edge values into the owner's halo. */
void mitgcm_b200_set_exch2_topology_(
    const int *nTiles, const int *maxNeighbours, const int *nNeighbours, const int *neighbourId,
    const int *opposingSend, const int *neighbourDir, const int *pij, const int *oi, const int *oj,
    const int *iLo, const int *iHi, const int *jLo, const int *jHi, const int *tBasex, const int *tBasey,
    const int *isNedge, const int *isSedge, const int *isEedge, const int *isWedge,
    const int *myTileList, int *ierr);
/* W2_tileProc(nTiles) (pkg/exch2/W2_EXCH2_TOPOLOGY.h:153-165, set by w2_map_procs.F:91): 1-based number of the process
 * that owns each tile; process p is rank p-1 of mitgcm_b200_comm_connect_.  Call before mitgcm_b200_set_exch2_topology_
 * when the graph is spread over several ranks. */
void mitgcm_b200_set_exch2_tile_proc_(const int *nTiles, const int *tileProc, int *ierr);
/* EXCH_UV_XY_RL/RS, EXCH_UV_XYZ_RL/RS on a pair of mirrors (eesupp/src/exch_uv_xyz_rx.template): on an exch2
 * tile graph EXCH2_UV_3D_RX (pkg/exch2/exch2_uv_3d_rx.template: two EXCH2_RX2_CUBE passes with the C-grid index
 * offsets of exch2_get_uv_bounds.F, u/v swapped and, when withSigns, negated across rotated facet edges, then
 * the cube-corner fix-ups); on the plain periodic tiling two scalar exchanges. */
void mitgcm_b200_exch_uv_(const int *idU, const int *idV, const int *withSigns, int *ierr);
/* Host-only helper (no device needed): the lists rank `myRank` of `nRanks` is given when the tile graph is spread
 * over ranks (what mitgcm_b200_set_exch2_topology_ uploads there): scalar gather list (2 ints per entry: dst,
 * owner << 28 | src), the width-1 push table of its tiles (owner << 28 | halo index) and the vector-pair gather list
 * (4 ints per entry), indices relative to the owning rank's (nTiles / nRanks, sNy+2*OL, sNx+2*OL) arrays.
 * dims3 = (sNx, sNy, OL); sizes(3): capacities in ints on entry, ints written on return. */
void mitgcm_b200_exch2_dist_lists_(
    const int *dims3, const int *nRanks, const int *myRank, const int *tileProc, const int *withSigns, const int *nTiles,
    const int *maxNeighbours, const int *nNeighbours, const int *neighbourId, const int *opposingSend,
    const int *neighbourDir, const int *pij, const int *oi, const int *oj, const int *iLo, const int *iHi, const int *jLo,
    const int *jHi, const int *tBasex, const int *tBasey, const int *isNedge, const int *isSedge, const int *isEedge,
    const int *isWedge, int *sizes, int *scalarOut, int *pushOut, int *uvOut, int *ierr);
/* Host-only helper (no device needed): the compiled vector-pair exchange of a tile graph as a list of 4-int
 * entries (dst array 0/1, dst flat cell, src array << 1 | negate, src flat cell) over (nTiles, sNy+2*OL,
 * sNx+2*OL) arrays; dims3 = {sNx, sNy, OL}.  Used by the host set-up code and the CPU tests. */
void mitgcm_b200_exch2_uv_map_(
    const int *dims3, const int *withSigns, const int *nTiles, const int *maxNeighbours, const int *nNeighbours,
    const int *neighbourId, const int *opposingSend, const int *neighbourDir, const int *pij, const int *oi,
    const int *oj, const int *iLo, const int *iHi, const int *jLo, const int *jHi, const int *tBasex,
    const int *tBasey, const int *isNedge, const int *isSedge, const int *isEedge, const int *isWedge,
    const int *maxEntries, int *nEntries, int *out4, int *ierr);

/* ---- multi-GPU (one process per GPU on one NVSwitch domain) --------------------------------
 * With nPx*nPy > 1 every rank keeps what its neighbours write into -- the exchanged state fields (uVel, vVel,
 * wVel, theta, salt, cg2d_x, etaN, etaH) and the CG2D workspace -- in one allocation, the peer arena, exports it
 * as a CUDA IPC handle (64 bytes + the 8-byte offset of the arena inside its allocation = 72 bytes) and, after
 * the handles of all ranks have been gathered (MPI_Allgather in the Fortran model, torch.distributed here), maps
 * its peers.  From then on
 *   - CG2D stores edge values directly into the neighbour's halo cells and combines dot products through
 *     peer-mapped mailboxes inside the persistent kernel (replaces EXCH_S3D_RL + GLOBAL_SUM_TILE_RL);
 *   - mitgcm_b200_halo_exchange_ fills the full-width halos (with corners) of several mirrors at once by storing
 *     edge strips and corner blocks into the 8 neighbours' halos over NVLink, ordered by peer-memory flags
 *     (replaces EXCH_XY_RL / EXCH_XYZ_RL, eesupp/src/exch1_rx.template:170-201): no MPI, no host round trips;
 *   - mitgcm_b200_forward_step_ runs across ranks, theta's halo travelling on a side stream under DYNAMICS.
 * rank = myPx + nPx*myPy (MPI_CART_CREATE order, eesupp/src/ini_procs.F:145).  The caller barriers between
 * connect and the first exchange. */
void mitgcm_b200_comm_handle_(unsigned char *handle72, int *ierr);
void mitgcm_b200_comm_connect_(const int *nRanks, const int *myRank, const unsigned char *handles72, int *ierr);
void mitgcm_b200_halo_exchange_(const int *nFields, const int *ids, int *ierr);
/* Unmaps the peers' arenas.  Shut-down order across ranks: disconnect on every rank, barrier, then
 * mitgcm_b200_finalize_ (which frees the arena the peers had mapped). */
void mitgcm_b200_comm_disconnect_(void);
/* The same exchanges over NCCL send/recv (alternative transport, mitgcm_b200/distributed.py): pack the strip
 * to send towards dir (0 W, 1 E, 2 S, 3 N) into buf, or (unpack != 0) scatter a received strip
 * into the halo on side dir; exch_dir does one periodic direction locally (nPx or nPy == 1). */
void mitgcm_b200_pack_(const int *id, const int *dir, double *buf, const int *unpack, int *ierr);
void mitgcm_b200_exch_dir_(const int *id, const int *ydir, int *ierr);
/* The resident step in two parts with the halo exchanges left to the caller (see step.cu). */
void mitgcm_b200_step_part_(const int *part, const int *myIter, double *cg2d_init_res, int *cg2d_iters,
                            double *cg2d_last_res, int *ierr);

#ifdef __cplusplus
}
#endif
#endif
