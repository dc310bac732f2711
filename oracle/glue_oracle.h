/* glue_oracle.h -- step-glue restatements (see glue_oracle.c). TEST INFRASTRUCTURE ONLY. */
#ifndef GLUE_ORACLE_H
#define GLUE_ORACLE_H
#include "mitgcm_oracle.h"
#ifdef __cplusplus
extern "C" {
#endif
void og_timestep(const og_grid *g, const og_params *p, int bi, int bj, int k,
                 int iMin, int iMax, int jMin, int jMax,
                 const double *dPhiHydX, const double *dPhiHydY,
                 const double *guDissip, const double *gvDissip,
                 const double *surfaceForcingU, const double *surfaceForcingV,
                 int momForcing, int momDissip_In_AB, double abFac,
                 const double *uVel, const double *vVel,
                 double *gU, double *gV, double *guNm1, double *gvNm1,
                 const double *phiSurfX, const double *phiSurfY);   /* NULL unless implicSurfPress != 1 */
void og_calc_grad_phi_surf(const og_grid *g, int bi, int bj, int iMin, int iMax, int jMin, int jMax,
                           const double *Bo_surf, const double *etaFld, double *phiSurfX, double *phiSurfY);
/* etaFS = the field of the free-surface term: etaH when exactConserv (solve_for_pressure.F:213-222),
 * else etaN (:224-233). */
void og_solve_rhs(const og_grid *g, const og_params *p, int bi, int bj, const double *Bo_surf,
                  const double *etaN, const double *etaFS, const double *gU, const double *gV,
                  double *cg2d_b, double *cg2d_x);
void og_correction_step(const og_grid *g, const og_params *p, int bi, int bj, const double *Bo_surf,
                        const double *etaN, const double *gU, const double *gV,
                        double *uVel, double *vVel);
void og_integrate_for_w(const og_grid *g, const og_params *p, int bi, int bj,
                        const double *uVel, const double *vVel, double *wVel);

/* ---- physics glue of config 2 (phys_oracle.c) ---- */
typedef struct { double rhoNil, rhoConst, tAlpha, sBeta; } og_eos;
void og_density_ivdc(const og_grid *g, const og_params *p, const og_eos *e, int bi, int bj,
                     const double *theta, const double *salt, const double *tRef, const double *sRef,
                     double *rhoInSitu, double *IVDConvCount);
void og_forcing_surf_relax_T(const og_grid *g, int bi, int bj, const double *theta, const double *SST,
                             const double *lambdaThetaClimRelax, double recip_Cp, double mass2rUnit,
                             double *surfaceForcingT);
void og_apply_forcing_T(const og_grid *g, int bi, int bj, int k, const double *surfaceForcingT, double *gtForc);
void og_calc_3d_diffusivity(const og_grid *g, int bi, int bj, const double *IVDConvCount, double ivdc_kappa,
                            const double *KbryanLewis79, const double *diffKrNrT, double *kappaRk);
int og_mom_implicit_r(const og_grid *g, const og_params *p, int bi, int bj, int isV, const double *kappaR, double *gFld);
int og_gad_implicit_r(const og_grid *g, int bi, int bj, int iMin, int iMax, int jMin, int jMax,
                      const double *deltaTLev, const double *kappaRX, const double *recip_hFac, double *gTracer);
void og_calc_phi_hyd(const og_grid *g, int bi, int bj, int iMin, int iMax, int jMin, int jMax, int k,
                     const double *rhoInSitu, const double *rF, const double *rC, double gravity,
                     double recip_rhoConst, const double *phi0surf,
                     double *phiHydF, double *phiHydC, double *dPhiHydX, double *dPhiHydY);
void og_integr_continuity_ec(const og_grid *g, const og_params *p, int bi, int bj, const double *uVel,
                             const double *vVel, const double *etaH, double *dEtaHdt, double *etaN,
                             int updateEtaN);

/* ---- the non-hydrostatic step around CG3D (nh_oracle.c) ---- */
int og_calc_gw(const og_grid *g, const og_params *p, int bi, int bj, const double *R_low, const double *Ro_surf,
               const double *rLowW, const double *rSurfW, const double *rLowS, const double *rSurfS,
               const double *rC, const double *kappaRU, const double *kappaRV, double viscAhW, double viscA4W,
               int momDissip_In_AB, double abFac, const double *uVel, const double *vVel, const double *wVel,
               double *gW, double *gwNm1);
void og_timestep_wvel(const og_grid *g, const og_params *p, int bi, int bj, double *gW, double *wVel);
void og_solve_rhs_nh(const og_grid *g, const og_params *p, int bi, int bj, const double *Bo_surf, const double *etaN,
                     const double *phi_nh, const double *gU, const double *gV, double *cg2d_b, double *cg2d_x,
                     double *cg3d_b);
void og_pre_cg3d(const og_grid *g, const og_params *p, int bi, int bj, const double *cg2d_x, const double *etaN,
                 const double *wVel, double *cg3d_b);
void og_correction_step_nh(const og_grid *g, const og_params *p, int bi, int bj, const double *Bo_surf, const double *etaN,
                           const double *phi_nh, const double *gU, const double *gV, double *uVel, double *vVel);
#ifdef __cplusplus
}
#endif
#endif
