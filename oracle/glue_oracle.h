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
                 double *gU, double *gV, double *guNm1, double *gvNm1);
void og_solve_rhs(const og_grid *g, const og_params *p, int bi, int bj, const double *Bo_surf,
                  const double *etaN, const double *gU, const double *gV,
                  double *cg2d_b, double *cg2d_x);
void og_correction_step(const og_grid *g, const og_params *p, int bi, int bj, const double *Bo_surf,
                        const double *etaN, const double *gU, const double *gV,
                        double *uVel, double *vVel);
void og_integrate_for_w(const og_grid *g, const og_params *p, int bi, int bj,
                        const double *uVel, const double *vVel, double *wVel);
#ifdef __cplusplus
}
#endif
#endif
