// mom.cuh -- point-wise restatement of MOM_FLUXFORM (pkg/mom_fluxform/mom_fluxform.F:202-1051) and
// the leaves it calls (mom_{u,v}_adv_*.F, mom_calc_rtrans.F, mom_{u,v}_{x,y}viscflux.F,
// mom_{u,v}_del2{u,v}.F, mom_{u,v}_coriolis.F, mom_{u,v}_metric_sphere.F; pkg/mom_common:
// mom_calc_hfacz.F, mom_calc_ke.F, mom_{u,v}_rviscflux.F, mom_{u,v}_sidedrag.F,
// mom_{u,v}_botdrag_coeff.F).  As in gad.cuh every face flux is a pure function of the inputs
// and is re-evaluated where it is needed instead of being staged through ~30 slab temporaries
// (mom_fluxform.F:123-166).  Expression order follows the Fortran (-fmad=false).
// Not on the B200 path (rejected by the entry point): variable viscosity, NH / 3-D Coriolis
// metric terms, cylindrical grid, r* and sigma coordinates, OBCS, shelf ice,
// sideDragFactor <= 0, deep-atmosphere / anelastic factors.
#pragma once
#include "gad.cuh"

namespace mg {

struct MomPar {
  double viscAhD, viscAhZ, viscA4D, viscA4Z, sideDragFactor, bottomDragLinear, bottomDragQuadratic, recip_rSphere;
  double afFacMom, vfFacMom, cfFacMom, mtFacMom, rkSign;
  int momAdvection, momViscosity, useBiharmonicVisc, implicitViscosity, no_slip_sides, no_slip_bottom,
      bottomVisc_pCell, selectBotDragQuadr, bottomDragTerms, useCDscheme, selectCoriScheme, metricTerms,
      usingSphericalPolarGrid, rigidLid, select_rStar;
};

// 3-D state of one tile (COMMON /DYNVARS_R/ uVel,vVel,wVel; kappaRU/V have Nr+1 levels).
struct MomState {
  const double *u, *v, *w, *kapU, *kapV;
};

#define MU(i, j, kk) st.u[g.s3(i, j, kk)]
#define MV(i, j, kk) st.v[g.s3(i, j, kk)]
#define MW(i, j, kk) st.w[g.s3(i, j, kk)]

// MOM_CALC_HFACZ (hZoption = 0): open-water fraction at the vorticity point.
__device__ __forceinline__ double mom_hfacz(const TileGrid &g, int k, int i, int j) {
  if (i < 2 - g.OLx || j < 2 - g.OLy) return 0.;
  double h = fmin(g.hFacW[g.s3(i, j, k)], g.hFacW[g.s3(i, j - 1, k)]);
  h = fmin(g.hFacS[g.s3(i, j, k)], h);
  h = fmin(g.hFacS[g.s3(i - 1, j, k)], h);
  return h;
}
__device__ __forceinline__ double mom_xA(const TileGrid &g, int k, int i, int j) {
  return g.dyG[g.s(i, j)] * g.drF[k - 1] * g.hFacW[g.s3(i, j, k)];
}
__device__ __forceinline__ double mom_yA(const TileGrid &g, int k, int i, int j) {
  return g.dxG[g.s(i, j)] * g.drF[k - 1] * g.hFacS[g.s3(i, j, k)];
}
__device__ __forceinline__ double mom_uTrans(const TileGrid &g, const MomState &st, int k, int i, int j) {
  return MU(i, j, k) * mom_xA(g, k, i, j);
}
__device__ __forceinline__ double mom_vTrans(const TileGrid &g, const MomState &st, int k, int i, int j) {
  return MV(i, j, k) * mom_yA(g, k, i, j);
}
// MOM_CALC_KE with the literal KEscheme = 2 of mom_fluxform.F:329
__device__ __forceinline__ double mom_ke(const TileGrid &g, const MomState &st, int k, int i, int j) {
  double u0 = MU(i, j, k), u1 = MU(i + 1, j, k), v0 = MV(i, j, k), v1 = MV(i, j + 1, k);
  return 0.25 * ((u0 * u0 * g.hFacW[g.s3(i, j, k)] + u1 * u1 * g.hFacW[g.s3(i + 1, j, k)]) +
                 (v0 * v0 * g.hFacS[g.s3(i, j, k)] + v1 * v1 * g.hFacS[g.s3(i, j + 1, k)])) *
         g.recip_hFacC[g.s3(i, j, k)];
}
// MOM_CALC_RTRANS at interface kk (1..Nr+1)
__device__ __forceinline__ double mom_rTransU(const TileGrid &g, const MomState &st, int kk, int i, int j) {
  if (kk > g.Nr) return 0.;
  return 0.5 * (MW(i - 1, j, kk) * g.rA[g.s(i - 1, j)] + MW(i, j, kk) * g.rA[g.s(i, j)]);
}
__device__ __forceinline__ double mom_rTransV(const TileGrid &g, const MomState &st, int kk, int i, int j) {
  if (kk > g.Nr) return 0.;
  return 0.5 * (MW(i, j - 1, kk) * g.rA[g.s(i, j - 1)] + MW(i, j, kk) * g.rA[g.s(i, j)]);
}
// MOM_U_ADV_WU / MOM_V_ADV_WV at interface kk
__device__ inline double mom_adv_wu(const TileGrid &g, const MomState &st, const MomPar &p, int kk, int i, int j) {
  if (kk > g.Nr || (kk == 1 && p.rigidLid)) return 0.;
  const double rT = mom_rTransU(g, st, kk, i, j);
  if (kk == 1) return rT * MU(i, j, kk);
  double f = rT * 0.5 * (MU(i, j, kk) + MU(i, j, kk - 1));
  if (p.select_rStar == 0 && !p.rigidLid)
    f = f + 0.25 * (MW(i, j, kk) * g.rA[g.s(i, j)] * (g.maskC[g.s3(i, j, kk)] - g.maskC[g.s3(i, j, kk - 1)]) +
                    MW(i - 1, j, kk) * g.rA[g.s(i - 1, j)] * (g.maskC[g.s3(i - 1, j, kk)] - g.maskC[g.s3(i - 1, j, kk - 1)])) *
                MU(i, j, kk);
  return f;
}
__device__ inline double mom_adv_wv(const TileGrid &g, const MomState &st, const MomPar &p, int kk, int i, int j) {
  if (kk > g.Nr || (kk == 1 && p.rigidLid)) return 0.;
  const double rT = mom_rTransV(g, st, kk, i, j);
  if (kk == 1) return rT * MV(i, j, kk);
  double f = rT * 0.5 * (MV(i, j, kk) + MV(i, j, kk - 1));
  if (p.select_rStar == 0 && !p.rigidLid)
    f = f + 0.25 * (MW(i, j, kk) * g.rA[g.s(i, j)] * (g.maskC[g.s3(i, j, kk)] - g.maskC[g.s3(i, j, kk - 1)]) +
                    MW(i, j - 1, kk) * g.rA[g.s(i, j - 1)] * (g.maskC[g.s3(i, j - 1, kk)] - g.maskC[g.s3(i, j - 1, kk - 1)])) *
                MV(i, j, kk);
  return f;
}
// horizontal advective fluxes
__device__ __forceinline__ double mom_adv_uu(const TileGrid &g, const MomState &st, int k, int i, int j) {
  return 0.25 * (mom_uTrans(g, st, k, i, j) + mom_uTrans(g, st, k, i + 1, j)) * (MU(i, j, k) + MU(i + 1, j, k));
}
__device__ __forceinline__ double mom_adv_vu(const TileGrid &g, const MomState &st, int k, int i, int j) {
  return 0.25 * (mom_vTrans(g, st, k, i, j) + mom_vTrans(g, st, k, i - 1, j)) * (MU(i, j, k) + MU(i, j - 1, k));
}
__device__ __forceinline__ double mom_adv_uv(const TileGrid &g, const MomState &st, int k, int i, int j) {
  return 0.25 * (mom_uTrans(g, st, k, i, j) + mom_uTrans(g, st, k, i, j - 1)) * (MV(i, j, k) + MV(i - 1, j, k));
}
__device__ __forceinline__ double mom_adv_vv(const TileGrid &g, const MomState &st, int k, int i, int j) {
  return 0.25 * (mom_vTrans(g, st, k, i, j) + mom_vTrans(g, st, k, i, j + 1)) * (MV(i, j, k) + MV(i, j + 1, k));
}
// MOM_U_RVISCFLUX / MOM_V_RVISCFLUX at interface kk
__device__ __forceinline__ double mom_u_rvisc(const TileGrid &g, const MomState &st, const MomPar &p, int kk, int i, int j) {
  if (kk <= 1 || kk > g.Nr) return 0.;
  return -st.kapU[g.s3(i, j, kk)] * g.rAw[g.s(i, j)] * (MU(i, j, kk) - MU(i, j, kk - 1)) * p.rkSign * g.recip_drC[kk - 1] *
         g.maskW[g.s3(i, j, kk)] * g.maskW[g.s3(i, j, kk - 1)];
}
__device__ __forceinline__ double mom_v_rvisc(const TileGrid &g, const MomState &st, const MomPar &p, int kk, int i, int j) {
  if (kk <= 1 || kk > g.Nr) return 0.;
  return -st.kapV[g.s3(i, j, kk)] * g.rAs[g.s(i, j)] * (MV(i, j, kk) - MV(i, j, kk - 1)) * p.rkSign * g.recip_drC[kk - 1] *
         g.maskS[g.s3(i, j, kk)] * g.maskS[g.s3(i, j, kk - 1)];
}
// MOM_U_DEL2U / MOM_V_DEL2V (zero outside 2-OL..sN+OL-1, like the zero-initialised v4F)
__device__ inline double mom_del2u(const TileGrid &g, const MomState &st, const MomPar &p, int k, int i, int j) {
  if (!p.useBiharmonicVisc) return 0.;
  if (i < 2 - g.OLx || i > g.sNx + g.OLx - 1 || j < 2 - g.OLy || j > g.sNy + g.OLy - 1) return 0.;
  auto fZ = [&](int ii) {
    return g.drF[k - 1] * g.hFacC[g.s3(ii, j, k)] * g.dyF[g.s(ii, j)] * g.recip_dxF[g.s(ii, j)] * (MU(ii + 1, j, k) - MU(ii, j, k));
  };
  auto fM = [&](int jj) {
    return g.drF[k - 1] * mom_hfacz(g, k, i, jj) * g.dxV[g.s(i, jj)] * g.recip_dyU[g.s(i, jj)] * (MU(i, jj, k) - MU(i, jj - 1, k));
  };
  double d = g.recip_drF[k - 1] * g.recip_hFacW[g.s3(i, j, k)] * g.recip_rAw[g.s(i, j)] *
             (fZ(i) - fZ(i - 1) + fM(j + 1) - fM(j)) * g.maskW[g.s3(i, j, k)];
  if (p.no_slip_sides) {
    double hS = g.hFacW[g.s3(i, j, k)] - mom_hfacz(g, k, i, j);
    double hN = g.hFacW[g.s3(i, j, k)] - mom_hfacz(g, k, i, j + 1);
    d = d - g.recip_hFacW[g.s3(i, j, k)] * g.recip_rAw[g.s(i, j)] *
                (hS * g.dxV[g.s(i, j)] * g.recip_dyU[g.s(i, j)] + hN * g.dxV[g.s(i, j + 1)] * g.recip_dyU[g.s(i, j + 1)]) *
                MU(i, j, k) * p.sideDragFactor * g.maskW[g.s3(i, j, k)];
  }
  return d;
}
__device__ inline double mom_del2v(const TileGrid &g, const MomState &st, const MomPar &p, int k, int i, int j) {
  if (!p.useBiharmonicVisc) return 0.;
  if (i < 2 - g.OLx || i > g.sNx + g.OLx - 1 || j < 2 - g.OLy || j > g.sNy + g.OLy - 1) return 0.;
  auto fZ = [&](int ii) {
    return g.drF[k - 1] * mom_hfacz(g, k, ii, j) * g.dyU[g.s(ii, j)] * g.recip_dxV[g.s(ii, j)] * (MV(ii, j, k) - MV(ii - 1, j, k));
  };
  auto fM = [&](int jj) {
    return g.drF[k - 1] * g.hFacC[g.s3(i, jj, k)] * g.dxF[g.s(i, jj)] * g.recip_dyF[g.s(i, jj)] * (MV(i, jj + 1, k) - MV(i, jj, k));
  };
  double d = g.recip_drF[k - 1] * g.recip_hFacS[g.s3(i, j, k)] * g.recip_rAs[g.s(i, j)] *
             (fZ(i + 1) - fZ(i) + fM(j) - fM(j - 1)) * g.maskS[g.s3(i, j, k)];
  if (p.no_slip_sides) {
    double hW = g.hFacS[g.s3(i, j, k)] - mom_hfacz(g, k, i, j);
    double hE = g.hFacS[g.s3(i, j, k)] - mom_hfacz(g, k, i + 1, j);
    d = d - g.recip_hFacS[g.s3(i, j, k)] * g.recip_rAs[g.s(i, j)] *
                (hW * g.dyU[g.s(i, j)] * g.recip_dxV[g.s(i, j)] + hE * g.dyU[g.s(i + 1, j)] * g.recip_dxV[g.s(i + 1, j)]) *
                MV(i, j, k) * p.sideDragFactor * g.maskS[g.s3(i, j, k)];
  }
  return d;
}
// viscous fluxes
__device__ __forceinline__ double mom_u_xvisc(const TileGrid &g, const MomState &st, const MomPar &p, int k, int i, int j) {
  const double cf = g.cosFacU[j + g.OLy - 1];
  return g.dyF[g.s(i, j)] * g.drF[k - 1] * g.hFacC[g.s3(i, j, k)] *
         (-p.viscAhD * (MU(i + 1, j, k) - MU(i, j, k)) * cf +
          p.viscA4D * (mom_del2u(g, st, p, k, i + 1, j) - mom_del2u(g, st, p, k, i, j)) * cf) *
         g.recip_dxF[g.s(i, j)];
}
__device__ __forceinline__ double mom_u_yvisc(const TileGrid &g, const MomState &st, const MomPar &p, int k, int i, int j) {
  return g.dxV[g.s(i, j)] * g.drF[k - 1] * mom_hfacz(g, k, i, j) *
         (-p.viscAhZ * (MU(i, j, k) - MU(i, j - 1, k)) + p.viscA4Z * (mom_del2u(g, st, p, k, i, j) - mom_del2u(g, st, p, k, i, j - 1))) *
         g.recip_dyU[g.s(i, j)];
}
__device__ __forceinline__ double mom_v_xvisc(const TileGrid &g, const MomState &st, const MomPar &p, int k, int i, int j) {
  const double cf = g.cosFacV[j + g.OLy - 1];
  return g.dyU[g.s(i, j)] * g.drF[k - 1] * mom_hfacz(g, k, i, j) *
         (-p.viscAhZ * (MV(i, j, k) - MV(i - 1, j, k)) * cf +
          p.viscA4Z * (mom_del2v(g, st, p, k, i, j) - mom_del2v(g, st, p, k, i - 1, j)) * cf) *
         g.recip_dxV[g.s(i, j)];
}
__device__ __forceinline__ double mom_v_yvisc(const TileGrid &g, const MomState &st, const MomPar &p, int k, int i, int j) {
  return g.dxF[g.s(i, j)] * g.drF[k - 1] * g.hFacC[g.s3(i, j, k)] *
         (-p.viscAhD * (MV(i, j + 1, k) - MV(i, j, k)) + p.viscA4D * (mom_del2v(g, st, p, k, i, j + 1) - mom_del2v(g, st, p, k, i, j))) *
         g.recip_dyF[g.s(i, j)];
}
// MOM_U_SIDEDRAG / MOM_V_SIDEDRAG, sideDragFactor > 0 branch, constant viscosity
__device__ inline double mom_u_sidedrag(const TileGrid &g, const MomState &st, const MomPar &p, int k, int i, int j) {
  double hS = g.hFacW[g.s3(i, j, k)] - mom_hfacz(g, k, i, j);
  double hN = g.hFacW[g.s3(i, j, k)] - mom_hfacz(g, k, i, j + 1);
  double t = p.viscAhZ * MU(i, j, k) - p.viscA4Z * mom_del2u(g, st, p, k, i, j);
  return -g.recip_hFacW[g.s3(i, j, k)] * g.recip_drF[k - 1] * g.recip_rAw[g.s(i, j)] *
         (hS * g.dxV[g.s(i, j)] * g.recip_dyU[g.s(i, j)] * t + hN * g.dxV[g.s(i, j + 1)] * g.recip_dyU[g.s(i, j + 1)] * t) *
         g.drF[k - 1] * p.sideDragFactor;
}
__device__ inline double mom_v_sidedrag(const TileGrid &g, const MomState &st, const MomPar &p, int k, int i, int j) {
  const double cf = g.cosFacV[j + g.OLy - 1];
  double hW = g.hFacS[g.s3(i, j, k)] - mom_hfacz(g, k, i, j);
  double hE = g.hFacS[g.s3(i, j, k)] - mom_hfacz(g, k, i + 1, j);
  double t = p.viscAhZ * MV(i, j, k) * cf - p.viscA4Z * mom_del2v(g, st, p, k, i, j) * cf;
  return -g.recip_hFacS[g.s3(i, j, k)] * g.recip_drF[k - 1] * g.recip_rAs[g.s(i, j)] *
         (hW * g.dyU[g.s(i, j)] * g.recip_dxV[g.s(i, j)] * t + hE * g.dyU[g.s(i + 1, j)] * g.recip_dxV[g.s(i + 1, j)] * t) *
         g.drF[k - 1] * p.sideDragFactor;
}
// MOM_U_BOTDRAG_COEFF / MOM_V_BOTDRAG_COEFF (z coordinates, inp_KE = .TRUE.)
// haveKs: MOM_VECINV passes KE(i,j) + KE(i-di,j-dj) computed with selectKEscheme; MOM_FLUXFORM uses KEscheme = 2
__device__ inline double mom_botdrag(const TileGrid &g, const MomState &st, const MomPar &p, int k, int isV, int i, int j,
                                 bool haveKs = false, double ksIn = 0.) {
  const int Nr = g.Nr;
  const double viscFac = p.no_slip_bottom ? 2. : 0.;
  const int kDown = min(k + 1, Nr), kLowF = k + 1;
  const double recDrC = (k == Nr) ? g.recip_drF[k - 1] : g.recip_drC[kLowF - 1];
  const double *mask = isV ? g.maskS : g.maskW;
  const int di = isV ? 0 : 1, dj = isV ? 1 : 0;
  double c = p.bottomDragLinear * 1.;
  const double kap = (isV ? st.kapV : st.kapU)[g.s3(i, j, kLowF)];
  if (p.no_slip_bottom && p.bottomVisc_pCell)
    c = c + kap * recDrC * viscFac * (isV ? g.recip_hFacS : g.recip_hFacW)[g.s3(i, j, k)];
  else if (p.no_slip_bottom)
    c = c + kap * recDrC * viscFac;
  if (p.selectBotDragQuadr == 0) {
    double ks = haveKs ? ksIn : mom_ke(g, st, k, i, j) + mom_ke(g, st, k, i - di, j - dj);
    if (ks > 0.) c = c + p.bottomDragQuadratic * sqrt(ks) * 1.;
  } else if (p.selectBotDragQuadr == 1 || p.selectBotDragQuadr == 2) {
    double uSq;
    if (!isV) {
      double a = (MV(i - 1, j, k) * MV(i - 1, j, k) * g.hFacS[g.s3(i - 1, j, k)] + MV(i, j, k) * MV(i, j, k) * g.hFacS[g.s3(i, j, k)]) +
                 (MV(i - 1, j + 1, k) * MV(i - 1, j + 1, k) * g.hFacS[g.s3(i - 1, j + 1, k)] +
                  MV(i, j + 1, k) * MV(i, j + 1, k) * g.hFacS[g.s3(i, j + 1, k)]);
      if (p.selectBotDragQuadr == 1) uSq = MU(i, j, k) * MU(i, j, k) + a * g.recip_hFacW[g.s3(i, j, k)] * 0.25;
      else {
        double h = (g.hFacS[g.s3(i - 1, j, k)] + g.hFacS[g.s3(i, j, k)]) + (g.hFacS[g.s3(i - 1, j + 1, k)] + g.hFacS[g.s3(i, j + 1, k)]);
        uSq = h > 0. ? MU(i, j, k) * MU(i, j, k) + a / h : MU(i, j, k) * MU(i, j, k);
      }
    } else {
      double a = (MU(i, j - 1, k) * MU(i, j - 1, k) * g.hFacW[g.s3(i, j - 1, k)] + MU(i, j, k) * MU(i, j, k) * g.hFacW[g.s3(i, j, k)]) +
                 (MU(i + 1, j - 1, k) * MU(i + 1, j - 1, k) * g.hFacW[g.s3(i + 1, j - 1, k)] +
                  MU(i + 1, j, k) * MU(i + 1, j, k) * g.hFacW[g.s3(i + 1, j, k)]);
      if (p.selectBotDragQuadr == 1) uSq = MV(i, j, k) * MV(i, j, k) + a * g.recip_hFacS[g.s3(i, j, k)] * 0.25;
      else {
        double h = (g.hFacW[g.s3(i, j - 1, k)] + g.hFacW[g.s3(i, j, k)]) + (g.hFacW[g.s3(i + 1, j - 1, k)] + g.hFacW[g.s3(i + 1, j, k)]);
        uSq = h > 0. ? MV(i, j, k) * MV(i, j, k) + a / h : MV(i, j, k) * MV(i, j, k);
      }
    }
    if (uSq > 0.) c = c + p.bottomDragQuadratic * sqrt(uSq) * 1.;
  }
  if (k == Nr) return c * mask[g.s3(i, j, k)];
  return c * mask[g.s3(i, j, k)] * (1. - mask[g.s3(i, j, kDown)]);
}

struct MomOut { double gU, gV, guDiss, gvDiss; };

// Tendencies of one cell (mom_fluxform.F:502-517, :602-662, :716-721, :757-772, :860-921,
// :975-980, :995-1022, :1044-1051) given the vertical advective fluxes above (km) and below (kp).
__device__ inline MomOut mom_cell(const TileGrid &g, const MomState &st, const MomPar &p, int k, int i, int j, double fVerUkm,
                           double fVerUkp, double fVerVkm, double fVerVkp) {
  MomOut o;
  const double rhW = g.recip_hFacW[g.s3(i, j, k)], rhS = g.recip_hFacS[g.s3(i, j, k)], rdrF = g.recip_drF[k - 1];
  const double uDudxFac = p.afFacMom, vDudyFac = p.afFacMom, rVelDudrFac = p.afFacMom;
  const double AhDudxFac = p.vfFacMom, AhDudyFac = p.vfFacMom;
  const double ArDudrFac = p.implicitViscosity ? 0. : p.vfFacMom;
  // ---- U ----
  if (p.momAdvection)
    o.gU = -rhW * rdrF * g.recip_rAw[g.s(i, j)] *
           ((mom_adv_uu(g, st, k, i, j) - mom_adv_uu(g, st, k, i - 1, j)) * uDudxFac +
            (mom_adv_vu(g, st, k, i, j + 1) - mom_adv_vu(g, st, k, i, j)) * vDudyFac + (fVerUkp - fVerUkm) * p.rkSign * rVelDudrFac);
  else o.gU = 0.;
  o.guDiss = 0.;
  if (p.momViscosity) {
    double fVrUp = 0., fVrDw = 0.;
    if (!p.implicitViscosity) { fVrUp = mom_u_rvisc(g, st, p, k, i, j); fVrDw = mom_u_rvisc(g, st, p, k + 1, i, j); }
    o.guDiss = -rhW * rdrF * g.recip_rAw[g.s(i, j)] *
               ((mom_u_xvisc(g, st, p, k, i, j) - mom_u_xvisc(g, st, p, k, i - 1, j)) * AhDudxFac +
                (mom_u_yvisc(g, st, p, k, i, j + 1) - mom_u_yvisc(g, st, p, k, i, j)) * AhDudyFac +
                (fVrDw - fVrUp) * p.rkSign * ArDudrFac);
    if (p.no_slip_sides) o.guDiss = o.guDiss + mom_u_sidedrag(g, st, p, k, i, j);
    if (p.bottomDragTerms) o.guDiss = o.guDiss - mom_botdrag(g, st, p, k, 0, i, j) * MU(i, j, k) * rhW * rdrF;
  }
  if (p.usingSphericalPolarGrid && p.metricTerms) {
    double mT = MU(i, j, k) * p.recip_rSphere * 0.25 * (MV(i, j, k) + MV(i - 1, j, k) + MV(i, j + 1, k) + MV(i - 1, j + 1, k)) *
                g.tanPhiAtU[g.s(i, j)];
    o.gU = o.gU + p.mtFacMom * mT;
  }
  // ---- V ----
  if (p.momAdvection)
    o.gV = -rhS * rdrF * g.recip_rAs[g.s(i, j)] *
           ((mom_adv_uv(g, st, k, i + 1, j) - mom_adv_uv(g, st, k, i, j)) * uDudxFac +
            (mom_adv_vv(g, st, k, i, j) - mom_adv_vv(g, st, k, i, j - 1)) * vDudyFac + (fVerVkp - fVerVkm) * p.rkSign * rVelDudrFac);
  else o.gV = 0.;
  o.gvDiss = 0.;
  if (p.momViscosity) {
    double fVrUp = 0., fVrDw = 0.;
    if (!p.implicitViscosity) { fVrUp = mom_v_rvisc(g, st, p, k, i, j); fVrDw = mom_v_rvisc(g, st, p, k + 1, i, j); }
    o.gvDiss = -rhS * rdrF * g.recip_rAs[g.s(i, j)] *
               ((mom_v_xvisc(g, st, p, k, i + 1, j) - mom_v_xvisc(g, st, p, k, i, j)) * AhDudxFac +
                (mom_v_yvisc(g, st, p, k, i, j) - mom_v_yvisc(g, st, p, k, i, j - 1)) * AhDudyFac +
                (fVrDw - fVrUp) * p.rkSign * ArDudrFac);
    if (p.no_slip_sides) o.gvDiss = o.gvDiss + mom_v_sidedrag(g, st, p, k, i, j);
    if (p.bottomDragTerms) o.gvDiss = o.gvDiss - mom_botdrag(g, st, p, k, 1, i, j) * MV(i, j, k) * rhS * rdrF;
  }
  if (p.usingSphericalPolarGrid && p.metricTerms) {
    double ub = 0.25 * (MU(i, j, k) + MU(i + 1, j, k) + MU(i, j - 1, k) + MU(i + 1, j - 1, k));
    double mT = -p.recip_rSphere * ub * ub * g.tanPhiAtV[g.s(i, j)];
    o.gV = o.gV + p.mtFacMom * mT;
  }
  // ---- Coriolis ----
  if (!p.useCDscheme) {
    double uCf, vCf;
    if (p.selectCoriScheme >= 2) {
      uCf = 0.5 * (g.fCori[g.s(i, j)] * 0.5 * (MV(i, j, k) + MV(i, j + 1, k)) +
                   g.fCori[g.s(i - 1, j)] * 0.5 * (MV(i - 1, j, k) + MV(i - 1, j + 1, k)));
      vCf = -0.5 * (g.fCori[g.s(i, j)] * 0.5 * (MU(i, j, k) + MU(i + 1, j, k)) +
                    g.fCori[g.s(i, j - 1)] * 0.5 * (MU(i, j - 1, k) + MU(i + 1, j - 1, k)));
    } else {
      uCf = 0.5 * (g.fCori[g.s(i, j)] + g.fCori[g.s(i - 1, j)]) * 0.25 *
            (MV(i, j, k) + MV(i, j + 1, k) + MV(i - 1, j, k) + MV(i - 1, j + 1, k));
      vCf = -0.5 * (g.fCori[g.s(i, j)] + g.fCori[g.s(i, j - 1)]) * 0.25 *
            (MU(i, j, k) + MU(i + 1, j, k) + MU(i, j - 1, k) + MU(i + 1, j - 1, k));
    }
    if (p.selectCoriScheme == 1 || p.selectCoriScheme == 3) {
      uCf = uCf * 4. / fmax(1., g.maskS[g.s3(i, j, k)] + g.maskS[g.s3(i, j + 1, k)] + g.maskS[g.s3(i - 1, j, k)] + g.maskS[g.s3(i - 1, j + 1, k)]);
      vCf = vCf * 4. / fmax(1., g.maskW[g.s3(i, j, k)] + g.maskW[g.s3(i + 1, j, k)] + g.maskW[g.s3(i, j - 1, k)] + g.maskW[g.s3(i + 1, j - 1, k)]);
    }
    o.gU = o.gU + p.cfFacMom * uCf;
    o.gV = o.gV + p.cfFacMom * vCf;
  }
  o.gU = o.gU * g.maskW[g.s3(i, j, k)];
  o.guDiss = o.guDiss * g.maskW[g.s3(i, j, k)];
  o.gV = o.gV * g.maskS[g.s3(i, j, k)];
  o.gvDiss = o.gvDiss * g.maskS[g.s3(i, j, k)];
  return o;
}

#undef MU
#undef MV
#undef MW

}  // namespace mg
