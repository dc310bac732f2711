// vecinv.cuh -- MOM_VECINV (pkg/mom_vecinv/mom_vecinv.F:10-1009) and the leaves it calls
// (pkg/mom_vecinv: mom_vi_coriolis.F, mom_vi_{u,v}_coriolis.F, mom_vi_{u,v}_grad_ke.F,
// mom_vi_{u,v}_vertshear.F, mom_vi_hdissip.F, mom_vi_del2uv.F; pkg/mom_common: mom_calc_hfacz.F,
// mom_calc_ke.F, mom_calc_relvort3.F, mom_calc_hdiv.F, mom_calc_absvort3.F, mom_{u,v}_rviscflux.F,
// mom_{u,v}_sidedrag.F, mom_{u,v}_botdrag_coeff.F; eesupp/src/fill_cs_corner_tr_rl.F).
// SURVEY.md section 8(f) rank 2.  Two stencil levels (vorticity / KE / divergence, then the tendencies)
// are staged through per-level scratch slabs, two more (del2u/del2v, then dStar/zStar) with
// biharmonic viscosity: the slabs of one level stay in L2 between the launches.
// Expression order follows the Fortran (-fmad=false): bit-identical to oracle/mom_oracle.c
// (og_mom_vecinv), which is pinned to verification/solid-body.cs-32x32x1/results/output.txt.
// Not on the B200 path (rejected by the entry point): momImplVertAdv, variable and strain-tension viscosity, Leith-QG,
// NH Coriolis / metric terms, GGL90-Langmuir, r* and sigma coordinates, OBCS, shelf ice.
#pragma once
#include "mom.cuh"

namespace mg {

struct ViPar {
  MomPar m;
  int useCoriolis, useAbsVorticity, selectVortScheme, useJamartMomAdv, upwindShear, selectKEscheme;
  int highOrderVorticity, upwindVorticity;   // MOM_VI_{U,V}_CORIOLIS_C4 instead of MOM_VI_{U,V}_CORIOLIS
  int harmonic;             // useHarmonicVisc: viscAh != 0
  int csCorners, myFace;    // facet corners this tile owns (1 SW, 2 SE, 4 NE, 8 NW), facet number
  int iMin, iMax, jMin, jMax;
};

bool make_vi_par(ViPar &p);   // host: from the run-time parameters; false + error for options not on the path

// per-level scratch slabs (PX*PY each)
struct ViScratch {
  double *hFacZ, *KE, *vort3, *hDiv, *del2u, *del2v, *dStar, *zStar, *omega3;
};

// ---- sources of the level's raw fields and of the metrics that are read at neighbouring points ----
// ViGlobalSrc: straight from global memory (through L1); uF, vF may be any slab pair (u, v of the level, or
// del2u, del2v).  ViPatchSrc: from the CTA's shared-memory patch (resident step).
struct ViGlobalSrc {
  const TileGrid &g; const double *uF, *vF; int k;
  __device__ double u(int i, int j) const { return uF[g.s(i, j)]; }
  __device__ double v(int i, int j) const { return vF[g.s(i, j)]; }
  __device__ double hW(int i, int j) const { return g.hFacW[g.s3(i, j, k)]; }
  __device__ double hS(int i, int j) const { return g.hFacS[g.s3(i, j, k)]; }
  __device__ double dxG(int i, int j) const { return g.dxG[g.s(i, j)]; }
  __device__ double dyG(int i, int j) const { return g.dyG[g.s(i, j)]; }
  __device__ double dxC(int i, int j) const { return g.dxC[g.s(i, j)]; }
  __device__ double dyC(int i, int j) const { return g.dyC[g.s(i, j)]; }
  __device__ double rAzI(int i, int j) const { return g.recip_rAz[g.s(i, j)]; }
};
constexpr int VT_X = 32, VT_Y = 8, VT_W = VT_X + 2, VT_H = VT_Y + 2, VT_N = VT_W * VT_H;
struct ViPatch { double u[VT_N], v[VT_N], hW[VT_N], hS[VT_N], dxG[VT_N], dyG[VT_N], dxC[VT_N], dyC[VT_N], rAzI[VT_N]; };
struct ViPatchSrc {
  const ViPatch &p; int i0, j0;      // patch origin: cell (li, lj) is column (i0 - 1 + li, j0 - 1 + lj)
  __device__ int e(int i, int j) const { return (j - j0 + 1) * VT_W + (i - i0 + 1); }
  __device__ double u(int i, int j) const { return p.u[e(i, j)]; }
  __device__ double v(int i, int j) const { return p.v[e(i, j)]; }
  __device__ double hW(int i, int j) const { return p.hW[e(i, j)]; }
  __device__ double hS(int i, int j) const { return p.hS[e(i, j)]; }
  __device__ double dxG(int i, int j) const { return p.dxG[e(i, j)]; }
  __device__ double dyG(int i, int j) const { return p.dyG[e(i, j)]; }
  __device__ double dxC(int i, int j) const { return p.dxC[e(i, j)]; }
  __device__ double dyC(int i, int j) const { return p.dyC[e(i, j)]; }
  __device__ double rAzI(int i, int j) const { return p.rAzI[e(i, j)]; }
};

// MOM_CALC_HFACZ (hZoption = 0): open-water fraction at the vorticity point
template <class F>
__device__ inline double vi_hfacz(const TileGrid &g, const F &f, int i, int j) {
  if (i < 2 - g.OLx || j < 2 - g.OLy) return 0.;
  double h = fmin(f.hW(i, j), f.hW(i, j - 1));
  h = fmin(f.hS(i, j), h);
  h = fmin(f.hS(i - 1, j), h);
  return h;
}

// MOM_CALC_KE (mom_calc_ke.F:55-125), zero outside 1-OL..sN+OL-1 like the zero-initialised slab
template <class F>
__device__ inline double vi_ke(const TileGrid &g, const F &f, int KEscheme, int k, int i, int j) {
  if (i > g.sNx + g.OLx - 1 || j > g.sNy + g.OLy - 1) return 0.;
  const double u0 = f.u(i, j), u1 = f.u(i + 1, j), v0 = f.v(i, j), v1 = f.v(i, j + 1);
  if (KEscheme == -1) return 0.125 * ((u0 + u1) * (u0 + u1) + (v0 + v1) * (v0 + v1));
  if (KEscheme == 0) return 0.25 * ((u0 * u0 + u1 * u1) + (v0 * v0 + v1 * v1));
  if (KEscheme == 1)
    return 0.25 * ((u0 * u0 * g.rAw[g.s(i, j)] + u1 * u1 * g.rAw[g.s(i + 1, j)]) +
                   (v0 * v0 * g.rAs[g.s(i, j)] + v1 * v1 * g.rAs[g.s(i, j + 1)])) * g.recip_rA[g.s(i, j)];
  if (KEscheme == 2)
    return 0.25 * ((u0 * u0 * f.hW(i, j) + u1 * u1 * f.hW(i + 1, j)) + (v0 * v0 * f.hS(i, j) + v1 * v1 * f.hS(i, j + 1))) *
           g.recip_hFacC[g.s3(i, j, k)];
  return 0.25 * ((u0 * u0 * f.hW(i, j) * g.rAw[g.s(i, j)] + u1 * u1 * f.hW(i + 1, j) * g.rAw[g.s(i + 1, j)]) +
                 (v0 * v0 * f.hS(i, j) * g.rAs[g.s(i, j)] + v1 * v1 * f.hS(i, j + 1) * g.rAs[g.s(i, j + 1)])) *
         g.recip_hFacC[g.s3(i, j, k)] * g.recip_rA[g.s(i, j)];
}

// MOM_CALC_RELVORT3 (mom_calc_relvort3.F:64-304), CALC_CS_CORNER_EXTENDED undefined
template <class F>
__device__ inline double vi_relvort3(const TileGrid &g, const F &f, int csCorners, int myFace, int i, int j) {
  if (i < 2 - g.OLx || j < 2 - g.OLy) return 0.;
  const double rz = f.rAzI(i, j);
  const double vdy = f.v(i, j) * f.dyC(i, j), vdym = f.v(i - 1, j) * f.dyC(i - 1, j);
  const double udx = f.u(i, j) * f.dxC(i, j), udxm = f.u(i, j - 1) * f.dxC(i, j - 1);
  if ((csCorners & 1) && i == 1 && j == 1) return +rz * ((vdy - udx) + udxm);
  if ((csCorners & 2) && i == g.sNx + 1 && j == 1) {
    if (myFace == 2) return +rz * ((-udx - vdym) + udxm);
    if (myFace == 4) return +rz * ((-vdym + udxm) - udx);
    return +rz * ((+udxm - udx) - vdym);
  }
  if ((csCorners & 8) && i == 1 && j == g.sNy + 1) {
    if (myFace == 1) return +rz * ((+udxm + vdy) - udx);
    if (myFace == 3) return +rz * ((-udx + udxm) + vdy);
    return +rz * ((+vdy - udx) + udxm);
  }
  if ((csCorners & 4) && i == g.sNx + 1 && j == g.sNy + 1) {
    if (myFace % 2 == 1) return +rz * ((-udx - vdym) + udxm);
    return +rz * ((+udxm - udx) - vdym);
  }
  return rz * ((vdy - vdym) - (udx - udxm));
}

// MOM_CALC_HDIV, hDivScheme = 2 (mom_calc_hdiv.F:56-72)
template <class F>
__device__ inline double vi_hdiv(const TileGrid &g, const F &f, int k, int i, int j) {
  if (i > g.sNx + g.OLx - 1 || j > g.sNy + g.OLy - 1) return 0.;
  return ((f.u(i + 1, j) * f.dyG(i + 1, j) * f.hW(i + 1, j) - f.u(i, j) * f.dyG(i, j) * f.hW(i, j)) +
          (f.v(i, j + 1) * f.dxG(i, j + 1) * f.hS(i, j + 1) - f.v(i, j) * f.dxG(i, j) * f.hS(i, j))) *
         g.recip_rA[g.s(i, j)] * g.recip_hFacC[g.s3(i, j, k)];
}

// FILL_CS_CORNER_TR_RL (fill_cs_corner_tr_rl.F:74-156) as a read map: the value the slab holds at (i,j)
// after the facet corners this tile owns were refilled for direction `dir` (1: x, 2: y); 0: untouched.
__device__ inline double vi_corner_read(const TileGrid &g, const double *f, int csCorners, int dir, int i, int j) {
  if (dir && csCorners) {
    const int sNx = g.sNx, sNy = g.sNy;
    if (i < 1 && j < 1 && (csCorners & 1)) {
      const int a = 1 - i, b = 1 - j;
      return dir == 1 ? f[g.s(1 - b, a)] : f[g.s(b, 1 - a)];
    }
    if (i > sNx && j < 1 && (csCorners & 2)) {
      const int a = i - sNx, b = 1 - j;
      return dir == 1 ? f[g.s(sNx + b, a)] : f[g.s(sNx + 1 - b, 1 - a)];
    }
    if (i < 1 && j > sNy && (csCorners & 8)) {
      const int a = 1 - i, b = j - sNy;
      return dir == 1 ? f[g.s(1 - b, sNy + 1 - a)] : f[g.s(b, sNy + a)];
    }
    if (i > sNx && j > sNy && (csCorners & 4)) {
      const int a = i - sNx, b = j - sNy;
      return dir == 1 ? f[g.s(sNx + b, sNy + 1 - a)] : f[g.s(sNx + 1 - b, sNy + a)];
    }
  }
  return f[g.s(i, j)];
}

// ---- stage 1: hFacZ, KE, vort3 (masked at hFacZ = 0), hDiv over the whole slab -----------------
static __global__ void __launch_bounds__(256) vi_stage1_kernel(TileGrid g, MomState st, ViPar p, int k, ViScratch w) {
  const int i = 1 - g.OLx + blockIdx.x * 32 + threadIdx.x;
  const int j = 1 - g.OLy + blockIdx.y * 8 + threadIdx.y;
  if (i > g.sNx + g.OLx || j > g.sNy + g.OLy) return;
  const size_t s = g.s(i, j);
  const ViGlobalSrc f{g, st.u + g.slab * (size_t)(k - 1), st.v + g.slab * (size_t)(k - 1), k};
  const double hz = vi_hfacz(g, f, i, j);
  w.hFacZ[s] = hz;
  w.KE[s] = vi_ke(g, f, p.selectKEscheme, k, i, j);
  double z = vi_relvort3(g, f, p.csCorners, p.myFace, i, j);
  if (hz == 0.) z = 0.;                                   // mom_vecinv.F:292-300
  w.vort3[s] = z;
  // MOM_CALC_ABSVORT3 (mom_calc_absvort3.F:34-46)
  w.omega3[s] = g.fCoriG[s] * (p.useCoriolis ? 1. : 0.) + z * (p.m.momAdvection ? 1. : 0.);
  w.hDiv[s] = p.m.momViscosity ? vi_hdiv(g, f, k, i, j) : 0.;
}

// ---- stage 2 (biharmonic): MOM_VI_DEL2UV (mom_vi_del2uv.F:78-124) --------------------------------
static __global__ void __launch_bounds__(256) vi_del2_kernel(TileGrid g, ViPar p, int k, ViScratch w) {
  const int i = 1 - g.OLx + blockIdx.x * 32 + threadIdx.x;
  const int j = 1 - g.OLy + blockIdx.y * 8 + threadIdx.y;
  if (i > g.sNx + g.OLx || j > g.sNy + g.OLy) return;
  const size_t s = g.s(i, j);
  double du = 0., dv = 0.;
  if (i >= 2 - g.OLx && i <= g.sNx + g.OLx - 1 && j >= 2 - g.OLy && j <= g.sNy + g.OLy - 1) {
    const size_t s3 = g.s3(i, j, k);
    const int c = p.csCorners;
    du = ((vi_corner_read(g, w.hDiv, c, 1, i, j) - vi_corner_read(g, w.hDiv, c, 1, i - 1, j)) * g.recip_dxC[s] -
          g.recip_hFacW[s3] * (w.hFacZ[g.s(i, j + 1)] * w.vort3[g.s(i, j + 1)] - w.hFacZ[s] * w.vort3[s]) * g.recip_dyG[s]) *
         g.maskW[s3];
    dv = ((vi_corner_read(g, w.hDiv, c, 2, i, j) - vi_corner_read(g, w.hDiv, c, 2, i, j - 1)) * g.recip_dyC[s] +
          g.recip_hFacS[s3] * (w.hFacZ[g.s(i + 1, j)] * w.vort3[g.s(i + 1, j)] - w.hFacZ[s] * w.vort3[s]) * g.recip_dxG[s]) *
         g.maskS[s3];
  }
  w.del2u[s] = du;
  w.del2v[s] = dv;
}

// ---- stage 3 (biharmonic): dStar, zStar (mom_vecinv.F:390-395) ----------------------------------
static __global__ void __launch_bounds__(256) vi_star_kernel(TileGrid g, ViPar p, int k, ViScratch w) {
  const int i = 1 - g.OLx + blockIdx.x * 32 + threadIdx.x;
  const int j = 1 - g.OLy + blockIdx.y * 8 + threadIdx.y;
  if (i > g.sNx + g.OLx || j > g.sNy + g.OLy) return;
  const size_t s = g.s(i, j);
  const ViGlobalSrc f{g, w.del2u, w.del2v, k};
  w.dStar[s] = vi_hdiv(g, f, k, i, j);
  w.zStar[s] = vi_relvort3(g, f, p.csCorners, p.myFace, i, j);
}

// ---- accessors of the level's intermediate fields ------------------------------------------------
// ViSlabAcc reads them from the scratch slabs the staging kernels wrote; ViFusedAcc re-evaluates them
// from u, v where they are read (no scratch traffic; everything but biharmonic viscosity).
struct ViSlabAcc {
  const TileGrid &g; ViScratch w; int csCorners, dirH;
  __device__ double hFacZ(int i, int j) const { return w.hFacZ[g.s(i, j)]; }
  __device__ double vort3(int i, int j) const { return w.vort3[g.s(i, j)]; }
  __device__ double omega3(int i, int j) const { return w.omega3[g.s(i, j)]; }
  __device__ double KE(int i, int j) const { return w.KE[g.s(i, j)]; }
  __device__ double hDiv(int i, int j) const { return vi_corner_read(g, w.hDiv, csCorners, dirH, i, j); }
  __device__ double del2u(int i, int j) const { return w.del2u[g.s(i, j)]; }
  __device__ double del2v(int i, int j) const { return w.del2v[g.s(i, j)]; }
  __device__ double dStar(int i, int j) const { return w.dStar[g.s(i, j)]; }
  __device__ double zStar(int i, int j) const { return w.zStar[g.s(i, j)]; }
};
template <class F>
struct ViFusedAcc {
  const TileGrid &g; const F &f; const ViPar &p; int k;
  __device__ double hFacZ(int i, int j) const { return vi_hfacz(g, f, i, j); }
  __device__ double vort3(int i, int j) const {
    if (vi_hfacz(g, f, i, j) == 0.) return 0.;
    return vi_relvort3(g, f, p.csCorners, p.myFace, i, j);
  }
  __device__ double omega3(int i, int j) const {
    return g.fCoriG[g.s(i, j)] * (p.useCoriolis ? 1. : 0.) + vort3(i, j) * (p.m.momAdvection ? 1. : 0.);
  }
  __device__ double KE(int i, int j) const { return vi_ke(g, f, p.selectKEscheme, k, i, j); }
  __device__ double hDiv(int i, int j) const { return vi_hdiv(g, f, k, i, j); }
  __device__ double del2u(int, int) const { return 0.; }
  __device__ double del2v(int, int) const { return 0.; }
  __device__ double dStar(int, int) const { return 0.; }
  __device__ double zStar(int, int) const { return 0.; }
};

// ViTileAcc: the CTA has evaluated them once per level for its 34 x 10 patch of columns (one-point rim)
// into shared memory; what the resident step uses.
struct ViTile { double hFacZ[VT_N], vort3[VT_N], omega3[VT_N], KE[VT_N], hDiv[VT_N]; };
struct ViTileAcc {
  const ViTile &t; int i0, j0;
  __device__ int e(int i, int j) const { return (j - j0 + 1) * VT_W + (i - i0 + 1); }
  __device__ double hFacZ(int i, int j) const { return t.hFacZ[e(i, j)]; }
  __device__ double vort3(int i, int j) const { return t.vort3[e(i, j)]; }
  __device__ double omega3(int i, int j) const { return t.omega3[e(i, j)]; }
  __device__ double KE(int i, int j) const { return t.KE[e(i, j)]; }
  __device__ double hDiv(int i, int j) const { return t.hDiv[e(i, j)]; }
  __device__ double del2u(int, int) const { return 0.; }
  __device__ double del2v(int, int) const { return 0.; }
  __device__ double dStar(int, int) const { return 0.; }
  __device__ double zStar(int, int) const { return 0.; }
};
// k-invariant metrics of the patch (once per CTA) and the raw fields of level k (all threads of a
// VT_X x VT_Y CTA; caller synchronises).  Columns past the halo'd slab are clamped; they are never consumed.
__device__ inline void vi_fill_patch_metrics(ViPatch &pt, const TileGrid &g, int i0, int j0, int tid) {
  for (int e = tid; e < VT_N; e += VT_X * VT_Y) {
    const size_t s = g.s(min(i0 - 1 + e % VT_W, g.sNx + g.OLx), min(j0 - 1 + e / VT_W, g.sNy + g.OLy));
    pt.dxG[e] = g.dxG[s]; pt.dyG[e] = g.dyG[s]; pt.dxC[e] = g.dxC[s]; pt.dyC[e] = g.dyC[s]; pt.rAzI[e] = g.recip_rAz[s];
  }
}
__device__ inline void vi_fill_patch(ViPatch &pt, const TileGrid &g, const MomState &st, int k, int i0, int j0, int tid) {
  for (int e = tid; e < VT_N; e += VT_X * VT_Y) {
    const size_t q = g.s3(min(i0 - 1 + e % VT_W, g.sNx + g.OLx), min(j0 - 1 + e / VT_W, g.sNy + g.OLy), k);
    pt.u[e] = st.u[q]; pt.v[e] = st.v[q]; pt.hW[e] = g.hFacW[q]; pt.hS[e] = g.hFacS[q];
  }
}
// derived fields of the patch from its raw fields; the rim cells whose stencil leaves the patch read global memory
__device__ inline void vi_fill_tile(ViTile &t, const ViPatch &pt, const TileGrid &g, const MomState &st, const ViPar &p, int k,
                                    int i0, int j0, int tid) {
  const ViPatchSrc fp{pt, i0, j0};
  const ViGlobalSrc fg{g, st.u + g.slab * (size_t)(k - 1), st.v + g.slab * (size_t)(k - 1), k};
  for (int e = tid; e < VT_N; e += VT_X * VT_Y) {
    const int li = e % VT_W, lj = e / VT_W, i = i0 - 1 + li, j = j0 - 1 + lj;
    double hz = 0., z = 0., om = 0., ke = 0., hd = 0.;
    if (i <= g.sNx + g.OLx && j <= g.sNy + g.OLy) {
      const bool lo = li >= 1 && lj >= 1, hi = li <= VT_W - 2 && lj <= VT_H - 2;
      hz = lo ? vi_hfacz(g, fp, i, j) : vi_hfacz(g, fg, i, j);
      if (hz != 0.) z = lo ? vi_relvort3(g, fp, p.csCorners, p.myFace, i, j) : vi_relvort3(g, fg, p.csCorners, p.myFace, i, j);
      om = g.fCoriG[g.s(i, j)] * (p.useCoriolis ? 1. : 0.) + z * (p.m.momAdvection ? 1. : 0.);
      ke = hi ? vi_ke(g, fp, p.selectKEscheme, k, i, j) : vi_ke(g, fg, p.selectKEscheme, k, i, j);
      if (p.m.momViscosity) hd = hi ? vi_hdiv(g, fp, k, i, j) : vi_hdiv(g, fg, k, i, j);
    }
    t.hFacZ[e] = hz; t.vort3[e] = z; t.omega3[e] = om; t.KE[e] = ke; t.hDiv[e] = hd;
  }
}

// MOM_VI_U_CORIOLIS / MOM_VI_V_CORIOLIS (mom_vi_{u,v}_coriolis.F:54-197), upwindVort3 = .FALSE.;
// absV selects omega3 (absolute) or vort3 (relative) as the advected vorticity
template <class F, class A>
__device__ inline double vi_u_coriolis(const TileGrid &g, const F &f, const ViPar &p, const A &a, bool absV, int k, int i, int j) {
  const double epsil = 1e-9, oneThird = 1. / 3.;
  const size_t s = g.s(i, j), s3 = g.s3(i, j, k);
  auto vdxh = [&](int ii, int jj) { return f.v(ii, jj) * f.dxG(ii, jj) * f.hS(ii, jj); };
  auto om = [&](int ii, int jj) { return absV ? a.omega3(ii, jj) : a.vort3(ii, jj); };
  auto rh = [&](int ii, int jj) { const double h = a.hFacZ(ii, jj); return h == 0. ? 0. : 1. / h; };
  auto rz = [&](int ii, int jj) { return rh(ii, jj) * om(ii, jj); };
  double r;
  const int sch = p.selectVortScheme;
  if (sch == 3 && i > g.sNx + g.OLx - 1) return 0.;
  if (sch == 0) {
    const double vBarXY = 0.25 * ((vdxh(i, j) + vdxh(i - 1, j)) + (vdxh(i, j + 1) + vdxh(i - 1, j + 1)));
    const double vort3u = 0.5 * (om(i, j) * rh(i, j) + om(i, j + 1) * rh(i, j + 1));
    r = +vort3u * vBarXY * g.recip_dxC[s] * g.maskW[s3];
  } else if (sch == 1) {
    const double h0 = a.hFacZ(i, j), h1 = a.hFacZ(i, j + 1);
    const double vBarXY = 0.5 * ((f.v(i, j) * f.dxG(i, j) * h0 + f.v(i - 1, j) * f.dxG(i - 1, j) * h0) +
                                 (f.v(i, j + 1) * f.dxG(i, j + 1) * h1 + f.v(i - 1, j + 1) * f.dxG(i - 1, j + 1) * h1)) /
                          fmax(epsil, h0 + h1);
    const double vort3u = 0.5 * (om(i, j) + om(i, j + 1));
    r = +vort3u * vBarXY * g.recip_dxC[s] * g.maskW[s3];
  } else if (sch == 2) {
    const double vBarXm = 0.5 * (vdxh(i, j) + vdxh(i - 1, j)), vBarXp = 0.5 * (vdxh(i, j + 1) + vdxh(i - 1, j + 1));
    const double vort3u = (vBarXm * rh(i, j) * om(i, j) + vBarXp * rh(i, j + 1) * om(i, j + 1)) * 0.5;
    r = +vort3u * g.recip_dxC[s] * g.maskW[s3];
  } else {
    const double vort3mj = (rz(i, j) + (rz(i, j + 1) + rz(i - 1, j))) * oneThird * vdxh(i - 1, j);
    const double vort3ij = (rz(i, j) + (rz(i, j + 1) + rz(i + 1, j))) * oneThird * vdxh(i, j);
    const double vort3mp = (rz(i, j + 1) + (rz(i, j) + rz(i - 1, j + 1))) * oneThird * vdxh(i - 1, j + 1);
    const double vort3ip = (rz(i, j + 1) + (rz(i, j) + rz(i + 1, j + 1))) * oneThird * vdxh(i, j + 1);
    r = +((vort3mj + vort3ij) + (vort3mp + vort3ip)) * 0.25 * g.recip_dxC[s] * g.maskW[s3];
  }
  if (p.useJamartMomAdv && i <= g.sNx + g.OLx - 1)
    r = r * 4. * f.hW(i, j) /
        fmax(epsil, (f.hS(i, j) + f.hS(i - 1, j)) + (f.hS(i, j + 1) + f.hS(i - 1, j + 1)));
  return r;
}
template <class F, class A>
__device__ inline double vi_v_coriolis(const TileGrid &g, const F &f, const ViPar &p, const A &a, bool absV, int k, int i, int j) {
  const double epsil = 1e-9, oneThird = 1. / 3.;
  const size_t s = g.s(i, j), s3 = g.s3(i, j, k);
  auto udyh = [&](int ii, int jj) { return f.u(ii, jj) * f.dyG(ii, jj) * f.hW(ii, jj); };
  auto om = [&](int ii, int jj) { return absV ? a.omega3(ii, jj) : a.vort3(ii, jj); };
  auto rh = [&](int ii, int jj) { const double h = a.hFacZ(ii, jj); return h == 0. ? 0. : 1. / h; };
  auto rz = [&](int ii, int jj) { return rh(ii, jj) * om(ii, jj); };
  double r;
  const int sch = p.selectVortScheme;
  if (sch == 3 && j > g.sNy + g.OLy - 1) return 0.;
  if (sch == 0) {
    const double uBarXY = 0.25 * ((udyh(i, j) + udyh(i, j - 1)) + (udyh(i + 1, j) + udyh(i + 1, j - 1)));
    const double vort3v = 0.5 * (om(i, j) * rh(i, j) + om(i + 1, j) * rh(i + 1, j));
    r = -vort3v * uBarXY * g.recip_dyC[s] * g.maskS[s3];
  } else if (sch == 1) {
    const double h0 = a.hFacZ(i, j), h1 = a.hFacZ(i + 1, j);
    const double uBarXY = 0.5 * ((f.u(i, j) * f.dyG(i, j) * h0 + f.u(i, j - 1) * f.dyG(i, j - 1) * h0) +
                                 (f.u(i + 1, j) * f.dyG(i + 1, j) * h1 + f.u(i + 1, j - 1) * f.dyG(i + 1, j - 1) * h1)) /
                          fmax(epsil, h0 + h1);
    const double vort3v = 0.5 * (om(i, j) + om(i + 1, j));
    r = -vort3v * uBarXY * g.recip_dyC[s] * g.maskS[s3];
  } else if (sch == 2) {
    const double uBarYm = 0.5 * (udyh(i, j) + udyh(i, j - 1)), uBarYp = 0.5 * (udyh(i + 1, j) + udyh(i + 1, j - 1));
    const double vort3v = (uBarYm * rh(i, j) * om(i, j) + uBarYp * rh(i + 1, j) * om(i + 1, j)) * 0.5;
    r = -vort3v * g.recip_dyC[s] * g.maskS[s3];
  } else {
    const double vort3im = (rz(i, j) + (rz(i + 1, j) + rz(i, j - 1))) * oneThird * udyh(i, j - 1);
    const double vort3ij = (rz(i, j) + (rz(i + 1, j) + rz(i, j + 1))) * oneThird * udyh(i, j);
    const double vort3pm = (rz(i + 1, j) + (rz(i, j) + rz(i + 1, j - 1))) * oneThird * udyh(i + 1, j - 1);
    const double vort3pj = (rz(i + 1, j) + (rz(i, j) + rz(i + 1, j + 1))) * oneThird * udyh(i + 1, j);
    r = -((vort3im + vort3ij) + (vort3pm + vort3pj)) * 0.25 * g.recip_dyC[s] * g.maskS[s3];
  }
  if (p.useJamartMomAdv && j <= g.sNy + g.OLy - 1)
    r = r * 4. * f.hS(i, j) /
        fmax(epsil, (f.hW(i, j) + f.hW(i, j - 1)) + (f.hW(i + 1, j) + f.hW(i + 1, j - 1)));
  return r;
}

// MOM_VI_U_CORIOLIS_C4 / MOM_VI_V_CORIOLIS_C4 (mom_vi_{u,v}_coriolis_c4.F:60-215): 4th-order (fourthVort3) or upwind
// interpolation of vort3r = r_hFacZ*omega3 along j (U) / i (V); selectVortScheme 0 and 2.  Defined on U: i = 1..sNx+1,
// j = 1..sNy, V: i = 1..sNx, j = 1..sNy+1 (the caller keeps the array's previous content elsewhere).
template <class F, class A>
__device__ inline double vi_coriolis_c4(const TileGrid &g, const F &f, const ViPar &p, const A &a, bool absV, int isV, int k, int i, int j) {
  const double oneSixth = 1. / 6., oneTwelve = 1. / 12.;
  const int sNx = g.sNx, sNy = g.sNy;
  auto base = [&](int ii, int jj) {
    const double h = a.hFacZ(ii, jj);
    return (h == 0. ? 0. : 1. / h) * (absV ? a.omega3(ii, jj) : a.vort3(ii, jj));
  };
  auto v3 = [&](int ii, int jj) {      // with the facet-corner averaging of the cube (highOrderVorticity only)
    if (p.csCorners && p.highOrderVorticity) {
      if (!isV) {
        if ((p.csCorners & 1) && ii == 1 && jj == 0) return (base(1, 0) + base(2, 1)) * 0.5;
        if ((p.csCorners & 2) && ii == sNx + 1 && jj == 0) return (base(sNx + 1, 0) + base(sNx, 1)) * 0.5;
        if ((p.csCorners & 8) && ii == 1 && jj == sNy + 2) return (base(1, sNy + 2) + base(2, sNy + 1)) * 0.5;
        if ((p.csCorners & 4) && ii == sNx + 1 && jj == sNy + 2) return (base(sNx + 1, sNy + 2) + base(sNx, sNy + 1)) * 0.5;
      } else {
        if ((p.csCorners & 1) && ii == 0 && jj == 1) return (base(0, 1) + base(1, 2)) * 0.5;
        if ((p.csCorners & 2) && ii == sNx + 2 && jj == 1) return (base(sNx + 2, 1) + base(sNx + 1, 2)) * 0.5;
        if ((p.csCorners & 8) && ii == 0 && jj == sNy + 1) return (base(0, sNy + 1) + base(1, sNy)) * 0.5;
        if ((p.csCorners & 4) && ii == sNx + 2 && jj == sNy + 1) return (base(sNx + 2, sNy + 1) + base(sNx + 1, sNy)) * 0.5;
      }
    }
    return base(ii, jj);
  };
  double bm, bp;
  if (!isV) {
    bm = f.v(i, j) * f.dxG(i, j) * f.hS(i, j) + f.v(i - 1, j) * f.dxG(i - 1, j) * f.hS(i - 1, j);
    bp = f.v(i, j + 1) * f.dxG(i, j + 1) * f.hS(i, j + 1) + f.v(i - 1, j + 1) * f.dxG(i - 1, j + 1) * f.hS(i - 1, j + 1);
  } else {
    bm = f.u(i, j) * f.dyG(i, j) * f.hW(i, j) + f.u(i, j - 1) * f.dyG(i, j - 1) * f.hW(i, j - 1);
    bp = f.u(i + 1, j) * f.dyG(i + 1, j) * f.hW(i + 1, j) + f.u(i + 1, j - 1) * f.dyG(i + 1, j - 1) * f.hW(i + 1, j - 1);
  }
  const int di = isV ? 1 : 0, dj = isV ? 0 : 1;
  const double z0 = v3(i, j), z1 = v3(i + di, j + dj), zm = v3(i - di, j - dj), z2 = v3(i + 2 * di, j + 2 * dj);
  const size_t s = g.s(i, j), s3 = g.s3(i, j, k);
  const double rd = isV ? g.recip_dyC[s] : g.recip_dxC[s], mk = isV ? g.maskS[s3] : g.maskW[s3];
  if (p.selectVortScheme == 0) {
    const double bXY = 0.25 * (bm + bp);
    double vort;
    if (p.upwindVorticity) vort = bXY > 0. ? z0 : z1;
    else {
      const double Rjp = z2 - z1, Rjm = z0 - zm;
      vort = 0.5 * ((z0 + z1) - oneTwelve * (Rjp - Rjm));
    }
    return isV ? -vort * bXY * rd * mk : vort * bXY * rd * mk;
  }
  const double bM = 0.5 * bm, bP = 0.5 * bp;
  double vort;
  if (p.upwindVorticity) vort = (bM + bP) > 0. ? bM * z0 : bP * z1;
  else {
    double Rjp = z2 - z1, Rjm = z0 - zm;
    const double Rj = z1 - z0;
    Rjp = z1 - oneSixth * (Rjp - Rj);
    Rjm = z0 - oneSixth * (Rj - Rjm);
    vort = 0.5 * (bM * Rjm + bP * Rjp);
  }
  return isV ? -vort * rd * mk : vort * rd * mk;
}

// MOM_VI_CORIOLIS (mom_vi_coriolis.F:48-190)
template <class F>
__device__ inline void vi_coriolis(const TileGrid &g, const F &f, int sch, int k, int i, int j, double &uCf, double &vCf) {
  const double epsil = 1e-9;
  const size_t s = g.s(i, j), s3 = g.s3(i, j, k);
  auto vdxh = [&](int ii, int jj) { return f.v(ii, jj) * f.dxG(ii, jj) * f.hS(ii, jj); };
  auto udyh = [&](int ii, int jj) { return f.u(ii, jj) * f.dyG(ii, jj) * f.hW(ii, jj); };
  {
    const double f0 = g.fCoriG[s], f1 = g.fCoriG[g.s(i, j + 1)];
    if (sch == 0) {
      const double vBarXY = 0.25 * ((f.v(i, j) * f.dxG(i, j) + f.v(i - 1, j) * f.dxG(i - 1, j)) +
                                    (f.v(i, j + 1) * f.dxG(i, j + 1) + f.v(i - 1, j + 1) * f.dxG(i - 1, j + 1)));
      uCf = +0.5 * (f0 + f1) * vBarXY * g.recip_dxC[s] * g.maskW[s3];
    } else if (sch == 1) {
      const double vBarXY = ((vdxh(i, j) + vdxh(i - 1, j)) + (vdxh(i, j + 1) + vdxh(i - 1, j + 1))) /
                            fmax(epsil, (f.hS(i, j) + f.hS(i - 1, j)) + (f.hS(i, j + 1) + f.hS(i - 1, j + 1)));
      uCf = +0.5 * (f0 + f1) * vBarXY * g.recip_dxC[s] * g.maskW[s3];
    } else if (sch == 2) {
      const double vBarXY = 0.25 * ((vdxh(i, j) + vdxh(i - 1, j)) + (vdxh(i, j + 1) + vdxh(i - 1, j + 1)));
      uCf = +0.5 * (f0 + f1) * vBarXY * g.recip_dxC[s] * g.recip_hFacW[s3];
    } else {
      const double vBarXm = 0.5 * (vdxh(i, j) + vdxh(i - 1, j)), vBarXp = 0.5 * (vdxh(i, j + 1) + vdxh(i - 1, j + 1));
      uCf = +0.5 * (vBarXm * f0 + vBarXp * f1) * g.recip_dxC[s] * g.recip_hFacW[s3];
    }
  }
  {
    const double f0 = g.fCoriG[s], f1 = g.fCoriG[g.s(i + 1, j)];
    if (sch == 0) {
      const double uBarXY = 0.25 * ((f.u(i, j) * f.dyG(i, j) + f.u(i, j - 1) * f.dyG(i, j - 1)) +
                                    (f.u(i + 1, j) * f.dyG(i + 1, j) + f.u(i + 1, j - 1) * f.dyG(i + 1, j - 1)));
      vCf = -0.5 * (f0 + f1) * uBarXY * g.recip_dyC[s] * g.maskS[s3];
    } else if (sch == 1) {
      const double uBarXY = ((udyh(i, j) + udyh(i, j - 1)) + (udyh(i + 1, j) + udyh(i + 1, j - 1))) /
                            fmax(epsil, (f.hW(i, j) + f.hW(i, j - 1)) + (f.hW(i + 1, j) + f.hW(i + 1, j - 1)));
      vCf = -0.5 * (f0 + f1) * uBarXY * g.recip_dyC[s] * g.maskS[s3];
    } else if (sch == 2) {
      const double uBarXY = 0.25 * ((udyh(i, j) + udyh(i, j - 1)) + (udyh(i + 1, j) + udyh(i + 1, j - 1)));
      vCf = -0.5 * (f0 + f1) * uBarXY * g.recip_dyC[s] * g.recip_hFacS[s3];
    } else {
      const double uBarYm = 0.5 * (udyh(i, j) + udyh(i, j - 1)), uBarYp = 0.5 * (udyh(i + 1, j) + udyh(i + 1, j - 1));
      vCf = -0.5 * (uBarYm * f0 + uBarYp * f1) * g.recip_dyC[s] * g.recip_hFacS[s3];
    }
  }
}

// MOM_VI_U_VERTSHEAR / MOM_VI_V_VERTSHEAR (mom_vi_{u,v}_vertshear.F:41-133)
__device__ inline double vi_vertshear(const TileGrid &g, const MomState &st, const ViPar &p, int isV, int k, int i, int j) {
  const int Nr = g.Nr;
  const bool rAdvAreaWeight = !(p.selectKEscheme == 1 || p.selectKEscheme == 3);
  const int Kp1 = min(k + 1, Nr), Km1 = max(k - 1, 1);
  const double mask_Kp1 = (k == Nr) ? 0. : 1., mask_Km1 = (k == 1) ? 0. : 1.;
  const int di = isV ? 0 : 1, dj = isV ? 1 : 0;
  const double *fld = isV ? st.v : st.u;
  const double rrA = (isV ? g.recip_rAs : g.recip_rAw)[g.s(i, j)];
  const double rh = (isV ? g.recip_hFacS : g.recip_hFacW)[g.s3(i, j, k)];
  double wBm, wBp;
  if (rAdvAreaWeight) {
    wBm = 0.5 * (st.w[g.s3(i, j, k)] * g.rA[g.s(i, j)] * g.maskC[g.s3(i, j, Km1)] +
                 st.w[g.s3(i - di, j - dj, k)] * g.rA[g.s(i - di, j - dj)] * g.maskC[g.s3(i - di, j - dj, Km1)]) * mask_Km1 * rrA;
    wBp = 0.5 * (st.w[g.s3(i, j, Kp1)] * g.rA[g.s(i, j)] + st.w[g.s3(i - di, j - dj, Kp1)] * g.rA[g.s(i - di, j - dj)]) * mask_Kp1 * rrA;
  } else {
    wBm = 0.5 * (st.w[g.s3(i, j, k)] * g.maskC[g.s3(i, j, Km1)] + st.w[g.s3(i - di, j - dj, k)] * g.maskC[g.s3(i - di, j - dj, Km1)]) * mask_Km1;
    wBp = 0.5 * (st.w[g.s3(i, j, Kp1)] + st.w[g.s3(i - di, j - dj, Kp1)]) * mask_Kp1;
  }
  const double fZm = (fld[g.s3(i, j, k)] - mask_Km1 * fld[g.s3(i, j, Km1)]) * p.m.rkSign;
  const double fZp = (mask_Kp1 * fld[g.s3(i, j, Kp1)] - fld[g.s3(i, j, k)]) * p.m.rkSign;
  if (p.upwindShear)
    return -0.5 * ((wBp * fZp + wBm * fZm) + (fabs(wBp) * fZp - fabs(wBm) * fZm)) * rh * g.recip_drF[k - 1];
  return -0.5 * (wBp * fZp + wBm * fZm) * rh * g.recip_drF[k - 1];
}

// MOM_{U,V}_SIDEDRAG with the vector-invariant del2u / del2v (sideDragFactor > 0, constant viscosity)
template <class F, class A>
__device__ inline double vi_sidedrag(const TileGrid &g, const F &f, const ViPar &p, const A &a, int isV, int k, int i, int j) {
  const size_t s = g.s(i, j), s3 = g.s3(i, j, k);
  if (!isV) {
    const double hS = f.hW(i, j) - a.hFacZ(i, j), hN = f.hW(i, j) - a.hFacZ(i, j + 1);
    const double d2 = p.m.useBiharmonicVisc ? a.del2u(i, j) : 0.;
    const double t = p.m.viscAhZ * f.u(i, j) - p.m.viscA4Z * d2;
    return -g.recip_hFacW[s3] * g.recip_drF[k - 1] * g.recip_rAw[s] *
           (hS * g.dxV[s] * g.recip_dyU[s] * t + hN * g.dxV[g.s(i, j + 1)] * g.recip_dyU[g.s(i, j + 1)] * t) * g.drF[k - 1] *
           p.m.sideDragFactor;
  }
  const double cf = g.cosFacV[j + g.OLy - 1];
  const double hW = f.hS(i, j) - a.hFacZ(i, j), hE = f.hS(i, j) - a.hFacZ(i + 1, j);
  const double d2 = p.m.useBiharmonicVisc ? a.del2v(i, j) : 0.;
  const double t = p.m.viscAhZ * f.v(i, j) * cf - p.m.viscA4Z * d2 * cf;
  return -g.recip_hFacS[s3] * g.recip_drF[k - 1] * g.recip_rAs[s] *
         (hW * g.dyU[s] * g.recip_dxV[s] * t + hE * g.dyU[g.s(i + 1, j)] * g.recip_dxV[g.s(i + 1, j)] * t) * g.drF[k - 1] *
         p.m.sideDragFactor;
}

struct ViOut { double gU, gV, guDiss, gvDiss, fVerUkp, fVerVkp; };

// Tendencies of one point (mom_vecinv.F:308-927).  inRange: the point is inside iMin..iMax x jMin..jMax
// (outside, only MOM_VI_HDISSIP's own range is written); fVer?km: viscous vertical flux at the upper
// interface; o.fVer?kp is valid only when inRange, momViscosity and not implicitViscosity.
template <class F, class A>
__device__ inline ViOut vi_cell(const TileGrid &g, const MomState &st, const F &f, const ViPar &p, const A &a, int k, int i, int j,
                                bool inRange, double fVerUkm, double fVerVkm) {
  ViOut o;
  const size_t s = g.s(i, j), s3 = g.s3(i, j, k);
  const MomPar &m = p.m;
  double uD = 0., vD = 0.;
  o.fVerUkp = 0.; o.fVerVkp = 0.;
  if (m.momViscosity) {
    // MOM_VI_HDISSIP (mom_vi_hdissip.F:60-271) on its own range, constant coefficients
    if (i >= 2 - g.OLx && i <= g.sNx + g.OLx - 1 && j >= 2 - g.OLy && j <= g.sNy + g.OLy - 1) {
      const double cU = g.cosFacU[j + g.OLy - 1], cV = g.cosFacV[j + g.OLy - 1];
      if (p.harmonic) {
        const double Dim = a.hDiv(i, j - 1), Dij = a.hDiv(i, j), Dmj = a.hDiv(i - 1, j);
        const double Zip = a.hFacZ(i, j + 1) * a.vort3(i, j + 1), Zij = a.hFacZ(i, j) * a.vort3(i, j),
                     Zpj = a.hFacZ(i + 1, j) * a.vort3(i + 1, j);
        const double uD2 = m.viscAhD * cU * (Dij - Dmj) * g.recip_dxC[s] - m.viscAhZ * g.recip_hFacW[s3] * (Zip - Zij) * g.recip_dyG[s];
        const double vD2 = m.viscAhZ * g.recip_hFacS[s3] * cV * (Zpj - Zij) * g.recip_dxG[s] + m.viscAhD * (Dij - Dim) * g.recip_dyC[s];
        uD = uD2 * g.maskW[s3];
        vD = vD2 * g.maskS[s3];
      }
      if (m.useBiharmonicVisc) {
        const double Dim = a.dStar(i, j - 1), Dij = a.dStar(i, j), Dmj = a.dStar(i - 1, j);
        const double Zip = a.hFacZ(i, j + 1) * a.zStar(i, j + 1), Zij = a.hFacZ(i, j) * a.zStar(i, j),
                     Zpj = a.hFacZ(i + 1, j) * a.zStar(i + 1, j);
        double uD4 = m.viscA4D * cU * (Dij - Dmj) * g.recip_dxC[s] - m.viscA4Z * g.recip_hFacW[s3] * (Zip - Zij) * g.recip_dyG[s];
        double vD4 = m.viscA4Z * g.recip_hFacS[s3] * cV * (Zpj - Zij) * g.recip_dxG[s] + m.viscA4D * (Dij - Dim) * g.recip_dyC[s];
        uD4 = -uD4 * g.maskW[s3];
        vD4 = -vD4 * g.maskS[s3];
        uD = uD + uD4;
        vD = vD + vD4;
      }
    }
    if (inRange) {
      const double rhW = g.recip_hFacW[s3], rhS = g.recip_hFacS[s3], rdrF = g.recip_drF[k - 1];
      const bool quad0 = m.bottomDragTerms && m.selectBotDragQuadr == 0;
      if (!m.implicitViscosity) {
        o.fVerUkp = m.vfFacMom * 1. * mom_u_rvisc(g, st, m, k + 1, i, j);
        uD = uD - rhW * rdrF * g.recip_rAw[s] * (o.fVerUkp - fVerUkm) * m.rkSign;
      }
      if (m.no_slip_sides) uD = uD + vi_sidedrag(g, f, p, a, 0, k, i, j);
      if (m.bottomDragTerms)
        uD = uD + (-mom_botdrag(g, st, m, k, 0, i, j, true, quad0 ? a.KE(i, j) + a.KE(i - 1, j) : 0.) * f.u(i, j) * rhW * rdrF);
      if (!m.implicitViscosity) {
        o.fVerVkp = m.vfFacMom * 1. * mom_v_rvisc(g, st, m, k + 1, i, j);
        vD = vD - rhS * rdrF * g.recip_rAs[s] * (o.fVerVkp - fVerVkm) * m.rkSign;
      }
      if (m.no_slip_sides) vD = vD + vi_sidedrag(g, f, p, a, 1, k, i, j);
      if (m.bottomDragTerms)
        vD = vD + (-mom_botdrag(g, st, m, k, 1, i, j, true, quad0 ? a.KE(i, j) + a.KE(i, j - 1) : 0.) * f.v(i, j) * rhS * rdrF);
    }
  }
  o.guDiss = uD;
  o.gvDiss = vD;
  o.gU = 0.; o.gV = 0.;
  if (!inRange) return o;
  // ---- Coriolis and advection (mom_vecinv.F:672-884)
  double tU = 0., tV = 0.;
  if (p.useCoriolis && !(m.useCDscheme || (p.useAbsVorticity && m.momAdvection))) {
    if (p.useAbsVorticity) {
      tU = vi_u_coriolis(g, f, p, a, true, k, i, j);
      tV = vi_v_coriolis(g, f, p, a, true, k, i, j);
    } else {
      vi_coriolis(g, f, m.selectCoriScheme, k, i, j, tU, tV);
    }
  }
  if (m.momAdvection) {
    if (p.highOrderVorticity || p.upwindVorticity) {
      // mom_vecinv.F:746-757, :772-783: outside its own range the C4 routine leaves uCf / vCf as the Coriolis call left them
      const bool inU = i >= 1 && i <= g.sNx + 1 && j >= 1 && j <= g.sNy, inV = i >= 1 && i <= g.sNx && j >= 1 && j <= g.sNy + 1;
      const double staleU = tU, staleV = tV;
      tU = tU + (inU ? vi_coriolis_c4(g, f, p, a, p.useAbsVorticity != 0, 0, k, i, j) : staleU);
      tV = tV + (inV ? vi_coriolis_c4(g, f, p, a, p.useAbsVorticity != 0, 1, k, i, j) : staleV);
    } else {
    tU = tU + vi_u_coriolis(g, f, p, a, p.useAbsVorticity != 0, k, i, j);
    tV = tV + vi_v_coriolis(g, f, p, a, p.useAbsVorticity != 0, k, i, j);
    }
    tU = tU + vi_vertshear(g, st, p, 0, k, i, j);
    tV = tV + vi_vertshear(g, st, p, 1, k, i, j);
    const double ke = a.KE(i, j);
    tU = tU + (-g.recip_dxC[s] * (ke - a.KE(i - 1, j)) * g.maskW[s3]);
    tV = tV + (-g.recip_dyC[s] * (ke - a.KE(i, j - 1)) * g.maskS[s3]);
  }
  o.gU = tU * g.maskW[s3];
  o.gV = tV * g.maskS[s3];
  return o;
}

// ---- final stage of the per-level entry point ----------------------------------------------------
template <bool FUSED>
__global__ void __launch_bounds__(256) vi_tend_kernel(TileGrid g, MomState st, ViPar p, int k, ViScratch w,
                                                      const double *fVerUkm, const double *fVerVkm, double *fVerUkp,
                                                      double *fVerVkp, double *guDiss, double *gvDiss, double *gU, double *gV) {
  const int i = 1 - g.OLx + blockIdx.x * 32 + threadIdx.x;
  const int j = 1 - g.OLy + blockIdx.y * 8 + threadIdx.y;
  if (i > g.sNx + g.OLx || j > g.sNy + g.OLy) return;
  const size_t s = g.s(i, j), s3 = g.s3(i, j, k);
  const bool inRange = i >= p.iMin && i <= p.iMax && j >= p.jMin && j <= p.jMax;
  ViOut o;
  if (FUSED) {
    // outside MOM_VI_HDISSIP's range nothing is evaluated (the accessors would read past the slab)
    const bool inner = i >= 2 - g.OLx && i <= g.sNx + g.OLx - 1 && j >= 2 - g.OLy && j <= g.sNy + g.OLy - 1;
    const ViGlobalSrc f{g, st.u + g.slab * (size_t)(k - 1), st.v + g.slab * (size_t)(k - 1), k};
    if (inner) o = vi_cell(g, st, f, p, ViFusedAcc<ViGlobalSrc>{g, f, p, k}, k, i, j, inRange, fVerUkm[s], fVerVkm[s]);
    else { o.guDiss = 0.; o.gvDiss = 0.; }
  } else {
    // hDiv keeps MOM_VI_DEL2UV's last facet-corner fill (direction 2) when that routine ran
    const int dirH = (p.m.useBiharmonicVisc && p.csCorners) ? 2 : 0;
    const ViGlobalSrc f{g, st.u + g.slab * (size_t)(k - 1), st.v + g.slab * (size_t)(k - 1), k};
    o = vi_cell(g, st, f, p, ViSlabAcc{g, w, p.csCorners, dirH}, k, i, j, inRange, fVerUkm[s], fVerVkm[s]);
  }
  guDiss[s] = o.guDiss;
  gvDiss[s] = o.gvDiss;
  if (!inRange) return;
  if (p.m.momViscosity && !p.m.implicitViscosity) { fVerUkp[s] = o.fVerUkp; fVerVkp[s] = o.fVerVkp; }
  gU[s3] = o.gU;
  gV[s3] = o.gV;
}


}  // namespace mg
