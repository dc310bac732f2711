// colgeom.cu -- column geometry: the nine time-invariant 3-D arrays hFacC/W/S, recip_hFacC/W/S, maskC/W/S of
// INI_MASKS_ETC (model/src/ini_masks_etc.F:100-478) compressed to three numbers per column and point type.
//
// With z levels, a linear free surface and the surface at the top of level 1, a column of open-water fractions is
// 1 down to its deepest wet level kLow, a partial value hLow there and 0 below; the mask is (hFac != 0) and the
// reciprocal 1/hFac where wet.  The 3-D kernels of the resident step read these nine arrays at every cell and every
// step (74 B per cell in MOM_FLUXFORM, 32 B in GAD_CALC_RHS, 40 B in the correction step, 16 B in CALC_DIV_GHAT:
// a third of the non-solver traffic).  When -- and only when -- ALL nine mirrors have exactly that form, checked
// element by element on the device, the step kernels take (kLow, hLow, recip at kLow) per column instead and rebuild
// the level values by compare + select: the same numbers, bit for bit, without the 3-D reads.  Anything else
// (partial cells at the top: shelf ice, p coordinates; non-linear free surface / r* with time-dependent hFac; hand-made
// masks) fails the check and the general kernels run.  The check is repeated whenever a geometry mirror is written
// (set_field / fill_field) or its address is handed out (field_ptr).  MITGCM_B200_NO_COLGEOM=1 switches it off.
#include <cstdlib>
#include "context.h"
#include "gad.cuh"

namespace mg {

struct ColGeomArgs {
  int PX, PY, Nr, nTiles;
  size_t slab;
  const double *h[3], *r[3], *m[3];
  int *k[3];
  double *hl[3], *rl[3];
  int *flag;
};

__global__ void col_geom_kernel(ColGeomArgs a) {
  const int ii = blockIdx.x * blockDim.x + threadIdx.x, jj = blockIdx.y * blockDim.y + threadIdx.y, tile = blockIdx.z;
  if (ii >= a.PX || jj >= a.PY) return;
  const size_t s2 = (size_t)ii + (size_t)a.PX * jj + a.slab * tile;
  const size_t s3 = (size_t)ii + (size_t)a.PX * jj + a.slab * (size_t)a.Nr * tile;
  bool ok = true;
  for (int t = 0; t < 3; t++) {
    int kLow = 0;
    for (int k = 1; k <= a.Nr; k++)
      if (a.h[t][s3 + a.slab * (k - 1)] != 0.) {
        if (kLow != k - 1) ok = false;      // a dry level above a wet one
        kLow = k;
      }
    double hLow = 0., rLow = 0.;
    for (int k = 1; k <= a.Nr; k++) {
      const size_t q = s3 + a.slab * (k - 1);
      const double h = a.h[t][q], r = a.r[t][q], m = a.m[t][q];
      if (k < kLow) ok = ok && h == 1. && r == 1. && m == 1.;
      else if (k == kLow) { hLow = h; rLow = r; ok = ok && m == 1. && h > 0.; }
      else ok = ok && h == 0. && r == 0. && m == 0.;
    }
    a.k[t][s2] = kLow; a.hl[t][s2] = hLow; a.rl[t][s2] = rLow;
  }
  if (!ok) atomicExch(a.flag, 1);
}

void col_geom_touch(int id) {
  switch (id) {
    case MG_HFACC: case MG_HFACW: case MG_HFACS: case MG_RECIP_HFACC: case MG_RECIP_HFACW: case MG_RECIP_HFACS:
    case MG_MASKC: case MG_MASKW: case MG_MASKS:
      if (ctx().cgState != 0) ctx().cgState = 0;
      break;
    default: break;
  }
}

void col_geom_free() {
  Ctx &c = ctx();
  for (int t = 0; t < 3; t++) {
    if (c.cgK[t]) cudaFree(c.cgK[t]);
    if (c.cgH[t]) cudaFree(c.cgH[t]);
    if (c.cgR[t]) cudaFree(c.cgR[t]);
    c.cgK[t] = nullptr; c.cgH[t] = c.cgR[t] = nullptr;
  }
  if (c.cgFlag) cudaFree(c.cgFlag);
  c.cgFlag = nullptr;
  c.cgState = 0; c.cgFails = 0;
}

bool col_geom_ready() {
  Ctx &c = ctx();
  if (getenv("MITGCM_B200_NO_COLGEOM")) return false;
  if (c.cgState == 1) return true;
  if (c.cgState == -1 || c.cgFails >= 3) return false;      // time-dependent geometry: stop checking every step
  const Geom &g = c.g;
  ColGeomArgs a;
  a.PX = g.PX; a.PY = g.PY; a.Nr = g.Nr; a.nTiles = g.nTiles; a.slab = g.slab;
  const int ids[3][3] = {{MG_HFACC, MG_RECIP_HFACC, MG_MASKC}, {MG_HFACW, MG_RECIP_HFACW, MG_MASKW}, {MG_HFACS, MG_RECIP_HFACS, MG_MASKS}};
  for (int t = 0; t < 3; t++) {
    a.h[t] = field(ids[t][0], false); a.r[t] = field(ids[t][1], false); a.m[t] = field(ids[t][2], false);
    if (!a.h[t] || !a.r[t] || !a.m[t]) return false;
  }
  for (int t = 0; t < 3; t++) {
    if (!c.cgK[t]) MG_CUDA(cudaMalloc(&c.cgK[t], g.n2 * sizeof(int)));
    if (!c.cgH[t]) MG_CUDA(cudaMalloc(&c.cgH[t], g.n2 * sizeof(double)));
    if (!c.cgR[t]) MG_CUDA(cudaMalloc(&c.cgR[t], g.n2 * sizeof(double)));
    a.k[t] = c.cgK[t]; a.hl[t] = c.cgH[t]; a.rl[t] = c.cgR[t];
  }
  if (!c.cgFlag) MG_CUDA(cudaMalloc(&c.cgFlag, sizeof(int)));
  MG_CUDA(cudaMemsetAsync(c.cgFlag, 0, sizeof(int), c.stream));
  a.flag = c.cgFlag;
  c.launches++;
  col_geom_kernel<<<dim3((g.PX + 31) / 32, (g.PY + 7) / 8, g.nTiles), dim3(32, 8), 0, c.stream>>>(a);
  MG_CUDA(cudaGetLastError());
  int flag = 1;
  MG_CUDA(cudaMemcpyAsync(&flag, c.cgFlag, sizeof(int), cudaMemcpyDeviceToHost, c.stream));
  MG_CUDA(cudaStreamSynchronize(c.stream));
  if (flag) { c.cgFails++; c.cgState = c.cgFails >= 3 ? -1 : 0; return false; }
  c.cgState = 1;
  return true;
}

bool attach_col_geom(int bi, int bj, TileGrid &t) {
  Ctx &c = ctx();
  if (c.cgState != 1 || getenv("MITGCM_B200_NO_COLGEOM")) return false;
  const size_t o = c.g.slab * ((size_t)(bi - 1) + (size_t)c.g.nSx * (size_t)(bj - 1));
  t.kLowC = c.cgK[0] + o; t.kLowW = c.cgK[1] + o; t.kLowS = c.cgK[2] + o;
  t.hLowC = c.cgH[0] + o; t.hLowW = c.cgH[1] + o; t.hLowS = c.cgH[2] + o;
  t.rhLowC = c.cgR[0] + o; t.rhLowW = c.cgR[1] + o; t.rhLowS = c.cgR[2] + o;
  return true;
}

}  // namespace mg

extern "C" int mitgcm_b200_col_geom_state_(void) { return mg::ctx().ready ? mg::ctx().cgState : 0; }
