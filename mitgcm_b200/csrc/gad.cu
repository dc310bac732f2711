// gad.cu -- GAD_CALC_RHS drop-in (pkg/generic_advdiff/gad_calc_rhs.F:10-795): one launch per
// (tile, level, tracer) call with the reference argument list.  All ~20 slab sweeps of the
// reference collapse into one pass over the halo'd slab; see gad.cuh for the point-wise form.
// Algorithmic traffic of this per-level form: 11 slab reads + tracer levels k-1,k + gTracer(k),
// fVerT(kDown) in; fZon, fMer, fVerT(kUp), gTracer(k) out  (DESIGN.md).
#include "gad.cuh"

namespace mg {

// Level-k inputs exactly as the reference passes them (gad_calc_rhs.F:89-123).
struct SlabAcc {
  int OLx, OLy, PX;
  size_t slab;
  const double *xA_, *yA_, *maskUp_, *uFld_, *vFld_, *wFld_, *uTrans_, *vTrans_, *rTrans_, *rTransKp1_, *KappaR_;
  const double *TracerN, *TracAB;   // (slab, Nr) of this tile; TracAB == TracerN unless applyAB_onTracer
  __device__ __forceinline__ size_t s(int i, int j) const {
    return (size_t)(i + OLx - 1) + (size_t)PX * (size_t)(j + OLy - 1);
  }
  __device__ __forceinline__ double T(int i, int j, int k) const { return TracerN[s(i, j) + slab * (size_t)(k - 1)]; }
  __device__ __forceinline__ double TA(int i, int j, int k) const { return TracAB[s(i, j) + slab * (size_t)(k - 1)]; }
  __device__ __forceinline__ double xA(int i, int j) const { return xA_[s(i, j)]; }
  __device__ __forceinline__ double yA(int i, int j) const { return yA_[s(i, j)]; }
  __device__ __forceinline__ double maskUp(int i, int j) const { return maskUp_[s(i, j)]; }
  __device__ __forceinline__ double uFld(int i, int j) const { return uFld_[s(i, j)]; }
  __device__ __forceinline__ double vFld(int i, int j) const { return vFld_[s(i, j)]; }
  __device__ __forceinline__ double wFld(int i, int j) const { return wFld_[s(i, j)]; }
  __device__ __forceinline__ double uTrans(int i, int j) const { return uTrans_[s(i, j)]; }
  __device__ __forceinline__ double vTrans(int i, int j) const { return vTrans_[s(i, j)]; }
  __device__ __forceinline__ double rTrans(int i, int j) const { return rTrans_[s(i, j)]; }
  __device__ __forceinline__ double rTransKp1(int i, int j) const { return rTransKp1_[s(i, j)]; }
  __device__ __forceinline__ double KappaR(int i, int j) const { return KappaR_[s(i, j)]; }
};

__global__ void __launch_bounds__(256) gad_level_kernel(TileGrid g, SlabAcc a, GadPar p, double *fZon, double *fMer,
                                                        double *fVerUp, const double *fVerDn, double *gTracer) {
  const int i = 1 - g.OLx + blockIdx.x * 32 + threadIdx.x;
  const int j = 1 - g.OLy + blockIdx.y * 8 + threadIdx.y;
  if (i > g.sNx + g.OLx || j > g.sNy + g.OLy) return;
  const size_t s = g.s(i, j);
  const double fz0 = gad_fzon(g, a, p, i, j);
  const double fm0 = gad_fmer(g, a, p, i, j);
  const double fv = gad_fver(g, a, p, i, j);
  fZon[s] = fz0;
  fMer[s] = fm0;
  fVerUp[s] = fv;
  if (i <= g.sNx + g.OLx - 1 && j <= g.sNy + g.OLy - 1) {
    const double fz1 = gad_fzon(g, a, p, i + 1, j);
    const double fm1 = gad_fmer(g, a, p, i, j + 1);
    const size_t s3 = g.s3(i, j, p.k);
    gTracer[s3] = gad_tendency(g, a, p, i, j, gTracer[s3], fz0, fz1, fm0, fm1, fv, fVerDn[s]);
  }
}

bool make_tile_grid(int bi, int bj, TileGrid &t) {
  Ctx &c = ctx();
  const Geom &g = c.g;
  if (bi < 1 || bi > g.nSx || bj < 1 || bj > g.nSy) return fail(40, "tile index out of range");
  const size_t tile = (size_t)(bi - 1) + (size_t)g.nSx * (size_t)(bj - 1);
  t.sNx = g.sNx; t.sNy = g.sNy; t.OLx = g.OLx; t.OLy = g.OLy; t.Nr = g.Nr; t.PX = g.PX; t.PY = g.PY; t.slab = g.slab;
  auto f2 = [&](int id) -> const double * { double *p = field(id); return p ? p + g.slab * tile : nullptr; };
  auto f3 = [&](int id) -> const double * { double *p = field(id); return p ? p + g.slab * g.Nr * tile : nullptr; };
  t.dxC = f2(MG_DXC); t.dyC = f2(MG_DYC); t.dxG = f2(MG_DXG); t.dyG = f2(MG_DYG); t.dxF = f2(MG_DXF); t.dyF = f2(MG_DYF);
  t.dxV = f2(MG_DXV); t.dyU = f2(MG_DYU); t.rA = f2(MG_RA); t.rAw = f2(MG_RAW); t.rAs = f2(MG_RAS);
  t.recip_dxC = f2(MG_RECIP_DXC); t.recip_dyC = f2(MG_RECIP_DYC); t.recip_dxF = f2(MG_RECIP_DXF);
  t.recip_dyF = f2(MG_RECIP_DYF); t.recip_dxV = f2(MG_RECIP_DXV); t.recip_dyU = f2(MG_RECIP_DYU);
  t.recip_rA = f2(MG_RECIP_RA); t.recip_rAw = f2(MG_RECIP_RAW); t.recip_rAs = f2(MG_RECIP_RAS);
  t.recip_dxG = f2(MG_RECIP_DXG); t.recip_dyG = f2(MG_RECIP_DYG); t.recip_rAz = f2(MG_RECIP_RAZ); t.fCoriG = f2(MG_FCORIG);
  t.fCori = f2(MG_FCORI); t.tanPhiAtU = f2(MG_TANPHIATU); t.tanPhiAtV = f2(MG_TANPHIATV);
  double *cu = field(MG_COSFACU), *cv = field(MG_COSFACV);
  t.cosFacU = cu ? cu + (size_t)g.PY * tile : nullptr;
  t.cosFacV = cv ? cv + (size_t)g.PY * tile : nullptr;
  t.drF = field(MG_DRF); t.drC = field(MG_DRC); t.recip_drF = field(MG_RECIP_DRF); t.recip_drC = field(MG_RECIP_DRC);
  t.hFacC = f3(MG_HFACC); t.hFacW = f3(MG_HFACW); t.hFacS = f3(MG_HFACS);
  t.recip_hFacC = f3(MG_RECIP_HFACC); t.recip_hFacW = f3(MG_RECIP_HFACW); t.recip_hFacS = f3(MG_RECIP_HFACS);
  t.maskC = f3(MG_MASKC); t.maskW = f3(MG_MASKW); t.maskS = f3(MG_MASKS);
  t.kLowC = t.kLowW = t.kLowS = nullptr;
  t.hLowC = t.hLowW = t.hLowS = t.rhLowC = t.rhLowW = t.rhLowS = nullptr;
  return t.maskS != nullptr && t.recip_drC != nullptr;
}

static bool supported_scheme(int s) {
  return s == ADV_UPWIND_1RST || s == ADV_CENTERED_2ND || s == ADV_UPWIND_3RD || s == ADV_CENTERED_4TH ||
         s == ADV_DST2 || s == ADV_FLUX_LIMIT || s == ADV_DST3 || s == ADV_DST3_FLUX_LIMIT || s == ADV_OS7MP;
}

}  // namespace mg

using namespace mg;

extern "C" void gad_calc_rhs_b200_(
    const int *bi, const int *bj, const int *iMin, const int *iMax, const int *jMin, const int *jMax, const int *k,
    const int *kM1, const int *kUp, const int *kDown, const double *xA, const double *yA, const double *maskUp,
    const double *uFld, const double *vFld, const double *wFld, const double *uTrans, const double *vTrans,
    const double *rTrans, const double *rTransKp1, const double *diffKh, const double *diffK4, const double *KappaR,
    const double *diffKr4, const double *TracerN, const double *TracAB, const double *deltaTLev, const int *trIdentity,
    const int *advectionSchArg, const int *vertAdvecSchArg, const int *calcAdvection, const int *implicitAdvection,
    const int *applyAB_onTracer, const int *trUseDiffKr4, const int *trUseGMRedi, const int *trUseKPP,
    const int *trUseSmolHack, double *fZon, double *fMer, double *fVerT, double *gTracer, const double *myTime,
    const int *myIter, const int *myThid) {
  (void)iMin; (void)iMax; (void)jMin; (void)jMax; (void)kM1; (void)trIdentity; (void)myTime; (void)myIter; (void)myThid;
  Ctx &c = ctx();
  c.lastError = 0;
  if (!c.ready) { fail(30, "mitgcm_b200_init_ not called"); return; }
  if (*trUseGMRedi || *trUseKPP || *trUseSmolHack) {
    fail(41, "gad_calc_rhs_b200_: GM/Redi, KPP and the Smolarkiewicz hack are not on the B200 path");
    return;
  }
  if (*calcAdvection && (!supported_scheme(*advectionSchArg) || (!*implicitAdvection && !supported_scheme(*vertAdvecSchArg)))) {
    fail(42, "gad_calc_rhs_b200_: unsupported advection scheme");
    return;
  }
  const Geom &g = c.g;
  if (*calcAdvection && *advectionSchArg == ADV_OS7MP && (g.OLx < 4 || g.OLy < 4)) {
    fail(42, "gad_calc_rhs_b200_: OS7MP needs OLx, OLy >= 4 (gad_check.F overlap test)");
    return;
  }
  const int K = *k, Nr = g.Nr;
  if (K < 1 || K > Nr || *kUp < 1 || *kUp > 2 || *kDown < 1 || *kDown > 2) { fail(40, "bad level index"); return; }
  TileGrid tg;
  if (!make_tile_grid(*bi, *bj, tg)) { if (!c.lastError) fail(43, "grid mirrors not set"); return; }
  const size_t ns = g.slab;
  SlabAcc a;
  a.OLx = g.OLx; a.OLy = g.OLy; a.PX = g.PX; a.slab = ns;
  const double *slabs[11] = {xA, yA, maskUp, uFld, vFld, wFld, uTrans, vTrans, rTrans, rTransKp1, KappaR};
  const double *dev[11];
  for (int n = 0; n < 11; n++) {
    dev[n] = to_device(slabs[n], ns, 10 + n, true);
    if (!dev[n]) return;
  }
  a.xA_ = dev[0]; a.yA_ = dev[1]; a.maskUp_ = dev[2]; a.uFld_ = dev[3]; a.vFld_ = dev[4]; a.wFld_ = dev[5];
  a.uTrans_ = dev[6]; a.vTrans_ = dev[7]; a.rTrans_ = dev[8]; a.rTransKp1_ = dev[9]; a.KappaR_ = dev[10];
  // tracer levels the stencils can touch: k-2 .. k+1 (k-3 .. k+2 is a safe superset); the 7-point
  // OS7MP stencil reaches k-4 .. k+3 (gad_os7mp_adv_r.F:100-107)
  const int reach = (*vertAdvecSchArg == ADV_OS7MP) ? 4 : 3;
  const int kLo = std::max(1, K - reach), kHi = std::min(Nr, K + reach - 1);
  auto stage3 = [&](const double *h, int slot) -> double * {
    if (is_device_ptr(h)) return const_cast<double *>(h);
    double *d = to_device(h, ns * Nr, slot, false);
    if (!d) return nullptr;
    if (cudaMemcpyAsync(d + ns * (kLo - 1), h + ns * (kLo - 1), ns * (size_t)(kHi - kLo + 1) * sizeof(double),
                        cudaMemcpyHostToDevice, c.stream) != cudaSuccess) { fail(4, "H2D copy failed"); return nullptr; }
    return d;
  };
  a.TracerN = stage3(TracerN, 30);
  a.TracAB = *applyAB_onTracer ? stage3(TracAB, 31) : a.TracerN;
  if (!a.TracerN || !a.TracAB) return;
  // gTracer: level k in/out
  double *gT = const_cast<double *>(gTracer);
  const bool gTdev = is_device_ptr(gTracer);
  if (!gTdev) {
    gT = to_device(gTracer, ns * Nr, 32, false);
    if (!gT) return;
    cudaMemcpyAsync(gT + ns * (K - 1), gTracer + ns * (K - 1), ns * sizeof(double), cudaMemcpyHostToDevice, c.stream);
  }
  double *fV = fVerT;
  const bool fVdev = is_device_ptr(fVerT);
  if (!fVdev) {
    fV = to_device(fVerT, ns * 2, 33, false);
    if (!fV) return;
    cudaMemcpyAsync(fV + ns * (*kDown - 1), fVerT + ns * (*kDown - 1), ns * sizeof(double), cudaMemcpyHostToDevice, c.stream);
  }
  double *fZ = to_device(fZon, ns, 34, false), *fM = to_device(fMer, ns, 35, false);
  if (!fZ || !fM) return;
  GadPar p;
  p.k = K; p.advScheme = *advectionSchArg; p.vertAdvScheme = *vertAdvecSchArg; p.calcAdvection = *calcAdvection;
  p.implicitAdvection = *implicitAdvection; p.applyAB = *applyAB_onTracer; p.useDiffKr4 = *trUseDiffKr4;
  p.implicitDiffusion = c.p.I(MI_IMPLICITDIFFUSION);
  p.diffKh = *diffKh; p.diffK4 = *diffK4; p.rkSign = c.p.D(MP_RKSIGN);
  p.deltaT = deltaTLev[K - 1];
  p.diffKr4k = *trUseDiffKr4 ? diffKr4[K - 1] : 0.0;
  dim3 blk(32, 8), grd((g.PX + 31) / 32, (g.PY + 7) / 8);
  c.launches++;
  gad_level_kernel<<<grd, blk, 0, c.stream>>>(tg, a, p, fZ, fM, fV + ns * (*kUp - 1), fV + ns * (*kDown - 1), gT);
  if (cudaGetLastError() != cudaSuccess) { fail(5, "gad_level_kernel launch failed"); return; }
  if (!from_device(fZon, fZ, ns) || !from_device(fMer, fM, ns)) return;
  if (!fVdev) cudaMemcpyAsync(fVerT + ns * (*kUp - 1), fV + ns * (*kUp - 1), ns * sizeof(double), cudaMemcpyDeviceToHost, c.stream);
  if (!gTdev) cudaMemcpyAsync(gTracer + ns * (K - 1), gT + ns * (K - 1), ns * sizeof(double), cudaMemcpyDeviceToHost, c.stream);
  if (cudaStreamSynchronize(c.stream) != cudaSuccess) fail(6, "gad_calc_rhs_b200_: stream error");
}
