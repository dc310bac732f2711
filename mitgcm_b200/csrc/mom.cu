// mom.cu -- MOM_FLUXFORM drop-in (pkg/mom_fluxform/mom_fluxform.F:42-1064): one launch per
// (tile, level) call with the reference argument list plus the COMMON /DYNVARS_R/ arrays.
// See mom.cuh for the point-wise form.
#include "mom.cuh"

namespace mg {

bool make_mom_par(MomPar &p) {
  const Params &q = ctx().p;
  p.viscAhD = q.D(MP_VISCAHD); p.viscAhZ = q.D(MP_VISCAHZ); p.viscA4D = q.D(MP_VISCA4D); p.viscA4Z = q.D(MP_VISCA4Z);
  p.sideDragFactor = q.D(MP_SIDEDRAGFACTOR); p.bottomDragLinear = q.D(MP_BOTTOMDRAGLINEAR);
  p.bottomDragQuadratic = q.D(MP_BOTTOMDRAGQUADRATIC); p.recip_rSphere = q.D(MP_RECIP_RSPHERE);
  p.afFacMom = q.D(MP_AFFACMOM); p.vfFacMom = q.D(MP_VFFACMOM); p.cfFacMom = q.D(MP_CFFACMOM); p.mtFacMom = q.D(MP_MTFACMOM);
  p.rkSign = q.D(MP_RKSIGN);
  p.momAdvection = q.I(MI_MOMADVECTION); p.momViscosity = q.I(MI_MOMVISCOSITY);
  p.useBiharmonicVisc = q.I(MI_USEBIHARMONICVISC); p.implicitViscosity = q.I(MI_IMPLICITVISCOSITY);
  p.no_slip_sides = q.I(MI_NO_SLIP_SIDES); p.no_slip_bottom = q.I(MI_NO_SLIP_BOTTOM);
  p.bottomVisc_pCell = q.I(MI_BOTTOMVISC_PCELL); p.selectBotDragQuadr = q.I(MI_SELECTBOTDRAGQUADR);
  // mom_fluxform.F:272-279
  p.bottomDragTerms = (q.I(MI_SELECTIMPLICITDRAG) == 0 &&
                       (p.no_slip_bottom || p.selectBotDragQuadr >= 0 || p.bottomDragLinear != 0.)) ? 1 : 0;
  p.useCDscheme = q.I(MI_USECDSCHEME); p.selectCoriScheme = q.I(MI_SELECTCORISCHEME);
  p.metricTerms = q.I(MI_SELECTMETRICTERMS) >= 1;
  p.usingSphericalPolarGrid = q.I(MI_USINGSPHERICALPOLARGRID); p.rigidLid = q.I(MI_RIGIDLID);
  p.select_rStar = q.I(MI_SELECT_RSTAR);
  if (p.select_rStar != 0) return fail(51, "mom_fluxform_b200_: r* coordinate is not on the B200 path");
  if (p.no_slip_sides && p.momViscosity && p.sideDragFactor <= 0.)
    return fail(51, "mom_fluxform_b200_: sideDragFactor <= 0 is not on the B200 path");
  if (p.selectBotDragQuadr < -1 || p.selectBotDragQuadr > 2) return fail(51, "invalid selectBotDragQuadr");
  return true;
}

__global__ void __launch_bounds__(256) mom_level_kernel(TileGrid g, MomState st, MomPar p, int k, int iMin, int iMax,
                                                        int jMin, int jMax, double *fVerUkm, double *fVerVkm,
                                                        double *fVerUkp, double *fVerVkp, double *guDiss,
                                                        double *gvDiss, double *gU, double *gV) {
  const int i = 1 - g.OLx + blockIdx.x * 32 + threadIdx.x;
  const int j = 1 - g.OLy + blockIdx.y * 8 + threadIdx.y;
  if (i > g.sNx + g.OLx || j > g.sNy + g.OLy) return;
  const size_t s = g.s(i, j);
  const bool inner = i >= 2 - g.OLx && j >= 2 - g.OLy;   // range of MOM_{U,V}_ADV_W{U,V}
  double ukm = fVerUkm[s], vkm = fVerVkm[s], ukp = fVerUkp[s], vkp = fVerVkp[s];
  if (p.momAdvection) {
    if (k == 1) {   // mom_fluxform.F:384-417
      if (p.rigidLid) { ukm = 0.; vkm = 0.; }
      else if (inner) { ukm = mom_adv_wu(g, st, p, 1, i, j); vkm = mom_adv_wv(g, st, p, 1, i, j); }
      fVerUkm[s] = ukm; fVerVkm[s] = vkm;
    }
    if (k + 1 > g.Nr) { ukp = 0.; vkp = 0.; }
    else if (inner) { ukp = mom_adv_wu(g, st, p, k + 1, i, j); vkp = mom_adv_wv(g, st, p, k + 1, i, j); }
    fVerUkp[s] = ukp; fVerVkp[s] = vkp;
  }
  const size_t s3 = g.s3(i, j, k);
  if (i >= iMin && i <= iMax && j >= jMin && j <= jMax) {
    MomOut o = mom_cell(g, st, p, k, i, j, ukm, ukp, vkm, vkp);
    gU[s3] = o.gU; gV[s3] = o.gV; guDiss[s] = o.guDiss; gvDiss[s] = o.gvDiss;
  } else {
    guDiss[s] = 0.; gvDiss[s] = 0.;                       // mom_fluxform.F:203-232
    if (!p.momAdvection) { gU[s3] = 0.; gV[s3] = 0.; }    // :548-553, :806-811
  }
}

}  // namespace mg

using namespace mg;

extern "C" void mom_fluxform_b200_(const int *bi, const int *bj, const int *k, const int *iMin, const int *iMax,
                                   const int *jMin, const int *jMax, const double *kappaRU, const double *kappaRV,
                                   double *fVerUkm, double *fVerVkm, double *fVerUkp, double *fVerVkp, double *guDiss,
                                   double *gvDiss, const double *myTime, const int *myIter, const int *myThid,
                                   const double *uVel, const double *vVel, const double *wVel, double *gU, double *gV) {
  (void)myTime; (void)myIter; (void)myThid;
  Ctx &c = ctx();
  c.lastError = 0;
  if (!c.ready) { fail(30, "mitgcm_b200_init_ not called"); return; }
  const Geom &g = c.g;
  const int K = *k, Nr = g.Nr;
  if (K < 1 || K > Nr) { fail(40, "bad level index"); return; }
  // the point-wise form is valid where every stencil stays inside the ranges the reference
  // leaves defined: 2-OL <= iMin, iMax <= sN+OL-1 (dynamics.F:191-192 uses 0..sN+1)
  if (*iMin < 2 - g.OLx || *iMax > g.sNx + g.OLx - 1 || *jMin < 2 - g.OLy || *jMax > g.sNy + g.OLy - 1) {
    fail(52, "mom_fluxform_b200_: iMin..iMax / jMin..jMax must leave a one-point rim inside the halo");
    return;
  }
  MomPar p;
  if (!make_mom_par(p)) return;
  if (p.useBiharmonicVisc && (*iMin < 3 - g.OLx || *iMax > g.sNx + g.OLx - 2 || *jMin < 3 - g.OLy || *jMax > g.sNy + g.OLy - 2)) {
    fail(52, "mom_fluxform_b200_: biharmonic viscosity needs OLx, OLy >= 3 for this range (gad_check.F:110-126)");
    return;
  }
  TileGrid tg;
  if (!make_tile_grid(*bi, *bj, tg)) { if (!c.lastError) fail(43, "grid mirrors not set"); return; }
  const size_t ns = g.slab, tile = (size_t)(*bi - 1) + (size_t)g.nSx * (size_t)(*bj - 1);
  const size_t off3 = ns * Nr * tile, off3p = ns * (Nr + 1) * tile;
  // stage the levels of the 3-D arrays this level touches (host pointers only)
  auto stage3 = [&](const double *h, int slot, size_t total, size_t off, int kLo, int kHi) -> double * {
    if (is_device_ptr(h)) return const_cast<double *>(h);
    double *d = to_device(h, total, slot, false);
    if (!d) return nullptr;
    if (cudaMemcpyAsync(d + off + ns * (kLo - 1), h + off + ns * (kLo - 1), ns * (size_t)(kHi - kLo + 1) * sizeof(double),
                        cudaMemcpyHostToDevice, c.stream) != cudaSuccess) { fail(4, "H2D copy failed"); return nullptr; }
    return d;
  };
  const int kLo = std::max(1, K - 1), kHi = std::min(Nr, K + 1);
  const double *dU = stage3(uVel, 40, g.n3, off3, kLo, kHi), *dV = stage3(vVel, 41, g.n3, off3, kLo, kHi);
  const double *dW = stage3(wVel, 42, g.n3, off3, K, kHi);
  const double *dKU = stage3(kappaRU, 43, ns * (Nr + 1), 0, K, K + 1), *dKV = stage3(kappaRV, 44, ns * (Nr + 1), 0, K, K + 1);
  if (!dU || !dV || !dW || !dKU || !dKV) return;
  (void)off3p;
  MomState st{dU + off3, dV + off3, dW + off3, dKU, dKV};
  double *dgU = is_device_ptr(gU) ? gU : to_device(gU, g.n3, 45, false);
  double *dgV = is_device_ptr(gV) ? gV : to_device(gV, g.n3, 46, false);
  double *slabH[6] = {fVerUkm, fVerVkm, fVerUkp, fVerVkp, guDiss, gvDiss};
  double *slabD[6];
  for (int n = 0; n < 6; n++) {
    // fVer?km is an input for k > 1; fVer?kp keeps its untouched rim -> upload all four
    slabD[n] = to_device(slabH[n], ns, 47 + n, n < 4);
    if (!slabD[n]) return;
  }
  if (!dgU || !dgV) return;
  if (!is_device_ptr(gU)) {   // keep the points outside iMin..iMax as the caller left them
    cudaMemcpyAsync(dgU + off3 + ns * (K - 1), gU + off3 + ns * (K - 1), ns * sizeof(double), cudaMemcpyHostToDevice, c.stream);
    cudaMemcpyAsync(dgV + off3 + ns * (K - 1), gV + off3 + ns * (K - 1), ns * sizeof(double), cudaMemcpyHostToDevice, c.stream);
  }
  dim3 blk(32, 8), grd((g.PX + 31) / 32, (g.PY + 7) / 8);
  c.launches++;
  mom_level_kernel<<<grd, blk, 0, c.stream>>>(tg, st, p, K, *iMin, *iMax, *jMin, *jMax, slabD[0], slabD[1], slabD[2],
                                              slabD[3], slabD[4], slabD[5], dgU + off3, dgV + off3);
  if (cudaGetLastError() != cudaSuccess) { fail(5, "mom_level_kernel launch failed"); return; }
  for (int n = 0; n < 6; n++)
    if (!from_device(slabH[n], slabD[n], ns)) return;
  if (!is_device_ptr(gU)) {
    cudaMemcpyAsync(gU + off3 + ns * (K - 1), dgU + off3 + ns * (K - 1), ns * sizeof(double), cudaMemcpyDeviceToHost, c.stream);
    cudaMemcpyAsync(gV + off3 + ns * (K - 1), dgV + off3 + ns * (K - 1), ns * sizeof(double), cudaMemcpyDeviceToHost, c.stream);
  }
  if (cudaStreamSynchronize(c.stream) != cudaSuccess) fail(6, "mom_fluxform_b200_: stream error");
}
