// vecinv.cu -- MOM_VECINV drop-in (pkg/mom_vecinv/mom_vecinv.F:10-1009; caller dynamics.F:527):
// per (tile, level) call with the reference argument list plus the COMMON /DYNVARS_R/ arrays and the
// two facts MOM_CALC_RELVORT3 / FILL_CS_CORNER_TR_RL read from W2_EXCH2_TOPOLOGY.h (facet corners,
// facet number).  2 launches per call, 4 with biharmonic viscosity; see vecinv.cuh.
#include <algorithm>
#include <cstdlib>
#include "vecinv.cuh"

namespace mg {
bool make_mom_par(MomPar &p);

bool make_vi_par(ViPar &p) {
  const Params &q = ctx().p;
  if (!make_mom_par(p.m)) return false;
  p.useCoriolis = q.I(MI_USECORIOLIS); p.useAbsVorticity = q.I(MI_USEABSVORTICITY);
  p.selectVortScheme = q.I(MI_SELECTVORTSCHEME); p.useJamartMomAdv = q.I(MI_USEJAMARTMOMADV);
  p.upwindShear = q.I(MI_UPWINDSHEAR); p.selectKEscheme = q.I(MI_SELECTKESCHEME);
  p.harmonic = (p.m.viscAhD != 0. || p.m.viscAhZ != 0.) ? 1 : 0;
  p.highOrderVorticity = q.I(MI_HIGHORDERVORTICITY); p.upwindVorticity = q.I(MI_UPWINDVORTICITY);
  if ((p.highOrderVorticity || p.upwindVorticity) && p.selectVortScheme != 0 && p.selectVortScheme != 2)
    return fail(53, "mom_vecinv_b200_: MOM_VI_*_CORIOLIS_C4 implements selectVortScheme 0 and 2 only");
  if (q.I(MI_MOMIMPLVERTADV)) return fail(53, "mom_vecinv_b200_: momImplVertAdv is not on the B200 path");
  if (p.selectVortScheme < 0 || p.selectVortScheme > 3) return fail(53, "mom_vecinv_b200_: selectVortScheme not implemented");
  if (p.m.selectCoriScheme < 0 || p.m.selectCoriScheme > 3) return fail(53, "mom_vecinv_b200_: invalid selectCoriScheme");
  if (p.selectKEscheme < -1 || p.selectKEscheme > 3) return fail(53, "mom_vecinv_b200_: invalid selectKEscheme");
  return true;
}
}  // namespace mg

using namespace mg;

extern "C" void mom_vecinv_b200_(const int *bi, const int *bj, const int *k, const int *iMin, const int *iMax,
                                 const int *jMin, const int *jMax, const double *kappaRU, const double *kappaRV,
                                 const double *fVerUkm, const double *fVerVkm, double *fVerUkp, double *fVerVkp,
                                 double *guDiss, double *gvDiss, const double *myTime, const int *myIter, const int *myThid,
                                 const double *uVel, const double *vVel, const double *wVel, double *gU, double *gV,
                                 const int *csCorners, const int *myFace) {
  (void)myTime; (void)myIter; (void)myThid;
  Ctx &c = ctx();
  c.lastError = 0;
  if (!c.ready) { fail(30, "mitgcm_b200_init_ not called"); return; }
  const Geom &g = c.g;
  const int K = *k, Nr = g.Nr;
  if (K < 1 || K > Nr) { fail(40, "bad level index"); return; }
  if (*iMin < 2 - g.OLx || *iMax > g.sNx + g.OLx - 1 || *jMin < 2 - g.OLy || *jMax > g.sNy + g.OLy - 1) {
    fail(52, "mom_vecinv_b200_: iMin..iMax / jMin..jMax must leave a one-point rim inside the halo");
    return;
  }
  if (*csCorners < 0 || *csCorners > 15 || (*csCorners && (*myFace < 1 || g.sNx + 1 > g.PX))) { fail(53, "mom_vecinv_b200_: bad facet corner mask"); return; }
  ViPar p;
  if (!make_vi_par(p)) return;
  p.csCorners = *csCorners; p.myFace = *myFace;
  p.iMin = *iMin; p.iMax = *iMax; p.jMin = *jMin; p.jMax = *jMax;
  TileGrid tg;
  if (!make_tile_grid(*bi, *bj, tg)) { if (!c.lastError) fail(43, "grid mirrors not set"); return; }
  if (!tg.recip_rAz || !tg.fCoriG || !tg.recip_dxG || !tg.recip_dyG) { fail(43, "mom_vecinv_b200_: rAz / fCoriG / dxG / dyG mirrors not set"); return; }
  const size_t ns = g.slab, tile = (size_t)(*bi - 1) + (size_t)g.nSx * (size_t)(*bj - 1);
  const size_t off3 = ns * Nr * tile;
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
  MomState st{dU + off3, dV + off3, dW + off3, dKU, dKV};
  double *dgU = is_device_ptr(gU) ? gU : to_device(gU, g.n3, 45, false);
  double *dgV = is_device_ptr(gV) ? gV : to_device(gV, g.n3, 46, false);
  if (!dgU || !dgV) return;
  // slabs: fVer?km in; fVer?kp keep the points outside iMin..iMax; guDiss / gvDiss are fully written
  const double *slabIn[4] = {fVerUkm, fVerVkm, fVerUkp, fVerVkp};
  double *slabD[6];
  for (int n = 0; n < 4; n++) {
    slabD[n] = to_device(slabIn[n], ns, 47 + n, true);
    if (!slabD[n]) return;
  }
  slabD[4] = to_device(guDiss, ns, 51, false);
  slabD[5] = to_device(gvDiss, ns, 52, false);
  static double dummy;
  double *scr = to_device(&dummy, 9 * ns, 53, false);
  if (!slabD[4] || !slabD[5] || !scr) return;
  ViScratch w{scr, scr + ns, scr + 2 * ns, scr + 3 * ns, scr + 4 * ns, scr + 5 * ns, scr + 6 * ns, scr + 7 * ns, scr + 8 * ns};
  if (!is_device_ptr(gU)) {   // keep the points outside iMin..iMax as the caller left them
    cudaMemcpyAsync(dgU + off3 + ns * (K - 1), gU + off3 + ns * (K - 1), ns * sizeof(double), cudaMemcpyHostToDevice, c.stream);
    cudaMemcpyAsync(dgV + off3 + ns * (K - 1), gV + off3 + ns * (K - 1), ns * sizeof(double), cudaMemcpyHostToDevice, c.stream);
  }
  dim3 blk(32, 8), grd((g.PX + 31) / 32, (g.PY + 7) / 8);
  const bool biharm = p.m.momViscosity && p.m.useBiharmonicVisc;
  if (!biharm && !getenv("MITGCM_B200_VI_STAGED")) {
    // one launch: vorticity, KE and divergence are re-evaluated where they are read
    c.launches++;
    vi_tend_kernel<true><<<grd, blk, 0, c.stream>>>(tg, st, p, K, w, slabD[0], slabD[1], slabD[2], slabD[3], slabD[4],
                                                    slabD[5], dgU + off3, dgV + off3);
  } else {
    c.launches += 2;
    vi_stage1_kernel<<<grd, blk, 0, c.stream>>>(tg, st, p, K, w);
    if (biharm) {
      c.launches += 2;
      vi_del2_kernel<<<grd, blk, 0, c.stream>>>(tg, p, K, w);
      vi_star_kernel<<<grd, blk, 0, c.stream>>>(tg, p, K, w);
    }
    vi_tend_kernel<false><<<grd, blk, 0, c.stream>>>(tg, st, p, K, w, slabD[0], slabD[1], slabD[2], slabD[3], slabD[4],
                                                     slabD[5], dgU + off3, dgV + off3);
  }
  if (cudaGetLastError() != cudaSuccess) { fail(5, "mom_vecinv kernel launch failed"); return; }
  double *slabOut[4] = {fVerUkp, fVerVkp, guDiss, gvDiss};
  for (int n = 0; n < 4; n++)
    if (!from_device(slabOut[n], slabD[2 + n], ns)) return;
  if (!is_device_ptr(gU)) {
    cudaMemcpyAsync(gU + off3 + ns * (K - 1), dgU + off3 + ns * (K - 1), ns * sizeof(double), cudaMemcpyDeviceToHost, c.stream);
    cudaMemcpyAsync(gV + off3 + ns * (K - 1), dgV + off3 + ns * (K - 1), ns * sizeof(double), cudaMemcpyDeviceToHost, c.stream);
  }
  if (cudaStreamSynchronize(c.stream) != cudaSuccess) fail(6, "mom_vecinv_b200_: stream error");
}
