// update_cg2d.cu -- UPDATE_CG2D (model/src/update_cg2d.F:57-192) on the device mirrors: rebuilds the CG2D operator
// aW2d, aS2d, aC2d (and, every cg2dPreCondFreq steps, the preconditioner pC, pW, pS) from the CURRENT hFacW / hFacS
// -- what the model does once per step with the non-linear free surface / r* (configs 3 and 4), when the column
// thickness follows eta.  cg2dNorm and the tolerance stay as INI_CG2D left them (update_cg2d.F:41-42).
// Three point-wise kernels in the reference's loop ranges and operation order (bit-identical to the Fortran):
//   aW, aS on 1..sN+1 (zero elsewhere: the reference does not exchange them here), the k sum marched in order;
//   aC on the interior, then EXCH_XY_RS(aC2d); the preconditioner on 1..sN+1.
// Not supported (refused by the shim's build options, not reachable through the parameters here): deepAtmosphere,
// OBCS masks, selectImplicitDrag = 2.
#include "context.h"

namespace mg {

struct UpdCg2dArgs {
  int sNx, sNy, OLx, OLy, PX, PY, Nr, nTiles;
  size_t slab;
  const double *dyG, *dxG, *recip_dxC, *recip_dyC, *drF, *hFacW, *hFacS, *recip_Bo, *rA;
  double *aW, *aS, *aC, *pW, *pS, *pC;
  double cg2dNorm, implicSurfPress, implicDiv2DFlow, freeSurfFac, deltaTMom, deltaTFreeSurf, pcOffDFac;
};

__global__ void upd_cg2d_aws_kernel(UpdCg2dArgs a) {
  const int ii = blockIdx.x * blockDim.x + threadIdx.x, jj = blockIdx.y * blockDim.y + threadIdx.y, t = blockIdx.z;
  if (ii >= a.PX || jj >= a.PY) return;
  const int i = ii - a.OLx + 1, j = jj - a.OLy + 1;      // Fortran indices
  const size_t s = (size_t)ii + (size_t)a.PX * jj + a.slab * t;
  double aw = 0., as = 0.;
  if (i >= 1 && i <= a.sNx + 1 && j >= 1 && j <= a.sNy + 1) {
    const size_t s3 = (size_t)ii + (size_t)a.PX * jj + a.slab * (size_t)a.Nr * t;
    for (int k = 0; k < a.Nr; k++) {
      double faceArea = a.dyG[s] * a.drF[k] * a.hFacW[s3 + a.slab * k];
      aw = aw + faceArea * a.recip_dxC[s];
      faceArea = a.dxG[s] * a.drF[k] * a.hFacS[s3 + a.slab * k];
      as = as + faceArea * a.recip_dyC[s];
    }
    aw = aw * a.cg2dNorm * a.implicSurfPress * a.implicDiv2DFlow;
    as = as * a.cg2dNorm * a.implicSurfPress * a.implicDiv2DFlow;
  }
  a.aW[s] = aw;
  a.aS[s] = as;
}

__global__ void upd_cg2d_ac_kernel(UpdCg2dArgs a) {
  const int i = 1 + blockIdx.x * blockDim.x + threadIdx.x, j = 1 + blockIdx.y * blockDim.y + threadIdx.y, t = blockIdx.z;
  if (i > a.sNx || j > a.sNy) return;
  const size_t s = (size_t)(i + a.OLx - 1) + (size_t)a.PX * (j + a.OLy - 1) + a.slab * t;
  a.aC[s] = -(a.aW[s] + a.aW[s + 1] + a.aS[s] + a.aS[s + a.PX] +
              a.freeSurfFac * a.cg2dNorm * a.recip_Bo[s] * a.rA[s] / a.deltaTMom / a.deltaTFreeSurf);
}

__global__ void upd_cg2d_pc_kernel(UpdCg2dArgs a) {
  const int i = 1 + blockIdx.x * blockDim.x + threadIdx.x, j = 1 + blockIdx.y * blockDim.y + threadIdx.y, t = blockIdx.z;
  if (i > a.sNx + 1 || j > a.sNy + 1) return;
  const size_t s = (size_t)(i + a.OLx - 1) + (size_t)a.PX * (j + a.OLy - 1) + a.slab * t;
  const double aC = a.aC[s];
  a.pC[s] = aC == 0. ? 1. : 1. / aC;
  const double pW_tmp = aC + a.aC[s - 1];
  if (pW_tmp == 0.) a.pW[s] = 0.;
  else {
    const double f = a.pcOffDFac * pW_tmp;
    a.pW[s] = -a.aW[s] / (f * f);
  }
  const double pS_tmp = aC + a.aC[s - a.PX];
  if (pS_tmp == 0.) a.pS[s] = 0.;
  else {
    const double f = a.pcOffDFac * pS_tmp;
    a.pS[s] = -a.aS[s] / (f * f);
  }
}

static bool update_cg2d(int myIter) {
  Ctx &c = ctx();
  if (!c.ready) return fail(30, "mitgcm_b200_init_ not called");
  const Geom &g = c.g;
  const Params &q = c.p;
  UpdCg2dArgs a;
  a.sNx = g.sNx; a.sNy = g.sNy; a.OLx = g.OLx; a.OLy = g.OLy; a.PX = g.PX; a.PY = g.PY; a.Nr = g.Nr; a.nTiles = g.nTiles;
  a.slab = g.slab;
  a.dyG = field(MG_DYG, false); a.dxG = field(MG_DXG, false); a.recip_dxC = field(MG_RECIP_DXC, false);
  a.recip_dyC = field(MG_RECIP_DYC, false); a.drF = field(MG_DRF, false); a.hFacW = field(MG_HFACW, false);
  a.hFacS = field(MG_HFACS, false); a.recip_Bo = field(MG_RECIP_BO, false); a.rA = field(MG_RA, false);
  if (!a.dyG || !a.dxG || !a.recip_dxC || !a.recip_dyC || !a.drF || !a.hFacW || !a.hFacS || !a.recip_Bo || !a.rA)
    return fail(43, "update_cg2d: grid mirrors not set (dxG, dyG, recip_dxC, recip_dyC, drF, hFacW, hFacS, recip_Bo, rA)");
  a.aW = field(MG_AW2D); a.aS = field(MG_AS2D); a.aC = field(MG_AC2D); a.pW = field(MG_PW); a.pS = field(MG_PS); a.pC = field(MG_PC);
  if (!a.aW || !a.aS || !a.aC || !a.pW || !a.pS || !a.pC) return false;
  a.cg2dNorm = q.D(MP_CG2DNORM); a.implicSurfPress = q.D(MP_IMPLICSURFPRESS); a.implicDiv2DFlow = q.D(MP_IMPLICDIV2DFLOW);
  a.freeSurfFac = q.D(MP_FREESURFFAC); a.deltaTMom = q.D(MP_DELTATMOM); a.deltaTFreeSurf = q.D(MP_DELTATFREESURF);
  a.pcOffDFac = q.D(MP_CG2DPCOFFDFAC);
  if (a.cg2dNorm == 0. || a.deltaTMom == 0. || a.deltaTFreeSurf == 0.)
    return fail(44, "update_cg2d: cg2dNorm / deltaTMom / deltaTFreeSurf not set (INI_CG2D comes first)");
  // update_cg2d.F:54-60
  const int freq = q.I(MI_CG2DPRECONDFREQ);
  bool updatePreCond = false;
  if (freq != 0) {
    updatePreCond = myIter == q.I(MI_NITER0);
    if (myIter % freq == 0) updatePreCond = true;
  }
  const dim3 blk(32, 8);
  c.launches += 2;
  upd_cg2d_aws_kernel<<<dim3((g.PX + 31) / 32, (g.PY + 7) / 8, g.nTiles), blk, 0, c.stream>>>(a);
  upd_cg2d_ac_kernel<<<dim3((g.sNx + 31) / 32, (g.sNy + 7) / 8, g.nTiles), blk, 0, c.stream>>>(a);
  MG_CUDA(cudaGetLastError());
  if (!updatePreCond) return true;
  if (!exch_field(a.aC, 1)) return false;      // EXCH_XY_RS(aC2d): periodic tiling, exch2 tile graph, or across ranks
  c.launches++;
  upd_cg2d_pc_kernel<<<dim3((g.sNx + 1 + 31) / 32, (g.sNy + 1 + 7) / 8, g.nTiles), blk, 0, c.stream>>>(a);
  MG_CUDA(cudaGetLastError());
  return true;
}

}  // namespace mg

extern "C" void update_cg2d_b200_(const double *myTime, const int *myIter, const int *myThid) {
  (void)myTime; (void)myThid;
  mg::ctx().lastError = 0;
  mg::update_cg2d(*myIter);
}
