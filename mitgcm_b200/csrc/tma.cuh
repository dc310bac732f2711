// tma.cuh -- the few pieces of the sm_100a bulk-tensor copy engine (TMA) the 3-D stencil kernels use:
// tensor maps of the eesupp tile arrays (host), mbarrier + cp.async.bulk.tensor (device, inline PTX).
// A tile3d array (1-OLx:sNx+OLx, 1-OLy:sNy+OLy, Nr, nSx, nSy) is described as the rank-3 tensor
// (PX, PY, Nr*nTiles) of doubles; one copy fetches a (boxX, boxY, 1) patch of one level into shared memory,
// rows packed densely (boxX doubles per row), cells outside the array zero-filled by the hardware.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>

namespace mg {

// ---- host ------------------------------------------------------------------------------------------
typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                        const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                        CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_tmapEncodeTiled tmap_encoder() {
  static PFN_tmapEncodeTiled fn = nullptr;
  if (!fn) {      // the driver entry point through the runtime: no link-time dependency on libcuda
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && p)
      fn = reinterpret_cast<PFN_tmapEncodeTiled>(p);
  }
  return fn;
}

// (PX, PY, nz) doubles at base, box (boxX, boxY, 1).  Needs PX even (16-byte row pitch), boxX even, and at copy time an
// EVEN x coordinate: the engine raises an illegal-instruction fault for a box row that does not start 16-byte aligned.
inline bool make_tmap3(CUtensorMap *m, const double *base, int PX, int PY, size_t nz, int boxX, int boxY) {
  PFN_tmapEncodeTiled enc = tmap_encoder();
  if (!enc) return false;
  cuuint64_t dims[3] = {(cuuint64_t)PX, (cuuint64_t)PY, (cuuint64_t)nz};
  cuuint64_t strides[2] = {(cuuint64_t)PX * 8, (cuuint64_t)PX * (cuuint64_t)PY * 8};
  cuuint32_t box[3] = {(cuuint32_t)boxX, (cuuint32_t)boxY, 1};
  cuuint32_t es[3] = {1, 1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<double *>(base), dims, strides, box, es,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// ---- device ----------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  const uint32_t a = smem_u32(bar);
  uint32_t ok;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(a), "r"(parity)
        : "memory");
  } while (!ok);
}
// one (boxX, boxY, 1) patch at array coordinates (x, y, z) -> dst (128-byte aligned shared memory)
__device__ __forceinline__ void tma_load3(void *dst, const CUtensorMap *m, int x, int y, int z, uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar))
      : "memory");
}

}  // namespace mg
