// Host-side TMA descriptor helpers (cuTensorMapEncodeTiled resolved through the runtime so we do not
// link libcuda directly; the CPU-only build box has no driver).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdexcept>
#include <string>

namespace sglb {

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    if (e != cudaSuccess || p == nullptr) throw std::runtime_error("sglb: cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
    fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

// rank-N tiled map over a tensor whose innermost dim is contiguous. dims/strides innermost-first;
// strides_bytes has rank-1 entries (dims 1..rank-1).
inline CUtensorMap make_tmap(const void* base, CUtensorMapDataType dt, int rank, const uint64_t* dims,
                             const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swz,
                             CUtensorMapL2promotion l2 = CU_TENSOR_MAP_L2_PROMOTION_L2_256B) {
  CUtensorMap m;
  cuuint64_t gd[5], gs[5];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
  for (int i = 0; i < rank - 1; ++i) gs[i] = strides_bytes[i];
  CUresult r = get_encode_fn()(&m, dt, (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, es,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, swz, l2, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[256];
    snprintf(buf, sizeof(buf), "sglb: cuTensorMapEncodeTiled failed (%d) rank=%d dims=[%llu,%llu,%llu] box=[%u,%u,%u]", (int)r,
             rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
             (unsigned long long)(rank > 2 ? dims[2] : 0), box[0], rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0);
    throw std::runtime_error(buf);
  }
  return m;
}

#define SGLB_CUDA_CHECK(x)                                                                              \
  do {                                                                                                  \
    cudaError_t e__ = (x);                                                                              \
    if (e__ != cudaSuccess)                                                                             \
      throw std::runtime_error(std::string("sglb CUDA error: ") + cudaGetErrorString(e__) + " at " + __FILE__ + ":" + \
                               std::to_string(__LINE__));                                               \
  } while (0)

}  // namespace sglb
