// SM-partitioned streams through CUDA green contexts (driver API, CUDA >= 12.4).
// Parity: reference sgl-kernel/csrc/spatial/greenctx_stream.cu (create_greenctx_stream_by_value: two streams over disjoint SM groups, used by
// PD-multiplexing to run prefill and decode side by side).  Here the split is N-way and the driver entry points are resolved at run time
// through cudaGetDriverEntryPoint, so the extension does not link libcuda and still loads on machines whose driver predates green contexts.
#include <cuda.h>
#include <cuda_runtime.h>

#include <stdexcept>
#include <string>
#include <vector>

#include "api.h"

namespace sglb {

namespace {
template <class F>
F entry(const char* name) {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint(name, &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || fn == nullptr)
    throw std::runtime_error(std::string("green contexts: the driver does not export ") + name);
  return reinterpret_cast<F>(fn);
}

void ck(CUresult r, const char* what) {
  if (r != CUDA_SUCCESS) throw std::runtime_error(std::string("green contexts: ") + what + " failed with CUresult " + std::to_string((int)r));
}
}  // namespace

// sm_counts: requested SMs for each partition but the last; the last partition receives what is left.  Returns, per partition,
// {stream handle, SM count actually provisioned} (the driver rounds requests up to its granularity: 8 SMs on sm_90+).
std::vector<long long> green_ctx_streams(int device, const std::vector<long long>& sm_counts) {
  using GetRes = CUresult (*)(CUdevice, CUdevResource*, CUdevResourceType);
  using Split = CUresult (*)(CUdevResource*, unsigned int*, const CUdevResource*, CUdevResource*, unsigned int, unsigned int);
  using GenDesc = CUresult (*)(CUdevResourceDesc*, CUdevResource*, unsigned int);
  using CtxCreate = CUresult (*)(CUgreenCtx*, CUdevResourceDesc, CUdevice, unsigned int);
  using StreamCreate = CUresult (*)(CUstream*, CUgreenCtx, unsigned int, int);
  using DevGet = CUresult (*)(CUdevice*, int);
  auto get_res = entry<GetRes>("cuDeviceGetDevResource");
  auto split = entry<Split>("cuDevSmResourceSplitByCount");
  auto gen_desc = entry<GenDesc>("cuDevResourceGenerateDesc");
  auto ctx_create = entry<CtxCreate>("cuGreenCtxCreate");
  auto stream_create = entry<StreamCreate>("cuGreenCtxStreamCreate");
  auto dev_get = entry<DevGet>("cuDeviceGet");

  if (cudaSetDevice(device) != cudaSuccess || cudaFree(nullptr) != cudaSuccess) throw std::runtime_error("green contexts: cannot initialise the device");
  CUdevice dev;
  ck(dev_get(&dev, device), "cuDeviceGet");
  CUdevResource rest;
  ck(get_res(dev, &rest, CU_DEV_RESOURCE_TYPE_SM), "cuDeviceGetDevResource");
  std::vector<CUdevResource> parts;
  for (long long want : sm_counts) {
    if (want <= 0) throw std::runtime_error("green contexts: SM counts must be positive");
    CUdevResource got, remaining;
    unsigned int groups = 1;
    ck(split(&got, &groups, &rest, &remaining, 0, (unsigned int)want), "cuDevSmResourceSplitByCount");
    if (groups < 1 || remaining.sm.smCount == 0) throw std::runtime_error("green contexts: not enough SMs left for the requested split");
    parts.push_back(got);
    rest = remaining;
  }
  parts.push_back(rest);
  std::vector<long long> out;
  for (auto& r : parts) {
    CUdevResourceDesc desc;
    ck(gen_desc(&desc, &r, 1), "cuDevResourceGenerateDesc");
    CUgreenCtx g;
    ck(ctx_create(&g, desc, dev, CU_GREEN_CTX_DEFAULT_STREAM), "cuGreenCtxCreate");       // lives for the life of the process
    CUstream s;
    ck(stream_create(&s, g, CU_STREAM_NON_BLOCKING, 0), "cuGreenCtxStreamCreate");
    out.push_back((long long)(uintptr_t)s);
    out.push_back((long long)r.sm.smCount);
  }
  return out;
}

}  // namespace sglb
