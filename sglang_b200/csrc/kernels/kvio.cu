// KV page movers between the device KV pool and the pinned-host pool of the hierarchical cache (L1 <-> L2), all layers in one launch.
//
// The host pool is pinned memory mapped into the device address space (UVA), so the SMs read / write it directly: one CTA per
// (page, layer x {K,V} plane) moves the page's contiguous chunk with 16-byte accesses — no staging tensor, no per-layer cudaMemcpy, no host
// scatter.  Host layouts: `layer_first` [planes][host_pages][chunk] (mirrors the device pool) and `page_first` [host_pages][planes][chunk]
// (every host page is one contiguous record -> one read / write per page for the L3 storage backends).
//
// Reference counterparts: kernels/aot/csrc/kvcacheio/transfer.cu (transfer_kv_all_layer / _lf_pf / direct variants),
// srt/mem_cache/memory_pool_host.py (layouts).
#include "common.cuh"
#include "tma_host.h"

namespace sglb {

struct KvIoParams {
  uint8_t* dev;               // device pool base
  uint8_t* host;              // pinned host pool base (device-visible pointer)
  const int* dev_pages;       // [n]
  const int* host_pages;      // [n]
  long long dev_plane_stride;   // bytes between planes in the device pool
  long long host_plane_stride;  // bytes between planes in the host pool (layer_first) or inside a page record (page_first)
  long long host_page_stride;   // bytes between host pages
  int chunk_bytes;              // bytes of one page in one plane (page_size * kv_heads * head_dim * itemsize)
  int to_host;
};

__global__ void __launch_bounds__(256) kv_transfer_pages_kernel(const KvIoParams p) {
  const int i = blockIdx.x, plane = blockIdx.y;
  uint8_t* d = p.dev + (long long)plane * p.dev_plane_stride + (long long)p.dev_pages[i] * p.chunk_bytes;
  uint8_t* h = p.host + (long long)plane * p.host_plane_stride + (long long)p.host_pages[i] * p.host_page_stride;
  const U4* src = reinterpret_cast<const U4*>(p.to_host ? d : h);
  U4* dst = reinterpret_cast<U4*>(p.to_host ? h : d);
  const int n16 = p.chunk_bytes >> 4;
  // 4 independent 16-byte loads in flight per thread before the stores (PCIe round trips are long)
  for (int k = threadIdx.x; k < n16; k += 4 * 256) {
    U4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (k + u * 256 < n16) v[u] = src[k + u * 256];
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (k + u * 256 < n16) dst[k + u * 256] = v[u];
  }
}

void kv_transfer_pages(void* dev, void* host_devptr, const int* dev_pages, const int* host_pages, int n, int planes, long long dev_plane_stride,
                       long long host_plane_stride, long long host_page_stride, int chunk_bytes, bool to_host, cudaStream_t stream) {
  if (n <= 0) return;
  if (chunk_bytes % 16) throw std::runtime_error("sglb kvio: page chunk must be a multiple of 16 bytes");
  KvIoParams p{(uint8_t*)dev, (uint8_t*)host_devptr, dev_pages, host_pages, dev_plane_stride, host_plane_stride, host_page_stride, chunk_bytes,
               to_host ? 1 : 0};
  kv_transfer_pages_kernel<<<dim3(n, planes), 256, 0, stream>>>(p);
  SGLB_CUDA_CHECK(cudaGetLastError());
}

}  // namespace sglb
