// Micro-benchmark: how fast can ONE persistent CTA per SM gather paged KV tiles with TMA, and what limits it?
// Same access pattern as attention_decode.cu (tile = 4 pages x 32 rows x 128 bf16 of one KV head, K and V pools [slots][Hkv][128],
// 128B-swizzled 64-element half boxes) with the math removed: a consumer warp waits for the tile, optionally spins `consume` cycles
// (stand-in for the S -> softmax -> PV chain) and releases the stage.  Variants: box shape (2 x 3D half boxes per page | 1 x 5D box per
// page), ring depth, one producer for K and V | one producer each, K only.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I sglang_b200/csrc/kernels tools/tma_gather_bench.cu -o tools/bin/tma_gather_bench
#include "common.cuh"
#include "tma_host.h"
#include <algorithm>
#include <random>
#include <vector>

using namespace sglb;

constexpr int PS = 32, D = 128, PAGES_PER_TILE = 4;
constexpr int HALF_BYTES = PAGES_PER_TILE * PS * 128;   // 16 KB: 128 rows x 64 bf16
constexpr int TILE_BYTES = 2 * HALF_BYTES;              // 32 KB per tensor

__device__ __forceinline__ void tma_load_5d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}

struct P {
  const int* tile_pages;   // [n_tiles][4]
  const int* tile_head;    // [n_tiles]
  int n_tiles;
  int box5d;               // 1: one 5D box per page (both halves), 0: two 3D half boxes per page
  int kv;                  // 1: K and V, 0: K only
  int consume;             // cycles the consumer holds each tensor's stage
  int chain;               // 1: consumer handles K then V of a tile serially (like S -> softmax -> PV); 0: release as soon as landed
};

template <int STAGES, int NPROD>
__global__ void __launch_bounds__(128) gather_kernel(const __grid_constant__ CUtensorMap tmK3, const __grid_constant__ CUtensorMap tmV3,
                                                     const __grid_constant__ CUtensorMap tmK5, const __grid_constant__ CUtensorMap tmV5, P p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sK = smem;
  uint8_t* sV = smem + STAGES * TILE_BYTES;
  uint64_t* bars = (uint64_t*)(smem + 2 * STAGES * TILE_BYTES);
  uint64_t *k_full = bars, *k_empty = bars + STAGES, *v_full = bars + 2 * STAGES, *v_empty = bars + 3 * STAGES;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1); mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1); }
    fence_barrier_init();
  }
  __syncthreads();
  auto load_tile = [&](const CUtensorMap* m3, const CUtensorMap* m5, uint8_t* dst0, uint64_t* full, int t) {
    mbar_expect_tx(full, TILE_BYTES);
    const int h = p.tile_head[t];
#pragma unroll
    for (int i = 0; i < PAGES_PER_TILE; ++i) {
      const int page = p.tile_pages[t * PAGES_PER_TILE + i];
      uint8_t* dst = dst0 + i * PS * 128;
      if (p.box5d) {
        // smem box order: [half][row][64]; the half stride inside the tile is HALF_BYTES only if the box covers the whole tile; per page the two
        // halves land 4 KB apart here (layout differs from the attention kernel's; irrelevant for a bandwidth probe)
        tma_load_5d(dst0 + i * 2 * PS * 128, m5, full, 0, 0, 0, h, page);
      } else {
        tma_load_3d(dst, m3, full, 0, h, page * PS);
        tma_load_3d(dst + HALF_BYTES, m3, full, 64, h, page * PS);
      }
    }
  };
  if (warp == 0 && lane == 0) {                   // producer (K, and V when NPROD == 1)
    int n = 0;
    for (int t = blockIdx.x; t < p.n_tiles; t += gridDim.x, ++n) {
      const int st = n % STAGES, ph = (n / STAGES) & 1;
      mbar_wait(&k_empty[st], ph ^ 1);
      load_tile(&tmK3, &tmK5, sK + st * TILE_BYTES, &k_full[st], t);
      if (NPROD == 1 && p.kv) {
        mbar_wait(&v_empty[st], ph ^ 1);
        load_tile(&tmV3, &tmV5, sV + st * TILE_BYTES, &v_full[st], t);
      }
    }
  } else if (warp == 1 && lane == 0 && NPROD == 2 && p.kv) {   // second producer: V
    int n = 0;
    for (int t = blockIdx.x; t < p.n_tiles; t += gridDim.x, ++n) {
      const int st = n % STAGES, ph = (n / STAGES) & 1;
      mbar_wait(&v_empty[st], ph ^ 1);
      load_tile(&tmV3, &tmV5, sV + st * TILE_BYTES, &v_full[st], t);
    }
  } else if (warp == 2 && lane == 0) {            // consumer
    int n = 0;
    for (int t = blockIdx.x; t < p.n_tiles; t += gridDim.x, ++n) {
      const int st = n % STAGES, ph = (n / STAGES) & 1;
      mbar_wait(&k_full[st], ph);
      if (p.consume) { const long long t0 = clock64(); while (clock64() - t0 < p.consume) {} }
      mbar_arrive(&k_empty[st]);
      if (p.kv) {
        mbar_wait(&v_full[st], ph);
        if (p.consume && p.chain) { const long long t0 = clock64(); while (clock64() - t0 < p.consume) {} }
        mbar_arrive(&v_empty[st]);
      }
    }
  }
}

// plain streaming read of both pools (ld.global.nc 16 B per thread, grid-stride): the read-only ceiling the gathers are compared against
__global__ void __launch_bounds__(1024) read_kernel(const uint4* __restrict__ a, const uint4* __restrict__ b, size_t n, unsigned* sink) {
  unsigned acc = 0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const uint4 x = __ldg(a + i), y = __ldg(b + i);
    acc ^= x.x ^ x.y ^ x.z ^ x.w ^ y.x ^ y.y ^ y.z ^ y.w;
  }
  if (acc == 0x12345678u) *sink = acc;
}

template <int STAGES, int NPROD>
float run(const CUtensorMap& k3, const CUtensorMap& v3, const CUtensorMap& k5, const CUtensorMap& v5, P p, int grid, void* flush, size_t flush_bytes) {
  const int smem = 2 * STAGES * TILE_BYTES + 1024 + 256;
  cudaFuncSetAttribute(gather_kernel<STAGES, NPROD>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  std::vector<float> ts;
  for (int it = 0; it < 6; ++it) {
    cudaMemsetAsync(flush, it, flush_bytes);
    cudaEventRecord(e0);
    gather_kernel<STAGES, NPROD><<<grid, 128, smem>>>(k3, v3, k5, v5, p);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); exit(1); }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (it) ts.push_back(ms);
  }
  std::sort(ts.begin(), ts.end());
  return ts[ts.size() / 2];
}

int main(int argc, char** argv) {
  const int Hkv = argc > 1 ? atoi(argv[1]) : 8;
  const int n_pages = argc > 2 ? atoi(argv[2]) : 8192 * 8 / Hkv;      // ~512 MB per tensor
  cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
  const int sms = prop.multiProcessorCount;
  const size_t pool_bytes = (size_t)n_pages * PS * Hkv * D * 2;
  void *kp, *vp, *flush;
  cudaMalloc(&kp, pool_bytes); cudaMalloc(&vp, pool_bytes);
  cudaMemset(kp, 1, pool_bytes); cudaMemset(vp, 2, pool_bytes);
  const size_t flush_bytes = 256u << 20;
  cudaMalloc(&flush, flush_bytes);
  // tiles: every (page, head) exactly once, pages shuffled, 4 pages of the same head per tile
  std::mt19937 rng(0);
  std::vector<int> tp, th;
  {
    // heads of the same 4 pages are adjacent tiles: they run on neighbouring CTAs at about the same time (what the request-major queue does)
    std::vector<int> perm(n_pages);
    for (int i = 0; i < n_pages; ++i) perm[i] = i;
    std::shuffle(perm.begin(), perm.end(), rng);
    for (int i = 0; i + PAGES_PER_TILE <= n_pages; i += PAGES_PER_TILE)
      for (int h = 0; h < Hkv; ++h) {
        for (int j = 0; j < PAGES_PER_TILE; ++j) tp.push_back(perm[i + j]);
        th.push_back(h);
      }
  }
  const int n_tiles = (int)th.size();
  int *d_tp, *d_th;
  cudaMalloc(&d_tp, tp.size() * 4); cudaMalloc(&d_th, th.size() * 4);
  cudaMemcpy(d_tp, tp.data(), tp.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(d_th, th.data(), th.size() * 4, cudaMemcpyHostToDevice);
  auto map3 = [&](void* base) {
    const uint64_t dims[3] = {(uint64_t)D, (uint64_t)Hkv, (uint64_t)n_pages * PS};
    const uint64_t strides[2] = {(uint64_t)D * 2, (uint64_t)Hkv * D * 2};
    const uint32_t box[3] = {64, 1, (uint32_t)PS};
    return make_tmap(base, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
  };
  auto map5 = [&](void* base) {
    // (64 elems, row in page, half, head, page)
    const uint64_t dims[5] = {64, (uint64_t)PS, 2, (uint64_t)Hkv, (uint64_t)n_pages};
    const uint64_t strides[4] = {(uint64_t)Hkv * D * 2, 128, (uint64_t)D * 2, (uint64_t)PS * Hkv * D * 2};
    const uint32_t box[5] = {64, (uint32_t)PS, 2, 1, 1};
    return make_tmap(base, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
  };
  CUtensorMap k3 = map3(kp), v3 = map3(vp), k5, v5;
  bool have5 = true;
  try { k5 = map5(kp); v5 = map5(vp); } catch (const std::exception& e) { printf("5D map rejected: %s\n", e.what()); have5 = false; k5 = k3; v5 = v3; }
  printf("Hkv=%d pages=%d tiles=%d pool=%.0f MB per tensor, %d SMs\n", Hkv, n_pages, n_tiles, pool_bytes / 1e6, sms);
  auto report = [&](const char* name, float ms, int kv) {
    const double bytes = (double)n_tiles * TILE_BYTES * (kv ? 2 : 1);
    printf("%-58s %8.1f us  %7.1f GB/s\n", name, ms * 1e3, bytes / ms / 1e6);
    fflush(stdout);
  };
  {
    unsigned* sink; cudaMalloc(&sink, 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9f;
    for (int it = 0; it < 5; ++it) {
      cudaMemsetAsync(flush, it, flush_bytes);
      cudaEventRecord(e0);
      read_kernel<<<sms * 2, 1024>>>((const uint4*)kp, (const uint4*)vp, pool_bytes / 16, sink);
      cudaEventRecord(e1); cudaDeviceSynchronize();
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (it) best = std::min(best, ms);
    }
    printf("%-58s %8.1f us  %7.1f GB/s\n", "streaming read of both pools (LDG.128)", best * 1e3, 2.0 * pool_bytes / best / 1e6);
  }
  P p{d_tp, d_th, n_tiles, 0, 1, 0, 0};
  char nm[128];
  for (int consume : {0, 500, 1000, 1500, 2000}) {
    p.box5d = 0; p.kv = 1; p.consume = consume; p.chain = 1;
    snprintf(nm, sizeof nm, "3D halves, 3 stages, 1 producer, chain consume=%d", consume);
    report(nm, run<3, 1>(k3, v3, k5, v5, p, sms, flush, flush_bytes), 1);
  }
  p.consume = 0; p.chain = 0;
  report("3D halves, 2 stages, 1 producer", run<2, 1>(k3, v3, k5, v5, p, sms, flush, flush_bytes), 1);
  report("3D halves, 3 stages, 2 producers", run<3, 2>(k3, v3, k5, v5, p, sms, flush, flush_bytes), 1);
  p.consume = 1500; p.chain = 1;
  report("3D halves, 3 stages, 2 producers, chain consume=1500", run<3, 2>(k3, v3, k5, v5, p, sms, flush, flush_bytes), 1);
  p.consume = 0; p.chain = 0;
  p.kv = 0;
  report("3D halves, K only, 3 stages", run<3, 1>(k3, v3, k5, v5, p, sms, flush, flush_bytes), 0);
  report("3D halves, K only, 1 stage", run<1, 1>(k3, v3, k5, v5, p, sms, flush, flush_bytes), 0);
  report("3D halves, K only, 2 stages", run<2, 1>(k3, v3, k5, v5, p, sms, flush, flush_bytes), 0);
  if (have5) {
    p.kv = 1; p.box5d = 1;
    report("5D box per page, 3 stages, 1 producer", run<3, 1>(k3, v3, k5, v5, p, sms, flush, flush_bytes), 1);
    report("5D box per page, 3 stages, 2 producers", run<3, 2>(k3, v3, k5, v5, p, sms, flush, flush_bytes), 1);
    p.consume = 1500; p.chain = 1;
    report("5D box per page, 3 stages, 2 producers, chain consume=1500", run<3, 2>(k3, v3, k5, v5, p, sms, flush, flush_bytes), 1);
  }
  return 0;
}
