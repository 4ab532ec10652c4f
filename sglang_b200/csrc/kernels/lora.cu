// Multi-adapter LoRA for a batch whose tokens use different adapters (SGMV): two small kernels per wrapped linear, graph-capturable.
//
//   shrink:  h[i, s, r]  = sum_k x[perm[i], k] * A[slot][s][r][k]              (s = sub-projection: q/k/v or gate/up; r < R)
//   expand:  y[perm[i], n] += sum_r h[i, sub(n), r] * B[slot][n][r]            (the adapter's alpha / r scaling is folded into B at load)
//
// Tokens are grouped by adapter slot on the host (stable sort -> `perm`) and cut into tiles of <= 16 tokens that share one slot:
// (tile_slot, tile_start, tile_cnt).  All of that lives in one STATIC int32 tensor that the host refreshes before every step, the grids
// are sized by capacity (unused tiles have cnt 0 and exit at once), so the kernels sit inside the captured decode graph — LoRA batches no
// longer fall off CUDA graphs.  Adapter factors live in a slot-paged pool: A_pool [slots][n_sub][R][K], B_pool [slots][N][R], slot 0 = none.
//
// Arithmetic intensity is tiny (R <= 64): CUDA-core FMAs with 16-byte loads; the A / B panels of the few adapters of a batch stay in L2.
//
// Reference counterparts: python/sglang/srt/lora/mem_pool.py (paged adapter pool), kernels/ops/gemm/sgemm_lora_{a,b}.py and the chunked
// SGMV Triton kernels (segment-grouped shrink / expand).
#include "common.cuh"
#include "tma_host.h"

namespace sglb {

constexpr int LORA_TM = 16;        // tokens per tile
constexpr int LORA_THREADS = 256;
constexpr int LORA_KC = 512;       // K chunk staged in shared memory (16 tokens x 512 x bf16 = 16 KB)

// grid (tile capacity, n_sub); thread = (rank row r, token quarter)
template <int R>
__global__ void __launch_bounds__(LORA_THREADS)
lora_shrink_kernel(const __nv_bfloat16* __restrict__ x, int ldx, const __nv_bfloat16* __restrict__ A_pool, int n_sub, int K,
                   const int* __restrict__ tile_slot, const int* __restrict__ tile_start, const int* __restrict__ tile_cnt,
                   const int* __restrict__ perm, float* __restrict__ h) {
  const int tile = blockIdx.x, sub = blockIdx.y;
  const int cnt = tile_cnt[tile];
  if (cnt <= 0) return;
  const int slot = tile_slot[tile], start = tile_start[tile];
  __shared__ __align__(16) __nv_bfloat16 xs[LORA_TM][LORA_KC];
  constexpr int GROUPS = (LORA_THREADS / R) < LORA_TM ? (LORA_THREADS / R) : LORA_TM;   // token groups in parallel (R = 64 -> 4 groups of 4 tokens)
  constexpr int TPG = LORA_TM / GROUPS;               // tokens per group
  const int r = threadIdx.x % R, grp = threadIdx.x / R;
  const __nv_bfloat16* a_row = A_pool + (((size_t)slot * n_sub + sub) * R + r) * (size_t)K;
  float acc[TPG];
#pragma unroll
  for (int i = 0; i < TPG; ++i) acc[i] = 0.f;
  for (int k0 = 0; k0 < K; k0 += LORA_KC) {
    const int kc = min(LORA_KC, K - k0);
    __syncthreads();
    for (int idx = threadIdx.x; idx < LORA_TM * (LORA_KC / 8); idx += LORA_THREADS) {
      const int i = idx / (LORA_KC / 8), v = idx % (LORA_KC / 8);
      U4 val = U4{0, 0, 0, 0};
      if (i < cnt && v * 8 < kc) val = ldg128(x + (size_t)perm[start + i] * ldx + k0 + v * 8);
      *reinterpret_cast<U4*>(&xs[i][v * 8]) = val;
    }
    __syncthreads();
    if (grp < GROUPS)
    for (int k = 0; k < kc; k += 8) {
      const U4 a = ldg128(a_row + k0 + k);
      const float af[8] = {bf16_lo(a.x), bf16_hi(a.x), bf16_lo(a.y), bf16_hi(a.y), bf16_lo(a.z), bf16_hi(a.z), bf16_lo(a.w), bf16_hi(a.w)};
#pragma unroll
      for (int i = 0; i < TPG; ++i) {
        const U4 xv = *reinterpret_cast<const U4*>(&xs[grp * TPG + i][k]);      // same address across the R lanes of a group: broadcast
        acc[i] += af[0] * bf16_lo(xv.x) + af[1] * bf16_hi(xv.x) + af[2] * bf16_lo(xv.y) + af[3] * bf16_hi(xv.y) + af[4] * bf16_lo(xv.z) +
                  af[5] * bf16_hi(xv.z) + af[6] * bf16_lo(xv.w) + af[7] * bf16_hi(xv.w);
      }
    }
  }
  if (grp < GROUPS) {
#pragma unroll
    for (int i = 0; i < TPG; ++i) {
      const int t = grp * TPG + i;
      if (t < cnt) h[((size_t)(start + t) * n_sub + sub) * R + r] = acc[i];
    }
  }
}

// grid (tile capacity, ceil(N / 256)); thread = output column
template <int R>
__global__ void __launch_bounds__(LORA_THREADS)
lora_expand_kernel(const float* __restrict__ h, const __nv_bfloat16* __restrict__ B_pool, int n_sub, int N, int b1, int b2,
                   const int* __restrict__ tile_slot, const int* __restrict__ tile_start, const int* __restrict__ tile_cnt,
                   const int* __restrict__ perm, __nv_bfloat16* __restrict__ y, int ldy) {
  const int tile = blockIdx.x;
  const int cnt = tile_cnt[tile];
  if (cnt <= 0) return;
  const int slot = tile_slot[tile], start = tile_start[tile];
  __shared__ float hs[LORA_TM][3][R];
  for (int idx = threadIdx.x; idx < LORA_TM * n_sub * R; idx += LORA_THREADS) {
    const int i = idx / (n_sub * R), rem = idx % (n_sub * R);
    hs[i][rem / R][rem % R] = (i < cnt) ? h[((size_t)(start + i) * n_sub) * R + rem] : 0.f;
  }
  __syncthreads();
  const int n = blockIdx.y * LORA_THREADS + threadIdx.x;
  if (n >= N) return;
  const int sub = (n >= b2) ? 2 : (n >= b1 ? 1 : 0);       // which sub-projection this output column belongs to
  const __nv_bfloat16* b_row = B_pool + ((size_t)slot * N + n) * R;
  float acc[LORA_TM];
#pragma unroll
  for (int i = 0; i < LORA_TM; ++i) acc[i] = 0.f;
#pragma unroll
  for (int r0 = 0; r0 < R; r0 += 8) {
    const U4 b = ldg128(b_row + r0);
    const float bf[8] = {bf16_lo(b.x), bf16_hi(b.x), bf16_lo(b.y), bf16_hi(b.y), bf16_lo(b.z), bf16_hi(b.z), bf16_lo(b.w), bf16_hi(b.w)};
#pragma unroll
    for (int i = 0; i < LORA_TM; ++i) {
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i] += hs[i][sub][r0 + j] * bf[j];
    }
  }
  for (int i = 0; i < cnt; ++i) {
    __nv_bfloat16* dst = y + (size_t)perm[start + i] * ldy + n;
    *dst = __float2bfloat16(__bfloat162float(*dst) + acc[i]);
  }
}

void lora_sgmv(const void* x, int ldx, const void* A_pool, const void* B_pool, int rank_cap, int n_sub, int K, int N, int b1, int b2,
               const int* tile_slot, const int* tile_start, const int* tile_cnt, const int* perm, int tile_cap, float* h, void* y, int ldy,
               cudaStream_t stream) {
  if (tile_cap <= 0) return;
  if (n_sub < 1 || n_sub > 3) throw std::runtime_error("sglb lora: 1..3 sub-projections per wrapped linear");
  if (K % 8 || ldx % 8) throw std::runtime_error("sglb lora: K and the input row stride must be multiples of 8");
  dim3 g1(tile_cap, n_sub), g2(tile_cap, (N + LORA_THREADS - 1) / LORA_THREADS);
#define SGLB_LORA_LAUNCH(R_)                                                                                                              \
  do {                                                                                                                                    \
    lora_shrink_kernel<R_><<<g1, LORA_THREADS, 0, stream>>>((const __nv_bfloat16*)x, ldx, (const __nv_bfloat16*)A_pool, n_sub, K, tile_slot, \
                                                            tile_start, tile_cnt, perm, h);                                               \
    lora_expand_kernel<R_><<<g2, LORA_THREADS, 0, stream>>>(h, (const __nv_bfloat16*)B_pool, n_sub, N, b1, b2, tile_slot, tile_start,     \
                                                            tile_cnt, perm, (__nv_bfloat16*)y, ldy);                                      \
  } while (0)
  switch (rank_cap) {
    case 8: SGLB_LORA_LAUNCH(8); break;
    case 16: SGLB_LORA_LAUNCH(16); break;
    case 32: SGLB_LORA_LAUNCH(32); break;
    case 64: SGLB_LORA_LAUNCH(64); break;
    default: throw std::runtime_error("sglb lora: pool rank must be 8, 16, 32 or 64");
  }
#undef SGLB_LORA_LAUNCH
  SGLB_CUDA_CHECK(cudaGetLastError());
}

}  // namespace sglb
