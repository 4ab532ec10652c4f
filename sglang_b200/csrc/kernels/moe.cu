// Mixture-of-experts kernels for sm_100a (everything around the expert-grouped tcgen05 GEMM in gemm_tc.cu):
//
//   moe_topk_softmax   one warp per token: softmax over E router logits in registers, K rounds of warp arg-max,
//                      optional renormalisation of the selected weights.
//   moe_align          counting sort of the T*K (token, k) pairs by LOCAL expert into 128-row-padded segments
//                      (the GEMM's M tile), producing sorted_row[pos] -> pair, inv_pos[pair] -> pos, the per-row
//                      routing weight, tile_expert[m_tile] and the number of live M tiles — all in device memory, so
//                      the whole MoE block is CUDA-graph capturable with no host round trip.
//   moe_gather         xs[pos,:] = x[sorted_row[pos] / K, :]     (zeros for padding rows)
//   moe_combine        out[t,:] = sum_k ys[inv_pos[t*K+k],:]     (fixed k order -> deterministic), optionally PUSHED into
//                      slot [rank] of every peer's NVLink arena so the expert-parallel all-reduce needs no further kernel
//                      on the producer side (consumer: nvl_local_reduce_add_rmsnorm in comm/nvl_allreduce.cu).
//
// Reference counterparts: kernels/aot/csrc/moe/moe_topk_softmax_kernels.cu:628-715, moe_align_kernel.cu:350-386,
// python/sglang/srt/layers/moe/fused_moe_triton/fused_moe_triton_kernels.py:324 (grouped GEMM), moe_sum_reduce.
#include "common.cuh"
#include "tma_host.h"

namespace sglb {

// ---------------------------------------------------------------- top-k softmax
template <int VPT, typename T>
__global__ void __launch_bounds__(128) moe_topk_softmax_kernel(const T* __restrict__ logits, float* __restrict__ out_w,
                                                               int* __restrict__ out_ids, int n_tokens, int E, int K, int stride,
                                                               int renorm) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t = blockIdx.x * 4 + warp;
  if (t >= n_tokens) return;
  float v[VPT];
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    const int e = i * 32 + lane;
    v[i] = e < E ? (float)logits[(size_t)t * stride + e] : -INFINITY;
    mx = fmaxf(mx, v[i]);
  }
  mx = warp_max(mx);
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    v[i] = (i * 32 + lane) < E ? __expf(v[i] - mx) : -1.f;
    sum += fmaxf(v[i], 0.f);
  }
  sum = warp_sum(sum);
  const float inv = 1.f / sum;
  float sel_sum = 0.f;
  float my_w = 0.f;
  int my_id = 0;
  for (int k = 0; k < K; ++k) {
    float best = -2.f;
    int best_e = 0x7fffffff;
#pragma unroll
    for (int i = 0; i < VPT; ++i)
      if (v[i] > best) { best = v[i]; best_e = i * 32 + lane; }
    // warp arg-max; ties -> lowest expert id (matches torch.topk on distinct values, deterministic otherwise)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oe = __shfl_xor_sync(0xffffffffu, best_e, o);
      if (ob > best || (ob == best && oe < best_e)) { best = ob; best_e = oe; }
    }
    if ((best_e & 31) == lane) {
#pragma unroll
      for (int i = 0; i < VPT; ++i)
        if (i == (best_e >> 5)) v[i] = -1.f;
    }
    sel_sum += best;
    if (lane == k) { my_w = best; my_id = best_e; }
  }
  if (lane < K) {
    out_w[(size_t)t * K + lane] = renorm ? my_w / sel_sum : my_w * inv;
    out_ids[(size_t)t * K + lane] = my_id;
  }
}

void moe_topk_softmax(const void* logits, bool bf16, float* out_w, int* out_ids, int T, int E, int K, int stride, bool renorm,
                      cudaStream_t stream) {
  if (T <= 0) return;
  if (E > 512 || K > 32 || K > E) throw std::runtime_error("sglb moe_topk_softmax: need E <= 512, K <= min(32, E)");
  const int grid = (T + 3) / 4;
  const int vpt = (E + 31) / 32;
#define SGLB_TOPK(V)                                                                                                              \
  do {                                                                                                                            \
    if (bf16) moe_topk_softmax_kernel<V, __nv_bfloat16><<<grid, 128, 0, stream>>>((const __nv_bfloat16*)logits, out_w, out_ids, T, \
                                                                                   E, K, stride, renorm);                         \
    else moe_topk_softmax_kernel<V, float><<<grid, 128, 0, stream>>>((const float*)logits, out_w, out_ids, T, E, K, stride, renorm); \
  } while (0)
  if (vpt <= 1) SGLB_TOPK(1); else if (vpt <= 2) SGLB_TOPK(2); else if (vpt <= 4) SGLB_TOPK(4); else if (vpt <= 8) SGLB_TOPK(8);
  else SGLB_TOPK(16);
#undef SGLB_TOPK
  SGLB_CUDA_CHECK(cudaGetLastError());
}

// ---------------------------------------------------------------- align (single CTA counting sort)
constexpr int MOE_MAX_E = 512;
constexpr int MOE_TILE = 128;

__global__ void __launch_bounds__(1024) moe_align_kernel(const int* __restrict__ topk_ids, const float* __restrict__ topk_w,
                                                         int n_pairs, int E, int expert_offset, int* __restrict__ sorted_row,
                                                         int* __restrict__ inv_pos, float* __restrict__ row_w,
                                                         int* __restrict__ tile_expert, int* __restrict__ num_tiles, int max_tiles) {
  __shared__ int cnt[MOE_MAX_E];
  __shared__ int start[MOE_MAX_E + 1];
  __shared__ int cursor[MOE_MAX_E];
  for (int e = threadIdx.x; e < E; e += blockDim.x) { cnt[e] = 0; cursor[e] = 0; }
  __syncthreads();
  for (int i = threadIdx.x; i < n_pairs; i += blockDim.x) {
    const int e = topk_ids[i] - expert_offset;
    if (e >= 0 && e < E) atomicAdd(&cnt[e], 1);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int run = 0, tiles = 0;
    for (int e = 0; e < E; ++e) {
      start[e] = run;
      const int nt = (cnt[e] + MOE_TILE - 1) / MOE_TILE;
      for (int j = 0; j < nt && tiles < max_tiles; ++j) tile_expert[tiles++] = e;
      run += nt * MOE_TILE;
    }
    start[E] = run;
    *num_tiles = tiles;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n_pairs; i += blockDim.x) {
    const int e = topk_ids[i] - expert_offset;
    if (e >= 0 && e < E) {
      const int p = start[e] + atomicAdd(&cursor[e], 1);
      sorted_row[p] = i;
      row_w[p] = topk_w[i];
      inv_pos[i] = p;
    } else {
      inv_pos[i] = -1;
    }
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < E * MOE_TILE; idx += blockDim.x) {
    const int e = idx / MOE_TILE, j = idx % MOE_TILE;
    const int p = start[e] + cnt[e] + j;
    if (p < start[e + 1]) { sorted_row[p] = -1; row_w[p] = 0.f; }
  }
}

void moe_align(const int* topk_ids, const float* topk_w, int n_pairs, int E, int expert_offset, int* sorted_row, int* inv_pos,
               float* row_w, int* tile_expert, int* num_tiles, int max_tiles, cudaStream_t stream) {
  if (E > MOE_MAX_E) throw std::runtime_error("sglb moe_align: more than 512 local experts");
  moe_align_kernel<<<1, 1024, 0, stream>>>(topk_ids, topk_w, n_pairs, E, expert_offset, sorted_row, inv_pos, row_w, tile_expert,
                                           num_tiles, max_tiles);
  SGLB_CUDA_CHECK(cudaGetLastError());
}

// ---------------------------------------------------------------- gather into expert-sorted order
__global__ void __launch_bounds__(256) moe_gather_kernel(const __nv_bfloat16* __restrict__ x, const int* __restrict__ sorted_row,
                                                         const int* __restrict__ num_tiles, __nv_bfloat16* __restrict__ xs, int H,
                                                         int x_stride, int K, int cap_rows) {
  const int live = min(cap_rows, __ldg(num_tiles) * MOE_TILE);
  const int nvec = H >> 3;
  for (int row = blockIdx.x; row < live; row += gridDim.x) {
    const int pair = __ldg(sorted_row + row);
    __nv_bfloat16* dst = xs + (size_t)row * H;
    if (pair >= 0) {
      const __nv_bfloat16* src = x + (size_t)(pair / K) * x_stride;
      for (int v = threadIdx.x; v < nvec; v += blockDim.x) st128(dst + v * 8, ldg128(src + v * 8));
    } else {
      for (int v = threadIdx.x; v < nvec; v += blockDim.x) st128(dst + v * 8, U4{0, 0, 0, 0});
    }
  }
}

void moe_gather(const void* x, const int* sorted_row, const int* num_tiles, void* xs, int H, int x_stride, int K, int cap_rows,
                int num_sms, cudaStream_t stream) {
  if (H % 8) throw std::runtime_error("sglb moe_gather: hidden size must be a multiple of 8");
  const int grid = max(1, min(cap_rows, num_sms * 8));
  moe_gather_kernel<<<grid, 256, 0, stream>>>((const __nv_bfloat16*)x, sorted_row, num_tiles, (__nv_bfloat16*)xs, H, x_stride, K,
                                              cap_rows);
  SGLB_CUDA_CHECK(cudaGetLastError());
}

// ---------------------------------------------------------------- combine (+ optional NVLink push)
struct MoePush {
  __nv_bfloat16* peer_slot[8];
  uint32_t* peer_flags[8];
  const uint32_t* epoch_ptr;
  uint32_t* done_ctr;
  long long slot_stride;
  int rank, world;
};

template <bool PUSH>
__global__ void __launch_bounds__(256) moe_combine_kernel(const __nv_bfloat16* __restrict__ ys, const int* __restrict__ inv_pos,
                                                          const float* __restrict__ pair_w, __nv_bfloat16* __restrict__ out, int T,
                                                          int H, int K, const MoePush push) {
  const int nvec = H >> 3;
  for (int t = blockIdx.x; t < T; t += gridDim.x) {
    for (int v = threadIdx.x; v < nvec; v += blockDim.x) {
      float a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      for (int k = 0; k < K; ++k) {
        const int p = __ldg(inv_pos + t * K + k);
        if (p < 0) continue;
        const float w = pair_w ? __ldg(pair_w + t * K + k) : 1.f;
        const U4 y = ld128(ys + (size_t)p * H + v * 8);
        a[0] += w * bf16_lo(y.x); a[1] += w * bf16_hi(y.x); a[2] += w * bf16_lo(y.y); a[3] += w * bf16_hi(y.y);
        a[4] += w * bf16_lo(y.z); a[5] += w * bf16_hi(y.z); a[6] += w * bf16_lo(y.w); a[7] += w * bf16_hi(y.w);
      }
      const U4 o = U4{pack_bf16(a[0], a[1]), pack_bf16(a[2], a[3]), pack_bf16(a[4], a[5]), pack_bf16(a[6], a[7])};
      if (PUSH) {
        const size_t off = (size_t)push.rank * push.slot_stride + (size_t)t * H + v * 8;
#pragma unroll
        for (int r = 0; r < 8; ++r)
          if (r < push.world) st128(push.peer_slot[r] + off, o);
      } else {
        st128(out + (size_t)t * H + v * 8, o);
      }
    }
  }
  if (PUSH) {
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      const uint32_t n = atomicAdd(push.done_ctr, 1u);
      if (n == gridDim.x - 1) {
        *push.done_ctr = 0;
        __threadfence_system();
        const uint32_t epoch = *((volatile const uint32_t*)push.epoch_ptr) + 1;
        for (int r = 0; r < push.world; ++r)
          asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(push.peer_flags[r] + push.rank), "r"(epoch) : "memory");
      }
    }
  }
}

void moe_combine(const void* ys, const int* inv_pos, const float* pair_w, void* out, int T, int H, int K, int num_sms,
                 cudaStream_t stream) {
  if (T <= 0) return;
  const int grid = max(1, min(T, num_sms * 8));
  moe_combine_kernel<false><<<grid, 256, 0, stream>>>((const __nv_bfloat16*)ys, inv_pos, pair_w, (__nv_bfloat16*)out, T, H, K, MoePush{});
  SGLB_CUDA_CHECK(cudaGetLastError());
}

void moe_combine_push(const void* ys, const int* inv_pos, const float* pair_w, int T, int H, int K, const int64_t* peer_slot_ptrs,
                      const int64_t* peer_flag_ptrs, int rank, int world, const uint32_t* epoch_ptr, uint32_t* done_ctr, int num_sms,
                      cudaStream_t stream) {
  if (T <= 0) return;
  if (world > 8) throw std::runtime_error("sglb moe_combine_push: world size > 8");
  MoePush p{};
  for (int r = 0; r < world; ++r) {
    p.peer_slot[r] = reinterpret_cast<__nv_bfloat16*>(peer_slot_ptrs[r]);
    p.peer_flags[r] = reinterpret_cast<uint32_t*>(peer_flag_ptrs[r]);
  }
  p.epoch_ptr = epoch_ptr; p.done_ctr = done_ctr; p.slot_stride = (long long)T * H; p.rank = rank; p.world = world;
  const int grid = max(1, min(T, num_sms * 4));
  moe_combine_kernel<true><<<grid, 256, 0, stream>>>((const __nv_bfloat16*)ys, inv_pos, pair_w, nullptr, T, H, K, p);
  SGLB_CUDA_CHECK(cudaGetLastError());
}

}  // namespace sglb
