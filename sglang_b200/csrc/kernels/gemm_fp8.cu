// Block-scaled fp8 (e4m3) GEMM on tcgen05 for sm_100a + the activation quantiser that feeds it.
//
//   C[M,N] (bf16) = sum_kb  xs[m,kb] * ws[n/128,kb] * ( A[m, kb*128:(kb+1)*128] . W[n, kb*128:(kb+1)*128] )
//
//   A  [M,K]  e4m3, xs [M, K/128] fp32  (per-token, per-128-channel group scales — produced by per_token_group_quant_fp8)
//   W  [N,K]  e4m3, ws [N/128, K/128] fp32 (128x128 weight blocks — the DeepSeek-V3 / "fp8 block" checkpoint format)
//
// Scales change every 128-wide K block, so partial products cannot simply accumulate in TMEM.  Each K block is
// therefore ONE 128-byte-swizzled smem stage (128 fp8 = 128 B per row) = 4 tcgen05.mma kind::f8f6f4 (K=32 each) into a
// fresh TMEM accumulator stage; the epilogue warps drain that stage with tcgen05.ld, apply xs*ws and accumulate in fp32
// registers ("promotion") while the tensor core already fills the next TMEM stage (4 stages x 128 columns = all 512).
//
//   warp 0 : TMA producer      warp 1 : MMA issuer      warp 2 : TMEM alloc      warps 4..7 : promotion + epilogue
//
// Reference counterparts: python/sglang/srt/layers/quantization/fp8_kernel.py (per_token_group_quant_fp8,
// w8a8_block_fp8_matmul), fp8_utils.py:552-808 (DeepGEMM / cutlass / triton dispatch), sgl-kernel fp8_blockwise_gemm.
#include "common.cuh"
#include "tma_host.h"

#include <algorithm>

namespace sglb {

constexpr int F8_BM = 128;
constexpr int F8_BK = 128;          // fp8 elements = bytes per swizzled row = one scale group
constexpr int F8_THREADS = 256;

template <int BN>
struct F8Cfg {
  static constexpr int A_BYTES = F8_BM * F8_BK;
  static constexpr int B_BYTES = BN * F8_BK;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN >= 128) ? 6 : 8;
  static constexpr int NACC = 512 / BN > 4 ? 4 : 512 / BN;      // TMEM accumulator stages
  static constexpr int TMEM_COLS = NACC * BN;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
};

template <int BN>
__global__ void __launch_bounds__(F8_THREADS, 1)
gemm_fp8_block_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                      const float* __restrict__ xs, const float* __restrict__ ws, __nv_bfloat16* __restrict__ C, int M, int N, int K,
                      int ldc, int xs_stride, int ws_stride) {
  using Cfg = F8Cfg<BN>;
  constexpr int STAGES = Cfg::STAGES, NACC = Cfg::NACC;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * Cfg::A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* tfull = bars + 2 * STAGES;
  uint64_t* tempty = bars + 2 * STAGES + NACC;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 2 * NACC);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tiles = (M + F8_BM - 1) / F8_BM, n_tiles = (N + BN - 1) / BN;
  const int num_tiles = m_tiles * n_tiles;
  const int num_kb = (K + F8_BK - 1) / F8_BK;

  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB); }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < NACC; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 4); }
    fence_barrier_init();
  }
  if (warp == 2) { tmem_alloc(tmem_ptr, Cfg::TMEM_COLS); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0 && lane == 0) {
    int stage = 0; uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_blk = tile % m_tiles, n_blk = tile / m_tiles;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&empty[stage], phase ^ 1);
        mbar_expect_tx(&full[stage], Cfg::STAGE_BYTES);
        tma_load_2d(sA + stage * Cfg::A_BYTES, &tmA, &full[stage], kb * F8_BK, m_blk * F8_BM);
        tma_load_2d(sB + stage * Cfg::B_BYTES, &tmB, &full[stage], kb * F8_BK, n_blk * BN);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1 && lane == 0) {
    constexpr uint32_t idesc = make_idesc(UMMA_E4M3, UMMA_E4M3, F8_BM, BN, 0, 0);
    int stage = 0; uint32_t phase = 0;
    int acc = 0; uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        const uint64_t adesc = make_smem_desc_sw128(smem_u32(sA + stage * Cfg::A_BYTES), 0, 1024);
        const uint64_t bdesc = make_smem_desc_sw128(smem_u32(sB + stage * Cfg::B_BYTES), 0, 1024);
#pragma unroll
        for (int k = 0; k < F8_BK / 32; ++k) umma_f8_ss(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, k != 0);
        umma_commit(&empty[stage]);
        umma_commit(&tfull[acc]);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
        if (++acc == NACC) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    const int q = warp - 4;
    int acc = 0; uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_blk = tile % m_tiles, n_blk = tile / m_tiles;
      const int row = m_blk * F8_BM + q * 32 + lane;
      const int row_c = min(row, M - 1);
      float sum[BN];
#pragma unroll
      for (int i = 0; i < BN; ++i) sum[i] = 0.f;
      const float* xs_row = xs + (size_t)row_c * xs_stride;
      float a_scale = __ldg(xs_row);
      for (int kb = 0; kb < num_kb; ++kb) {
        const float a_next = (kb + 1 < num_kb) ? __ldg(xs_row + kb + 1) : 0.f;
        mbar_wait(&tfull[acc], acc_phase);
        tc_fence_after();
        const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN;
#pragma unroll
        for (int c = 0; c < BN / 32; ++c) {
          // weight-block scale of these 32 columns (a 128-column block never straddles a 32-column chunk)
          const int nb = min(n_blk * BN + c * 32, N - 1) >> 7;
          const float s = a_scale * __ldg(ws + (size_t)nb * ws_stride + kb);
          uint32_t r[32];
          tmem_ld_32x32b_x32(t_row + c * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) sum[c * 32 + i] = fmaf(s, __uint_as_float(r[i]), sum[c * 32 + i]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[acc]);
        if (++acc == NACC) { acc = 0; acc_phase ^= 1; }
        a_scale = a_next;
      }
      if (row < M) {
#pragma unroll
        for (int v = 0; v < BN / 8; ++v) {
          const int col = n_blk * BN + v * 8;
          if (col < N) {
            const U4 o = U4{pack_bf16(sum[v * 8], sum[v * 8 + 1]), pack_bf16(sum[v * 8 + 2], sum[v * 8 + 3]),
                            pack_bf16(sum[v * 8 + 4], sum[v * 8 + 5]), pack_bf16(sum[v * 8 + 6], sum[v * 8 + 7])};
            st128(C + (size_t)row * ldc + col, o);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

template <int BN>
static void launch_fp8(const void* A, const float* xs, const void* W, const float* ws, void* C, int M, int N, int K, int lda, int ldw,
                       int ldc, int xs_stride, int ws_stride, int num_sms, cudaStream_t stream) {
  using Cfg = F8Cfg<BN>;
  uint64_t dA[2] = {(uint64_t)K, (uint64_t)M}, sAs[1] = {(uint64_t)lda};
  uint64_t dB[2] = {(uint64_t)K, (uint64_t)N}, sBs[1] = {(uint64_t)ldw};
  uint32_t bA[2] = {F8_BK, F8_BM}, bB[2] = {F8_BK, (uint32_t)BN};
  CUtensorMap tmA = make_tmap(A, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, dA, sAs, bA, CU_TENSOR_MAP_SWIZZLE_128B);
  CUtensorMap tmB = make_tmap(W, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, dB, sBs, bB, CU_TENSOR_MAP_SWIZZLE_128B);
  auto kern = gemm_fp8_block_kernel<BN>;
  static bool configured = false;
  if (!configured) {
    SGLB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    configured = true;
  }
  const int m_tiles = (M + F8_BM - 1) / F8_BM, n_tiles = (N + BN - 1) / BN;
  const int grid = (int)std::min<long long>((long long)m_tiles * n_tiles, num_sms);
  kern<<<grid, F8_THREADS, Cfg::SMEM_BYTES, stream>>>(tmA, tmB, xs, ws, (__nv_bfloat16*)C, M, N, K, ldc, xs_stride, ws_stride);
  SGLB_CUDA_CHECK(cudaGetLastError());
}

void gemm_fp8_block(const void* A, const float* xs, const void* W, const float* ws, void* C, int M, int N, int K, int lda, int ldw,
                    int ldc, int xs_stride, int ws_stride, int num_sms, cudaStream_t stream) {
  if (M <= 0 || N <= 0) return;
  if (K % 16 || N % 8 || lda % 16 || ldw % 16) throw std::runtime_error("sglb fp8 gemm: K, lda, ldw must be multiples of 16 and N of 8");
  const int m_tiles = (M + F8_BM - 1) / F8_BM;
  // small M is weight-bandwidth bound: narrower tiles -> more CTAs streaming W
  int bn = 128;
  if (m_tiles * ((N + 127) / 128) < num_sms) bn = 64;
  if (m_tiles * ((N + 63) / 64) < num_sms) bn = 32;
  switch (bn) {
    case 128: launch_fp8<128>(A, xs, W, ws, C, M, N, K, lda, ldw, ldc, xs_stride, ws_stride, num_sms, stream); break;
    case 64: launch_fp8<64>(A, xs, W, ws, C, M, N, K, lda, ldw, ldc, xs_stride, ws_stride, num_sms, stream); break;
    default: launch_fp8<32>(A, xs, W, ws, C, M, N, K, lda, ldw, ldc, xs_stride, ws_stride, num_sms, stream); break;
  }
}

// ---------------------------------------------------------------- activation quantiser: one warp per (row, 128-group)
__global__ void __launch_bounds__(256) per_token_group_quant_fp8_kernel(const __nv_bfloat16* __restrict__ x, uint8_t* __restrict__ q,
                                                                        float* __restrict__ scales, int M, int K, int x_stride,
                                                                        int groups, int ue8m0) {
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (gw >= M * groups) return;
  const int row = gw / groups, g = gw % groups;
  const int col = g * 128 + lane * 4;
  float v[4] = {0.f, 0.f, 0.f, 0.f};
  if (col + 3 < K) {
    const uint2 raw = *reinterpret_cast<const uint2*>(x + (size_t)row * x_stride + col);
    v[0] = bf16_lo(raw.x); v[1] = bf16_hi(raw.x); v[2] = bf16_lo(raw.y); v[3] = bf16_hi(raw.y);
  } else {
    for (int i = 0; i < 4; ++i)
      if (col + i < K) v[i] = __bfloat162float(x[(size_t)row * x_stride + col + i]);
  }
  float amax = fmaxf(fmaxf(fabsf(v[0]), fabsf(v[1])), fmaxf(fabsf(v[2]), fabsf(v[3])));
  amax = warp_max(amax);
  float scale = fmaxf(amax, 1e-10f) / 448.f;
  if (ue8m0) scale = exp2f(ceilf(log2f(scale)));
  const float inv = 1.f / scale;
  if (lane == 0) scales[(size_t)row * groups + g] = scale;
  uint32_t packed = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __nv_fp8_storage_t b = __nv_cvt_float_to_fp8(v[i] * inv, __NV_SATFINITE, __NV_E4M3);
    packed |= (uint32_t)b << (8 * i);
  }
  if (col + 3 < K) {
    *reinterpret_cast<uint32_t*>(q + (size_t)row * K + col) = packed;
  } else {
    for (int i = 0; i < 4; ++i)
      if (col + i < K) q[(size_t)row * K + col + i] = (uint8_t)(packed >> (8 * i));
  }
}

void per_token_group_quant_fp8(const void* x, void* q, float* scales, int M, int K, int x_stride, bool ue8m0, cudaStream_t stream) {
  if (M <= 0) return;
  if (K % 4) throw std::runtime_error("sglb fp8 quant: K must be a multiple of 4");
  const int groups = (K + 127) / 128;
  const long long warps = (long long)M * groups;
  const int grid = (int)((warps * 32 + 255) / 256);
  per_token_group_quant_fp8_kernel<<<grid, 256, 0, stream>>>((const __nv_bfloat16*)x, (uint8_t*)q, scales, M, K, x_stride, groups, ue8m0);
  SGLB_CUDA_CHECK(cudaGetLastError());
}

}  // namespace sglb
