// Decode-shape ("skinny") bf16 GEMM for sm_100a:  C[M,N] = A[M,K] * W[N,K]^T  with M <= 256 tokens.
//
// For M this small the classic orientation wastes the machine twice: the 128-row A box is mostly zero padding (smem and
// tcgen05 rows that carry no data) and every CTA re-reads all of A, so a weight-streaming GEMM ends up latency-bound at a
// fraction of HBM bandwidth (profiles/microbench_r1.md: cuBLAS and our wide kernel both sit at 25-30 us for 33-50 MB of
// weights).  This kernel swaps the operands and splits K:
//
//   * W is the UMMA *M* operand: each CTA owns 128 output features and streams ONLY useful bytes (16 KB of W per 64-wide K
//     block), with up to 12 TMA stages in flight (Little's law: ~190 KB of weights in flight per SM);
//   * the tokens are the UMMA *N* operand (N = MT in {16,32,64,128,256}), accumulator D^T[128 features, MT tokens] in TMEM;
//   * grid = (N/128 feature tiles) x (K splits) so that ~148-296 CTAs stream disjoint slices of W; split partials land in
//     bf16 "slabs" [split][M][N] that the CONSUMER kernel sums (fused add+RMSNorm, RoPE+KV-write) — no reduction pass;
//   * the epilogue transposes through shared memory so global stores are 16-byte, row-contiguous — locally, or into every
//     tensor-parallel peer's NVLink arena (fused GEMM + all-reduce producer; slab index = rank*splits + split);
//   * SPLIT_SILU: the 128 W rows of a tile are 64 gate rows + the matching 64 up rows; the write-out computes silu(g)*u, so
//     gate_up_proj + SiluAndMul is one kernel with the checkpoint's un-interleaved weight layout.
//
// Reference counterparts: cuBLASLt / CuTe-DSL bf16 GEMM (layers/quantization/unquant.py:72-169), K3 GEMM+AR (tcgen05,
// layers/k3_gemm_ar.py), silu_and_mul (kernels/aot/csrc/elementwise/activation.cu:98).
#include "common.cuh"
#include "tma_host.h"

#include <algorithm>

namespace sglb {

constexpr int SK_BF = 128;   // output features per CTA (UMMA M)
constexpr int SK_BK = 64;
constexpr int SK_THREADS = 256;

struct SkinnyPush {
  __nv_bfloat16* peer_slot[8];
  uint32_t* peer_flags[8];
  const uint32_t* epoch_ptr;
  uint32_t* done_ctr;
  int rank, world;       // world == 0 -> local output only
};

template <int MT>
struct SkCfg {
  static constexpr int W_BYTES = SK_BF * SK_BK * 2;
  static constexpr int A_BYTES = MT * SK_BK * 2;
  static constexpr int STAGE_BYTES = W_BYTES + A_BYTES;
  static constexpr int OUT_BYTES = MT * SK_BF * 2;                     // transposed staging [MT][128] bf16 (aliases the stages)
  static constexpr int RAW_STAGES = (200 * 1024) / STAGE_BYTES;
  static constexpr int STAGES = RAW_STAGES > 12 ? 12 : RAW_STAGES;
  static constexpr int TMEM_COLS = MT < 32 ? 32 : MT;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
  static_assert(OUT_BYTES <= STAGES * STAGE_BYTES, "staging must fit in the pipeline buffers");
};

enum { SK_EPI_NONE = 0, SK_EPI_BIAS = 1, SK_EPI_SILU_SPLIT = 4 };

template <int MT, int EPI>
__global__ void __launch_bounds__(SK_THREADS, 1)
gemm_skinny_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmA, __nv_bfloat16* __restrict__ C,
                   const __nv_bfloat16* __restrict__ bias, int M, int N, int K, int ldc, int kb_per_split, long long slab_stride,
                   int n_half, const SkinnyPush push) {
  using Cfg = SkCfg<MT>;
  constexpr int STAGES = Cfg::STAGES;
  constexpr bool SILU = (EPI == SK_EPI_SILU_SPLIT);
  constexpr int F_OUT = SILU ? SK_BF / 2 : SK_BF;      // output features written per tile
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sW = smem;
  uint8_t* sA = smem + STAGES * Cfg::W_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* tfull = bars + 2 * STAGES;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_blk = blockIdx.x, ks = blockIdx.y;
  const int num_kb = (K + SK_BK - 1) / SK_BK;
  const int kb0 = ks * kb_per_split, kb1 = min(num_kb, kb0 + kb_per_split);
  const int f0 = n_blk * F_OUT;                        // first output feature of this tile

  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmW); tma_prefetch_desc(&tmA); }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(tfull, 1);
    fence_barrier_init();
  }
  if (warp == 2) { tmem_alloc(tmem_ptr, Cfg::TMEM_COLS); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0 && lane == 0) {
    int stage = 0; uint32_t phase = 0;
    for (int kb = kb0; kb < kb1; ++kb) {
      mbar_wait(&empty[stage], phase ^ 1);
      mbar_expect_tx(&full[stage], Cfg::STAGE_BYTES);
      uint8_t* w_dst = sW + stage * Cfg::W_BYTES;
      if constexpr (SILU) {
        tma_load_2d(w_dst, &tmW, &full[stage], kb * SK_BK, f0);
        tma_load_2d(w_dst + Cfg::W_BYTES / 2, &tmW, &full[stage], kb * SK_BK, n_half + f0);
      } else {
        tma_load_2d(w_dst, &tmW, &full[stage], kb * SK_BK, f0);
      }
      tma_load_2d(sA + stage * Cfg::A_BYTES, &tmA, &full[stage], kb * SK_BK, 0);
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1 && lane == 0) {
    constexpr uint32_t idesc = make_idesc(UMMA_BF16, UMMA_BF16, SK_BF, MT, 0, 0);
    int stage = 0; uint32_t phase = 0;
    for (int kb = kb0; kb < kb1; ++kb) {
      mbar_wait(&full[stage], phase);
      tc_fence_after();
      const uint64_t wdesc = make_smem_desc_sw128(smem_u32(sW + stage * Cfg::W_BYTES), 0, 1024);
      const uint64_t adesc = make_smem_desc_sw128(smem_u32(sA + stage * Cfg::A_BYTES), 0, 1024);
#pragma unroll
      for (int k = 0; k < SK_BK / 16; ++k) umma_f16_ss(tmem_base, wdesc + 2 * k, adesc + 2 * k, idesc, (kb > kb0) || k != 0);
      umma_commit(&empty[stage]);
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }
    umma_commit(tfull);
  } else if (warp >= 4) {
    // ------------------------------------------------------------ epilogue: TMEM -> (bias) -> bf16 -> smem^T -> 16-B global stores
    const int q = warp - 4;
    const int feat = q * 32 + lane;                     // TMEM lane = feature row of this tile
    __nv_bfloat16* sOut = reinterpret_cast<__nv_bfloat16*>(smem);   // [MT][128]; the pipeline stages are drained by now
    mbar_wait(tfull, 0);
    tc_fence_after();
    float b = 0.f;
    if (EPI == SK_EPI_BIAS && ks == 0 && f0 + feat < N) b = __bfloat162float(bias[f0 + feat]);
    const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16);
    constexpr int CH = MT < 32 ? MT : 32;
#pragma unroll 1
    for (int c = 0; c < MT / CH; ++c) {
      uint32_t r[32];
      if constexpr (CH == 32) {
        tmem_ld_32x32b_x32(t_row + c * 32, r);
      } else {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                       "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                     : "r"(t_row) : "memory");
      }
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < CH; ++i) sOut[(c * CH + i) * SK_BF + feat] = __float2bfloat16(__uint_as_float(r[i]) + b);
    }
    tc_fence_before();
    asm volatile("bar.sync 1, 128;" ::: "memory");      // the four epilogue warps only
    // write-out: token rows of F_OUT contiguous features, 16 bytes per thread per step
    const int et = threadIdx.x - 128;                   // 0..127
    constexpr int VEC_PER_ROW = F_OUT / 8;              // 16 (plain) or 8 (silu)
    const size_t slab = (size_t)((push.world ? push.rank * (int)gridDim.y : 0) + ks) * (size_t)slab_stride;
    for (int idx = et; idx < MT * VEC_PER_ROW; idx += 128) {
      const int tok = idx / VEC_PER_ROW, v = idx % VEC_PER_ROW;
      if (tok >= M) break;
      const int col = f0 + v * 8;
      if (col >= N) continue;
      U4 o;
      if constexpr (SILU) {
        const U4 g = *reinterpret_cast<const U4*>(sOut + tok * SK_BF + v * 8);
        const U4 u = *reinterpret_cast<const U4*>(sOut + tok * SK_BF + 64 + v * 8);
        auto f = [](uint32_t gg, uint32_t uu) {
          const float g0 = bf16_lo(gg), g1 = bf16_hi(gg);
          return pack_bf16(g0 / (1.f + __expf(-g0)) * bf16_lo(uu), g1 / (1.f + __expf(-g1)) * bf16_hi(uu));
        };
        o = U4{f(g.x, u.x), f(g.y, u.y), f(g.z, u.z), f(g.w, u.w)};
      } else {
        o = *reinterpret_cast<const U4*>(sOut + tok * SK_BF + v * 8);
      }
      const size_t off = slab + (size_t)tok * ldc + col;
      if (push.world) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
          if (r < push.world) st128(push.peer_slot[r] + off, o);
      } else {
        st128(C + off, o);
      }
    }
    if (push.world) __threadfence_system();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  if (push.world && threadIdx.x == 0) {
    __threadfence();
    const uint32_t t = atomicAdd(push.done_ctr, 1u);
    if (t == gridDim.x * gridDim.y - 1) {               // last CTA of this rank: every slab tile is out -> publish the epoch
      *push.done_ctr = 0;
      __threadfence_system();
      const uint32_t epoch = *((volatile const uint32_t*)push.epoch_ptr) + 1;
      for (int r = 0; r < push.world; ++r)
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(push.peer_flags[r] + push.rank), "r"(epoch) : "memory");
    }
  }
}

template <int MT, int EPI>
static void launch_skinny(const void* A, const void* W, void* C, const void* bias, int M, int N, int K, int lda, int ldw, int ldc,
                          int splits, long long slab_stride, int n_half, const SkinnyPush& push, cudaStream_t stream) {
  using Cfg = SkCfg<MT>;
  constexpr bool SILU = (EPI == SK_EPI_SILU_SPLIT);
  const int w_rows = SILU ? 2 * n_half : N;
  uint64_t dW[2] = {(uint64_t)K, (uint64_t)w_rows}, sWs[1] = {(uint64_t)ldw * 2};
  uint64_t dA[2] = {(uint64_t)K, (uint64_t)M}, sAs[1] = {(uint64_t)lda * 2};
  uint32_t bW[2] = {SK_BK, (uint32_t)(SILU ? SK_BF / 2 : SK_BF)}, bA[2] = {SK_BK, (uint32_t)MT};
  CUtensorMap tmW = make_tmap(W, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dW, sWs, bW, CU_TENSOR_MAP_SWIZZLE_128B);
  CUtensorMap tmA = make_tmap(A, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dA, sAs, bA, CU_TENSOR_MAP_SWIZZLE_128B);
  auto kern = gemm_skinny_kernel<MT, EPI>;
  static bool configured = false;
  if (!configured) {
    SGLB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    configured = true;
  }
  const int f_out = SILU ? SK_BF / 2 : SK_BF;
  const int n_tiles = (N + f_out - 1) / f_out;
  const int num_kb = (K + SK_BK - 1) / SK_BK;
  const int kb_per = (num_kb + splits - 1) / splits;
  const int eff_splits = (num_kb + kb_per - 1) / kb_per;
  if (eff_splits != splits) throw std::runtime_error("sglb skinny gemm: split count is not normalised (use gemm_skinny_norm_splits)");
  dim3 grid(n_tiles, eff_splits);
  kern<<<grid, SK_THREADS, Cfg::SMEM_BYTES, stream>>>(tmW, tmA, (__nv_bfloat16*)C, (const __nv_bfloat16*)bias, M, N, K, ldc, kb_per,
                                                      slab_stride, n_half, push);
  SGLB_CUDA_CHECK(cudaGetLastError());
}

template <int EPI>
static void dispatch_mt(int mt, const void* A, const void* W, void* C, const void* bias, int M, int N, int K, int lda, int ldw, int ldc,
                        int splits, long long slab_stride, int n_half, const SkinnyPush& push, cudaStream_t stream) {
  switch (mt) {
    case 16: launch_skinny<16, EPI>(A, W, C, bias, M, N, K, lda, ldw, ldc, splits, slab_stride, n_half, push, stream); break;
    case 32: launch_skinny<32, EPI>(A, W, C, bias, M, N, K, lda, ldw, ldc, splits, slab_stride, n_half, push, stream); break;
    case 64: launch_skinny<64, EPI>(A, W, C, bias, M, N, K, lda, ldw, ldc, splits, slab_stride, n_half, push, stream); break;
    case 128: launch_skinny<128, EPI>(A, W, C, bias, M, N, K, lda, ldw, ldc, splits, slab_stride, n_half, push, stream); break;
    default: launch_skinny<256, EPI>(A, W, C, bias, M, N, K, lda, ldw, ldc, splits, slab_stride, n_half, push, stream); break;
  }
}

static int pick_mt(int M) { return M <= 16 ? 16 : M <= 32 ? 32 : M <= 64 ? 64 : M <= 128 ? 128 : 256; }

// largest s' <= s with ceil(nkb / ceil(nkb / s')) == s'  (every split non-empty; the grid's y extent equals the slab count)
int gemm_skinny_norm_splits(int K, int s) {
  const int num_kb = (K + SK_BK - 1) / SK_BK;
  s = std::max(1, std::min(s, num_kb));
  for (;;) {
    const int kb_per = (num_kb + s - 1) / s;
    const int s2 = (num_kb + kb_per - 1) / kb_per;
    if (s2 == s) return s;
    s = s2;
  }
}

// how many K splits to use for this problem (callers size their slab workspace with it)
int gemm_skinny_splits(int M, int N, int K, int max_splits, int num_sms) {
  const int n_tiles = (N + SK_BF - 1) / SK_BF;
  const int num_kb = (K + SK_BK - 1) / SK_BK;
  // exactly ONE wave: n_tiles * s <= num_sms (a partial second wave doubles the kernel time — measured 160 CTAs on 148 SMs)
  int s = std::max(1, std::min(max_splits, num_sms / std::max(1, n_tiles)));
  s = std::min(s, std::max(1, num_kb / 8));                                      // keep >= 8 K blocks per split
  return gemm_skinny_norm_splits(K, s);
}

// epilogue: 0 none, 1 bias (splits must be 1 or the bias is added by split 0), 4 split-silu (W = [gate; up], N = n_half)
void gemm_bf16_skinny(const void* A, const void* W, void* C, const void* bias, int M, int N, int K, int lda, int ldw, int ldc,
                      int epilogue, int splits, long long slab_stride, cudaStream_t stream) {
  if (M <= 0 || N <= 0) return;
  if (M > 256) throw std::runtime_error("sglb skinny gemm: M must be <= 256");
  if (K % 8 || lda % 8 || ldw % 8 || ldc % 8) throw std::runtime_error("sglb skinny gemm: K, lda, ldw, ldc must be multiples of 8");
  const int mt = pick_mt(M);
  SkinnyPush none{};
  switch (epilogue) {
    case SK_EPI_NONE: dispatch_mt<SK_EPI_NONE>(mt, A, W, C, bias, M, N, K, lda, ldw, ldc, splits, slab_stride, 0, none, stream); break;
    case SK_EPI_BIAS: dispatch_mt<SK_EPI_BIAS>(mt, A, W, C, bias, M, N, K, lda, ldw, ldc, splits, slab_stride, 0, none, stream); break;
    case SK_EPI_SILU_SPLIT:
      if (N % 8) throw std::runtime_error("sglb skinny gemm: intermediate size must be a multiple of 8");
      dispatch_mt<SK_EPI_SILU_SPLIT>(mt, A, W, C, bias, M, N, K, lda, ldw, ldc, 1, 0, N, none, stream);
      break;
    default: throw std::runtime_error("sglb skinny gemm: bad epilogue");
  }
}

// fused GEMM + all-reduce producer for decode shapes: slabs [world*splits][M][N] in every peer's arena
void gemm_bf16_skinny_push(const void* A, const void* W, int M, int N, int K, int lda, int ldw, int splits, const int64_t* peer_slot_ptrs,
                           const int64_t* peer_flag_ptrs, int rank, int world, const uint32_t* epoch_ptr, uint32_t* done_ctr,
                           cudaStream_t stream) {
  if (M <= 0 || N <= 0) return;
  if (M > 256 || world > 8 || world < 1) throw std::runtime_error("sglb skinny gemm push: need M <= 256 and 1 <= world <= 8");
  SkinnyPush p{};
  for (int r = 0; r < world; ++r) {
    p.peer_slot[r] = reinterpret_cast<__nv_bfloat16*>(peer_slot_ptrs[r]);
    p.peer_flags[r] = reinterpret_cast<uint32_t*>(peer_flag_ptrs[r]);
  }
  p.epoch_ptr = epoch_ptr; p.done_ctr = done_ctr; p.rank = rank; p.world = world;
  dispatch_mt<SK_EPI_NONE>(pick_mt(M), A, W, nullptr, nullptr, M, N, K, lda, ldw, N, splits, (long long)M * N, 0, p, stream);
}

}  // namespace sglb
