// Persistent, warp-specialised tcgen05 GEMM for sm_100a:   C[M,N] = A[M,K] * W[N,K]^T  (+ epilogue)
//
//   warp 0      : TMA producer  (A and W tiles, 128-B swizzle, multi-stage mbarrier ring)
//   warp 1      : MMA issuer    (one thread; tcgen05.mma cta_group::1 kind::f16, fp32 accum in TMEM)
//   warp 2      : TMEM allocator (2 accumulator buffers so tile i's epilogue overlaps tile i+1's mainloop)
//   warps 4..7  : epilogue      (tcgen05.ld -> bias / SiLU*mul -> bf16 -> 128-bit global stores)
//
// Stands in for the reference's cuBLASLt / CuTe-DSL bf16 GEMM (python/sglang/srt/layers/quantization/
// unquant.py:72-169, kernels/ops/gemm/cutedsl_bf16_gemm.py) and its separate silu_and_mul kernel
// (kernels/aot/csrc/elementwise/activation.cu:98) when EPI_SILU_MUL is fused.
#include "common.cuh"
#include "tma_host.h"

#include <algorithm>

namespace sglb {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
constexpr int GEMM_THREADS = 256;

enum GemmEpilogue { EPI_NONE = 0, EPI_BIAS = 1, EPI_SILU_MUL = 2, EPI_PUSH = 3, EPI_SILU_SPLIT = 4, EPI_ROWSCALE = 5 };
// EPI_SILU_SPLIT: W = [gate rows (n_half) ; up rows (n_half)] un-interleaved (the checkpoint layout).  Each B stage is
//   filled by TWO TMA boxes of BN/2 rows — gate rows j.. and up rows n_half+j.. — so one UMMA of width BN produces the
//   gate half in accumulator columns [0,BN/2) and the matching up half in [BN/2,BN); the epilogue writes silu(g)*u.
// EPI_ROWSCALE : C[row,:] = acc[row,:] * scale[row]  (fp32 per-row scale passed through `bias`; MoE routing weights).
// GROUPED      : A rows are expert-sorted and padded to 128-row tiles; W is [E][N][K] behind a 3-D tensor map and
//   tile_expert[m_tile] selects the expert slab; the number of live M tiles is read from device memory.

// EPI_PUSH: fused GEMM + all-reduce producer.  The epilogue stores every output tile into slot [rank] of EVERY peer's
// symmetric arena (P2P stores over NVLink / NVSwitch, peer pointers mapped into this address space), tile by tile while
// the tensor core already works on the next tile; the last CTA to finish publishes `epoch` to all peers with a
// system-scope release store.  The consumer (nvl_local_reduce_add_rmsnorm_kernel) then only reads LOCAL memory.
struct GemmPush {
  __nv_bfloat16* peer_slot[8];   // peer r's arena slot base for this call: layout [world][M][ldc]
  uint32_t* peer_flags[8];       // peer r's flag array for this channel: flags[src_rank]
  const uint32_t* epoch_ptr;     // device-resident epoch (the consumer kernel advances it)
  uint32_t* done_ctr;            // local CTA-completion counter
  long long slot_stride;         // elements between two source ranks' regions (= M * ldc)
  int rank, world;
};

template <int BN>
struct GemmCfg {
  static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;
  static constexpr int B_BYTES = BN * GEMM_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN >= 256) ? 4 : (BN >= 128 ? 6 : 8);
  static constexpr int TMEM_COLS = (2 * BN < 32) ? 32 : 2 * BN;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

template <int BN, int EPI, bool GROUPED>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               __nv_bfloat16* __restrict__ C, const __nv_bfloat16* __restrict__ bias, int M, int N, int K, int ldc,
               const GemmPush push, const int* __restrict__ tile_expert, const int* __restrict__ num_m_tiles_ptr, int n_half) {
  constexpr bool SPLIT = (EPI == EPI_SILU_SPLIT);
  constexpr int BN_OUT = SPLIT ? BN / 2 : BN;       // output columns covered by one tile
  using Cfg = GemmCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * Cfg::A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* tfull = bars + 2 * STAGES;
  uint64_t* tempty = bars + 2 * STAGES + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m_tiles = GROUPED ? min(__ldg(num_m_tiles_ptr), (M + GEMM_BM - 1) / GEMM_BM) : (M + GEMM_BM - 1) / GEMM_BM;
  const int n_tiles = (N + BN_OUT - 1) / BN_OUT;
  const int num_tiles = m_tiles * n_tiles;
  const int num_kb = (K + GEMM_BK - 1) / GEMM_BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 4); }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_ptr, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0 && lane == 0) {
    // ------------------------------------------------------------ TMA producer
    int stage = 0; uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_blk = tile % m_tiles, n_blk = tile / m_tiles;
      const int expert = GROUPED ? __ldg(tile_expert + m_blk) : 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&empty[stage], phase ^ 1);
        mbar_expect_tx(&full[stage], Cfg::STAGE_BYTES);
        tma_load_2d(sA + stage * Cfg::A_BYTES, &tmA, &full[stage], kb * GEMM_BK, m_blk * GEMM_BM);
        uint8_t* b_dst = sB + stage * Cfg::B_BYTES;
        if constexpr (GROUPED) {
          tma_load_3d(b_dst, &tmB, &full[stage], kb * GEMM_BK, n_blk * BN_OUT, expert);
          if constexpr (SPLIT) tma_load_3d(b_dst + Cfg::B_BYTES / 2, &tmB, &full[stage], kb * GEMM_BK, n_half + n_blk * BN_OUT, expert);
        } else {
          tma_load_2d(b_dst, &tmB, &full[stage], kb * GEMM_BK, n_blk * BN_OUT);
          if constexpr (SPLIT) tma_load_2d(b_dst + Cfg::B_BYTES / 2, &tmB, &full[stage], kb * GEMM_BK, n_half + n_blk * BN_OUT);
        }
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = make_idesc(UMMA_BF16, UMMA_BF16, GEMM_BM, BN, 0, 0);
    int stage = 0; uint32_t phase = 0;
    int acc = 0; uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      mbar_wait(&tempty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint64_t adesc = make_smem_desc_sw128(smem_u32(sA + stage * Cfg::A_BYTES), 0, 1024);
        const uint64_t bdesc = make_smem_desc_sw128(smem_u32(sB + stage * Cfg::B_BYTES), 0, 1024);
#pragma unroll
        for (int k = 0; k < GEMM_BK / 16; ++k) {
          // +32 B per K=16 step inside the 128-B swizzled row (start-address field is in 16-B units)
          umma_f16_ss(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
        }
        umma_commit(&empty[stage]);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      umma_commit(&tfull[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------ epilogue
    const int q = warp - 4;  // == warp % 4 : TMEM lane quadrant this warp may access
    int acc = 0; uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_blk = tile % m_tiles, n_blk = tile / m_tiles;
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const int row = m_blk * GEMM_BM + q * 32 + lane;
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN;
      float row_scale = 1.f;
      if (EPI == EPI_ROWSCALE && row < M) row_scale = __ldg(reinterpret_cast<const float*>(bias) + row);
#pragma unroll 1
      for (int c = 0; c < BN_OUT / 32; ++c) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(t_row + c * 32, r);
        if constexpr (SPLIT) {
          uint32_t u[32];
          tmem_ld_32x32b_x32(t_row + BN / 2 + c * 32, u);
          tmem_ld_wait();
          const int col0 = n_blk * BN_OUT + c * 32;
          if (row < M) {
#pragma unroll
            for (int v = 0; v < 4; ++v) {
              if (col0 + v * 8 < N) {
                uint32_t o[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const float g0 = __uint_as_float(r[v * 8 + 2 * i]), g1 = __uint_as_float(r[v * 8 + 2 * i + 1]);
                  const float u0 = __uint_as_float(u[v * 8 + 2 * i]), u1 = __uint_as_float(u[v * 8 + 2 * i + 1]);
                  o[i] = pack_bf16(g0 / (1.f + __expf(-g0)) * u0, g1 / (1.f + __expf(-g1)) * u1);
                }
                st128(C + (size_t)row * ldc + col0 + v * 8, U4{o[0], o[1], o[2], o[3]});
              }
            }
          }
          continue;
        }
        tmem_ld_wait();
        const int col0 = n_blk * BN + c * 32;
        if (EPI == EPI_SILU_MUL) {
          // columns [0,16) = gate, [16,32) = up of the same 16 intermediate features
          if (row < M && col0 < N) {
            uint32_t o[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              float g0 = __uint_as_float(r[2 * i]), g1 = __uint_as_float(r[2 * i + 1]);
              float u0 = __uint_as_float(r[16 + 2 * i]), u1 = __uint_as_float(r[16 + 2 * i + 1]);
              float a0 = g0 / (1.f + __expf(-g0)) * u0;
              float a1 = g1 / (1.f + __expf(-g1)) * u1;
              o[i] = pack_bf16(a0, a1);
            }
            __nv_bfloat16* dst = C + (size_t)row * ldc + (col0 >> 1);
            st128(dst, U4{o[0], o[1], o[2], o[3]});
            st128(dst + 8, U4{o[4], o[5], o[6], o[7]});
          }
        } else {
          if (row < M) {
#pragma unroll
            for (int v = 0; v < 4; ++v) {
              const int col = col0 + v * 8;
              if (col < N) {
                float f[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(r[v * 8 + i]);
                if (EPI == EPI_ROWSCALE) {
#pragma unroll
                  for (int i = 0; i < 8; ++i) f[i] *= row_scale;
                }
                if (EPI == EPI_BIAS) {
                  U4 b = ldg128(bias + col);
                  f[0] += bf16_lo(b.x); f[1] += bf16_hi(b.x); f[2] += bf16_lo(b.y); f[3] += bf16_hi(b.y);
                  f[4] += bf16_lo(b.z); f[5] += bf16_hi(b.z); f[6] += bf16_lo(b.w); f[7] += bf16_hi(b.w);
                }
                const U4 o = U4{pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7])};
                if (EPI == EPI_PUSH) {
                  const size_t off = (size_t)push.rank * push.slot_stride + (size_t)row * ldc + col;
#pragma unroll
                  for (int r = 0; r < 8; ++r)
                    if (r < push.world) st128(push.peer_slot[r] + off, o);
                } else {
                  st128(C + (size_t)row * ldc + col, o);
                }
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  if (EPI == EPI_PUSH) {
    // every thread that stored to peers makes its stores visible system-wide before the CTA signs off
    if (warp >= 4) __threadfence_system();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  if (EPI == EPI_PUSH && threadIdx.x == 0) {
    __threadfence();
    const uint32_t t = atomicAdd(push.done_ctr, 1u);
    if (t == gridDim.x - 1) {           // last CTA of this rank: all tiles of all CTAs are out -> publish
      *push.done_ctr = 0;
      __threadfence_system();
      const uint32_t epoch = *((volatile const uint32_t*)push.epoch_ptr) + 1;
      for (int r = 0; r < push.world; ++r)
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(push.peer_flags[r] + push.rank), "r"(epoch) : "memory");
    }
  }
}

// host-side extras: expert-grouped B ([E][rows][K], 3-D tensor map) and the split gate/up epilogue
struct GemmExtra {
  int num_experts = 0;                 // > 0 -> GROUPED
  const int* tile_expert = nullptr;    // [max m tiles]
  const int* num_m_tiles = nullptr;    // device scalar
  int w_rows = 0;                      // rows of W per expert (N for plain epilogues, 2*n_half for EPI_SILU_SPLIT)
  int n_half = 0;                      // EPI_SILU_SPLIT: row offset of the `up` half
};

template <int BN, int EPI, bool GROUPED>
static void launch_gemm(const void* A, const void* W, void* C, const void* bias, int M, int N, int K, int lda, int ldw,
                        int ldc, int num_sms, cudaStream_t stream, const GemmPush& push, const GemmExtra& ex) {
  using Cfg = GemmCfg<BN>;
  constexpr bool SPLIT = (EPI == EPI_SILU_SPLIT);
  constexpr int BN_OUT = SPLIT ? BN / 2 : BN;
  if (SPLIT && BN < 64) throw std::runtime_error("sglb gemm: split-silu epilogue needs BN >= 64");
  const int w_rows = ex.w_rows > 0 ? ex.w_rows : N;
  uint64_t dA[2] = {(uint64_t)K, (uint64_t)M}, sAs[1] = {(uint64_t)lda * 2};
  uint32_t bA[2] = {GEMM_BK, GEMM_BM};
  CUtensorMap tmA = make_tmap(A, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dA, sAs, bA, CU_TENSOR_MAP_SWIZZLE_128B);
  CUtensorMap tmB;
  if (GROUPED) {
    uint64_t dB[3] = {(uint64_t)K, (uint64_t)w_rows, (uint64_t)ex.num_experts};
    uint64_t sBs[2] = {(uint64_t)ldw * 2, (uint64_t)w_rows * ldw * 2};
    uint32_t bB[3] = {GEMM_BK, (uint32_t)BN_OUT, 1};
    tmB = make_tmap(W, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, dB, sBs, bB, CU_TENSOR_MAP_SWIZZLE_128B);
  } else {
    uint64_t dB[2] = {(uint64_t)K, (uint64_t)w_rows}, sBs[1] = {(uint64_t)ldw * 2};
    uint32_t bB[2] = {GEMM_BK, (uint32_t)BN_OUT};
    tmB = make_tmap(W, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dB, sBs, bB, CU_TENSOR_MAP_SWIZZLE_128B);
  }
  auto kern = gemm_tc_kernel<BN, EPI, GROUPED>;
  static bool configured = false;
  if (!configured) {
    SGLB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    configured = true;
  }
  const int m_tiles = (M + GEMM_BM - 1) / GEMM_BM, n_tiles = (N + BN_OUT - 1) / BN_OUT;
  const int grid = (int)std::min<long long>((long long)m_tiles * n_tiles, num_sms);
  kern<<<grid, GEMM_THREADS, Cfg::SMEM_BYTES, stream>>>(tmA, tmB, (__nv_bfloat16*)C, (const __nv_bfloat16*)bias, M, N, K, ldc, push,
                                                        ex.tile_expert, ex.num_m_tiles, ex.n_half);
  SGLB_CUDA_CHECK(cudaGetLastError());
}

template <int EPI, bool GROUPED = false>
static void dispatch_bn(int bn, const void* A, const void* W, void* C, const void* bias, int M, int N, int K, int lda,
                        int ldw, int ldc, int num_sms, cudaStream_t stream, const GemmPush& push = GemmPush{},
                        const GemmExtra& ex = GemmExtra{}) {
  switch (bn) {
    case 32: launch_gemm<32, EPI, GROUPED>(A, W, C, bias, M, N, K, lda, ldw, ldc, num_sms, stream, push, ex); break;
    case 64: launch_gemm<64, EPI, GROUPED>(A, W, C, bias, M, N, K, lda, ldw, ldc, num_sms, stream, push, ex); break;
    case 128: launch_gemm<128, EPI, GROUPED>(A, W, C, bias, M, N, K, lda, ldw, ldc, num_sms, stream, push, ex); break;
    case 256: launch_gemm<256, EPI, GROUPED>(A, W, C, bias, M, N, K, lda, ldw, ldc, num_sms, stream, push, ex); break;
    default: throw std::runtime_error("sglb gemm: unsupported BN");
  }
}

// Heuristic tile-N: enough CTAs to keep HBM busy when weight-bandwidth bound (small M), wide tiles when
// compute bound (large M).  `bn <= 0` selects automatically.
int gemm_pick_bn(int M, int N, int K, int num_sms) {
  const int m_tiles = (M + GEMM_BM - 1) / GEMM_BM;
  if (m_tiles >= 8) return (N % 256 == 0) ? 256 : 128;
  int best = 128;
  for (int bn : {256, 128, 64, 32}) {
    if (N % bn) continue;
    best = bn;
    if (m_tiles * (N / bn) >= num_sms) break;
  }
  return best;
}

void gemm_bf16_tc(const void* A, const void* W, void* C, const void* bias, int M, int N, int K, int lda, int ldw,
                  int ldc, int epilogue, int bn, int num_sms, cudaStream_t stream) {
  if (M <= 0 || N <= 0) return;
  if (K % 8 || N % 8 || lda % 8 || ldw % 8) throw std::runtime_error("sglb gemm: K, N, lda, ldw must be multiples of 8");
  if (bn <= 0) bn = gemm_pick_bn(M, N, K, num_sms);
  switch (epilogue) {
    case EPI_NONE: dispatch_bn<EPI_NONE>(bn, A, W, C, bias, M, N, K, lda, ldw, ldc, num_sms, stream); break;
    case EPI_BIAS: dispatch_bn<EPI_BIAS>(bn, A, W, C, bias, M, N, K, lda, ldw, ldc, num_sms, stream); break;
    case EPI_SILU_MUL: dispatch_bn<EPI_SILU_MUL>(bn, A, W, C, bias, M, N, K, lda, ldw, ldc, num_sms, stream); break;
    case EPI_ROWSCALE: dispatch_bn<EPI_ROWSCALE>(bn, A, W, C, bias, M, N, K, lda, ldw, ldc, num_sms, stream); break;
    case EPI_SILU_SPLIT: {
      // W is [2*n_half, K] = [gate ; up]; C is [M, n_half].  N counts W rows on entry.
      const int n_half = N / 2;
      if (bn < 64 || n_half % (bn / 2)) {      // the epilogue walks 32-column chunks of each half: BN/2 >= 32
        bn = 0;
        const int m_tiles = (M + GEMM_BM - 1) / GEMM_BM;
        for (int cand : {256, 128, 64}) {
          if (n_half % (cand / 2)) continue;
          bn = cand;
          if ((long long)m_tiles * (n_half / (cand / 2)) >= num_sms) break;
        }
        if (!bn) throw std::runtime_error("sglb gemm: split-silu needs intermediate size % 32 == 0");
      }
      GemmExtra ex; ex.w_rows = N; ex.n_half = n_half;
      dispatch_bn<EPI_SILU_SPLIT>(bn, A, W, C, bias, M, n_half, K, lda, ldw, ldc, num_sms, stream, GemmPush{}, ex);
      break;
    }
    default: throw std::runtime_error("sglb gemm: bad epilogue");
  }
}

// Expert-grouped GEMM over expert-sorted, 128-row-padded activations (see moe.cu for the alignment kernels).
//   split_silu: W = [E][2*n_out][K] (gate;up), C = [M_cap][n_out] = silu(gate)*up
//   else      : W = [E][n_out][K],             C = [M_cap][n_out] = acc * row_scale[row]   (row_scale may be null)
void gemm_bf16_tc_grouped(const void* A, const void* W, void* C, const float* row_scale, int M_cap, int n_out, int K, int lda,
                          int ldc, int num_experts, const int* tile_expert, const int* num_m_tiles, bool split_silu, int num_sms,
                          cudaStream_t stream) {
  if (M_cap <= 0 || n_out <= 0) return;
  if (K % 8 || n_out % 32 || lda % 8) throw std::runtime_error("sglb grouped gemm: K % 8, n_out % 32, lda % 8 required");
  GemmExtra ex; ex.num_experts = num_experts; ex.tile_expert = tile_expert; ex.num_m_tiles = num_m_tiles;
  const int m_tiles = M_cap / GEMM_BM;
  int bn = 0;
  if (split_silu) {
    ex.w_rows = 2 * n_out; ex.n_half = n_out;
    for (int cand : {256, 128, 64}) {
      if (n_out % (cand / 2)) continue;
      bn = cand;
      if ((long long)m_tiles * (n_out / (cand / 2)) >= 2LL * num_sms) break;
    }
    if (!bn) throw std::runtime_error("sglb grouped gemm: bad intermediate size");
    dispatch_bn<EPI_SILU_SPLIT, true>(bn, A, W, C, nullptr, M_cap, n_out, K, lda, K, ldc, num_sms, stream, GemmPush{}, ex);
  } else {
    ex.w_rows = n_out;
    for (int cand : {256, 128, 64, 32}) {
      if (n_out % cand) continue;
      bn = cand;
      if ((long long)m_tiles * (n_out / cand) >= 2LL * num_sms) break;
    }
    if (!bn) throw std::runtime_error("sglb grouped gemm: n_out must be a multiple of 32");
    if (row_scale) dispatch_bn<EPI_ROWSCALE, true>(bn, A, W, C, row_scale, M_cap, n_out, K, lda, K, ldc, num_sms, stream, GemmPush{}, ex);
    else dispatch_bn<EPI_NONE, true>(bn, A, W, C, nullptr, M_cap, n_out, K, lda, K, ldc, num_sms, stream, GemmPush{}, ex);
  }
}

// Fused GEMM + all-reduce producer: C is not written locally; every tile goes to slot [rank] of every peer's arena.
void gemm_bf16_tc_push(const void* A, const void* W, int M, int N, int K, int lda, int ldw, const int64_t* peer_slot_ptrs,
                       const int64_t* peer_flag_ptrs, int rank, int world, const uint32_t* epoch_ptr, uint32_t* done_ctr, int bn,
                       int num_sms, cudaStream_t stream) {
  if (M <= 0 || N <= 0) return;
  if (K % 8 || N % 8 || lda % 8 || ldw % 8) throw std::runtime_error("sglb gemm push: K, N, lda, ldw must be multiples of 8");
  if (world > 8) throw std::runtime_error("sglb gemm push: world size > 8");
  GemmPush p{};
  for (int r = 0; r < world; ++r) {
    p.peer_slot[r] = reinterpret_cast<__nv_bfloat16*>(peer_slot_ptrs[r]);
    p.peer_flags[r] = reinterpret_cast<uint32_t*>(peer_flag_ptrs[r]);
  }
  p.epoch_ptr = epoch_ptr; p.done_ctr = done_ctr; p.slot_stride = (long long)M * N; p.rank = rank; p.world = world;
  if (bn <= 0) bn = gemm_pick_bn(M, N, K, num_sms);
  dispatch_bn<EPI_PUSH>(bn, A, W, nullptr, nullptr, M, N, K, lda, ldw, N, num_sms, stream, p);
}

}  // namespace sglb
