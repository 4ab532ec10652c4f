// Persistent stream-K weight-streaming GEMM for sm_100a decode shapes:   C[M,N] = A[M,K] * W[N,K]^T   (M <= 256 tokens)
//
// Why this kernel exists (profiles/gemm_decode_r1.md): at decode the four projections of a layer are pure weight streaming, and both the
// library and the first swap-AB kernel sit at 0.5-0.7 of the HBM roofline in-graph because their work decomposition is tile x wave:
//   * gate_up = 224 tiles on 148 SMs = 1.51 waves, o/down = 32 tiles (21 % of the machine unless K is split by hand),
//   * every CTA pays prologue + pipeline fill + epilogue drain once per tile on a 10-40 us kernel.
// Here ONE CTA per SM stays resident and the flattened (feature tile, K block) space is cut into 148 equal contiguous ranges
// ("stream-K"): every SM streams the same number of weight bytes (+-16 KB), the TMA ring never drains between tiles (the producer is
// already fetching the next tile's first K blocks while the previous accumulator is written out), and the accumulators are double
// buffered in TMEM so the tensor pipe never waits for an epilogue.
//
//   warp 0 : TMA producer   W tile [128 features x 64 K] (swap-AB: the weights are the UMMA *M* operand, every streamed byte is useful)
//                           + the token block [MT x 64 K] (UMMA *N* operand, MT in {16..256}); 6-12 stage mbarrier ring
//   warp 1 : MMA issuer     tcgen05.mma cta_group::1 kind::f16, D^T[128 features, MT tokens] fp32 in TMEM, 2 accumulator stages
//   warp 2 : TMEM allocator
//   warps 4-7 : epilogue    tcgen05.ld -> (segment covers the whole K range) bias / SiLU*mul -> bf16 -> smem [tokens][features] -> ONE TMA
//                           store (cp.async.bulk.tensor, UTMASTG) which also clips the M edge
//                           (partial K range) fp32 partial -> workspace, arrival counter; the LAST contributor of a tile sums all
//                           partials in contributor order (bit-deterministic whatever the arrival order) and runs the final epilogue.
//                           Nobody ever waits for another CTA: no spin, no co-residency requirement, CUDA-graph safe (the finisher
//                           re-arms the counter).
//   SILU_SPLIT: a tile is 64 gate rows + the matching 64 up rows of the checkpoint's [gate; up] matrix (two TMA boxes), the write-out is
//   silu(g) * u, so gate_up_proj + SiluAndMul is one kernel and the [M, 2I] intermediate never exists.
//
// Reference counterparts: cuBLASLt / CuTe-DSL bf16 GEMM (python/sglang/srt/layers/quantization/unquant.py:72-169,
// kernels/ops/gemm/cutedsl_bf16_gemm.py), silu_and_mul (kernels/jit/csrc/elementwise/activation.cuh).
#include "common.cuh"
#include "tma_host.h"

#include <algorithm>
#include <mutex>
#include <unordered_map>

namespace sglb {

constexpr int STK_BF = 128;      // features per tile (UMMA M)
constexpr int STK_BK = 64;       // K block (one 128-byte swizzle row)
constexpr int STK_THREADS = 256;
constexpr int STK_MAX_CTAS = 160;

enum { STK_EPI_NONE = 0, STK_EPI_BIAS = 1, STK_EPI_SILU_SPLIT = 4 };

template <int MT>
struct StkCfg {
  static constexpr int W_BYTES = STK_BF * STK_BK * 2;                  // 16 KB
  static constexpr int A_BYTES = MT * STK_BK * 2;
  static constexpr int STAGE_BYTES = W_BYTES + A_BYTES;
  static constexpr int TOK = MT < 64 ? MT : 64;                        // tokens per epilogue staging pass / TMA store
  static constexpr int OUT_BYTES = TOK * STK_BF * 2;                   // bf16 staging [TOK][128] (dedicated: the ring keeps streaming)
  static constexpr int OUT2_BYTES = TOK * (STK_BF / 2) * 2;            // SiLU result staging [TOK][64]
  static constexpr int RING_BUDGET = 212 * 1024 - OUT_BYTES - OUT2_BYTES;
  static constexpr int RAW_STAGES = RING_BUDGET / STAGE_BYTES;
  static constexpr int STAGES = RAW_STAGES > 12 ? 12 : RAW_STAGES;
  static constexpr int ACC_COLS = MT < 32 ? 32 : MT;                   // one accumulator stage
  static constexpr int TMEM_COLS = 2 * ACC_COLS < 32 ? 32 : 2 * ACC_COLS;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + OUT_BYTES + OUT2_BYTES + 1024 + 512;
  static_assert(STAGES >= 3, "pipeline too shallow");
  static_assert(TMEM_COLS <= 512, "TMEM");
};

struct StkParams {
  const __nv_bfloat16* bias;
  float* ws;            // [grid][2][128 * MT] fp32 partial tiles (slot 0: a CTA's first segment, slot 1: its last)
  int* tile_ctr;        // [n_tiles] arrival counters, zero between launches
  int M, N, n_tiles, num_kb, n_half;
};

// unit range of CTA c:  [c*U/G, (c+1)*U/G)   (64-bit: U up to 1002 tiles x 224 blocks)
__device__ __forceinline__ int stk_cut(long long U, int c, int G) { return (int)((U * c) / G); }

template <int MT, int EPI>
__global__ void __launch_bounds__(STK_THREADS, 1)
gemm_streamk_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmA,
                    const __grid_constant__ CUtensorMap tmC, const StkParams p) {
  using Cfg = StkCfg<MT>;
  constexpr int STAGES = Cfg::STAGES;
  constexpr bool SILU = (EPI == STK_EPI_SILU_SPLIT);
  constexpr int F_OUT = SILU ? STK_BF / 2 : STK_BF;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sW = smem;
  uint8_t* sA = smem + STAGES * Cfg::W_BYTES;
  __nv_bfloat16* sOut = reinterpret_cast<__nv_bfloat16*>(smem + STAGES * Cfg::STAGE_BYTES);
  __nv_bfloat16* sOut2 = reinterpret_cast<__nv_bfloat16*>(smem + STAGES * Cfg::STAGE_BYTES + Cfg::OUT_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES + Cfg::OUT_BYTES + Cfg::OUT2_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* tfull = bars + 2 * STAGES;        // [2]
  uint64_t* tempty = bars + 2 * STAGES + 2;   // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
  volatile int* s_flag = reinterpret_cast<volatile int*>(bars + 2 * STAGES + 5);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int G = (int)gridDim.x, cta = (int)blockIdx.x;
  const int nkb = p.num_kb;
  const long long U = (long long)p.n_tiles * nkb;
  const int u0 = stk_cut(U, cta, G), u1 = stk_cut(U, cta + 1, G);

  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmW); tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmC); }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 128); }
    fence_barrier_init();
  }
  if (warp == 2) { tmem_alloc(tmem_ptr, Cfg::TMEM_COLS); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0 && lane == 0) {
    // ------------------------------------------------------------ producer: one continuous stream over the CTA's unit range
    int stage = 0; uint32_t phase = 0;
    for (int u = u0; u < u1; ++u) {
      const int t = u / nkb, kb = u - t * nkb;
      mbar_wait(&empty[stage], phase ^ 1);
      mbar_expect_tx(&full[stage], Cfg::STAGE_BYTES);
      uint8_t* w_dst = sW + stage * Cfg::W_BYTES;
      if constexpr (SILU) {
        tma_load_2d(w_dst, &tmW, &full[stage], kb * STK_BK, t * F_OUT);
        tma_load_2d(w_dst + Cfg::W_BYTES / 2, &tmW, &full[stage], kb * STK_BK, p.n_half + t * F_OUT);
      } else {
        tma_load_2d(w_dst, &tmW, &full[stage], kb * STK_BK, t * STK_BF);
      }
      tma_load_2d(sA + stage * Cfg::A_BYTES, &tmA, &full[stage], kb * STK_BK, 0);
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1 && lane == 0) {
    // ------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = make_idesc(UMMA_BF16, UMMA_BF16, STK_BF, MT, 0, 0);
    int stage = 0; uint32_t phase = 0;
    int seg = 0;
    for (int u = u0; u < u1; ++seg) {
      const int t = u / nkb, a = u - t * nkb;
      const int b = min(nkb, a + (u1 - u));
      const int as = seg & 1;
      mbar_wait(&tempty[as], ((seg >> 1) & 1) ^ 1);          // the epilogue has drained this accumulator stage
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + as * Cfg::ACC_COLS;
      for (int kb = a; kb < b; ++kb) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint64_t wdesc = make_smem_desc_sw128(smem_u32(sW + stage * Cfg::W_BYTES), 0, 1024);
        const uint64_t adesc = make_smem_desc_sw128(smem_u32(sA + stage * Cfg::A_BYTES), 0, 1024);
#pragma unroll
        for (int k = 0; k < STK_BK / 16; ++k) umma_f16_ss(d_tmem, wdesc + 2 * k, adesc + 2 * k, idesc, (kb > a) || k != 0);
        umma_commit(&empty[stage]);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      umma_commit(&tfull[as]);
      u += b - a;
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------ epilogue
    const int q = warp - 4;
    const int feat = q * 32 + lane;                       // TMEM lane = feature row of the tile
    const int et = threadIdx.x - 128;                     // 0..127
    constexpr int CH = MT < 32 ? MT : 32;                 // TMEM columns (tokens) per tcgen05.ld
    constexpr int TOK = Cfg::TOK;                         // tokens per staging pass / TMA store
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;

    auto load_cols = [&](uint32_t taddr, uint32_t (&r)[32]) {
      if constexpr (CH == 32) {
        tmem_ld_32x32b_x32(taddr, r);
      } else {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                       "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                     : "r"(taddr) : "memory");
      }
      tmem_ld_wait();
    };
    // staging [TOK][128] bf16 (token-major) holds tokens [tok0, tok0 + TOK) of tile t: SiLU combine, then ONE TMA store
    auto store_chunk = [&](int t, int tok0) {
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if constexpr (SILU) {
        for (int idx = et; idx < TOK * 8; idx += 128) {
          const int tok = idx >> 3, v = idx & 7;
          const U4 g = *reinterpret_cast<const U4*>(sOut + tok * STK_BF + v * 8);
          const U4 uu = *reinterpret_cast<const U4*>(sOut + tok * STK_BF + 64 + v * 8);
          auto f = [](uint32_t gg, uint32_t x) {
            const float g0 = bf16_lo(gg), g1 = bf16_hi(gg);
            return pack_bf16(g0 / (1.f + __expf(-g0)) * bf16_lo(x), g1 / (1.f + __expf(-g1)) * bf16_hi(x));
          };
          *reinterpret_cast<U4*>(sOut2 + tok * 64 + v * 8) = U4{f(g.x, uu.x), f(g.y, uu.y), f(g.z, uu.z), f(g.w, uu.w)};
        }
      }
      fence_proxy_async_smem();                           // generic-proxy smem writes -> visible to the TMA (async proxy)
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (et == 0 && tok0 < p.M) {
        tma_store_2d(&tmC, SILU ? (const void*)sOut2 : (const void*)sOut, t * F_OUT, tok0);   // rows >= M are clipped by the tensor map
        tma_store_commit();
        tma_store_wait_read0();                           // staging may be overwritten once the TMA has READ it
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
    };

    int seg = 0;
    for (int u = u0; u < u1; ++seg) {
      const int t = u / nkb, a = u - t * nkb;
      const int b = min(nkb, a + (u1 - u));
      const int as = seg & 1;
      const bool whole = (a == 0 && b == nkb);
      const uint32_t t_row = tmem_base + as * Cfg::ACC_COLS + lane_off;
      mbar_wait(&tfull[as], (seg >> 1) & 1);
      tc_fence_after();
      if (whole) {
        float bsum = 0.f;
        if (EPI == STK_EPI_BIAS && t * STK_BF + feat < p.N) bsum = __bfloat162float(p.bias[t * STK_BF + feat]);
#pragma unroll 1
        for (int tok0 = 0; tok0 < MT; tok0 += TOK) {
          if (tok0 < p.M) {
#pragma unroll 1
            for (int c = 0; c < TOK / CH; ++c) {
              uint32_t r[32];
              load_cols(t_row + tok0 + c * CH, r);
#pragma unroll
              for (int i = 0; i < CH; ++i) sOut[(c * CH + i) * STK_BF + feat] = __float2bfloat16(__uint_as_float(r[i]) + bsum);
            }
          }
          if (tok0 + TOK >= MT) {                         // the accumulator stage is free once its last columns sit in smem
            tc_fence_before();
            mbar_arrive(&tempty[as]);
          }
          store_chunk(t, tok0);
        }
      } else {
        // fp32 partial [token][feature] -> workspace (a warp writes 128 contiguous bytes per token)
        float* my_ws = p.ws + ((size_t)cta * 2 + (u == u0 ? 0 : 1)) * (size_t)(STK_BF * MT);
#pragma unroll 1
        for (int c = 0; c < MT / CH; ++c) {
          uint32_t r[32];
          load_cols(t_row + c * CH, r);
#pragma unroll
          for (int i = 0; i < CH; ++i) my_ws[(c * CH + i) * STK_BF + feat] = __uint_as_float(r[i]);
        }
        tc_fence_before();
        mbar_arrive(&tempty[as]);
        // publish; the LAST contributor of the tile sums every partial in contributor order and finishes the tile
        __threadfence();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (et == 0) {
          const long long lo = (long long)t * nkb, hi = lo + nkb;      // contributors: CTAs whose unit range intersects [lo, hi)
          int c_first = (int)((lo * G) / U);
          while (stk_cut(U, c_first + 1, G) <= lo) ++c_first;
          while (c_first > 0 && stk_cut(U, c_first, G) > lo) --c_first;
          int c_last = c_first;
          while (c_last + 1 < G && stk_cut(U, c_last + 1, G) < hi) ++c_last;
          const int n_contrib = c_last - c_first + 1;
          const int prev = atomicAdd(p.tile_ctr + t, 1);
          int f = 0;
          if (prev == n_contrib - 1) {
            p.tile_ctr[t] = 0;                            // re-arm for the next launch / graph replay
            __threadfence();
            f = 1 | (c_first << 1) | (n_contrib << 12);
          }
          *s_flag = f;
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        const int f = *s_flag;
        if (f & 1) {
          const int c_first = (f >> 1) & 0x7ff, n_contrib = f >> 12;
          const long long lo = (long long)t * nkb;
#pragma unroll 1
          for (int tok0 = 0; tok0 < MT; tok0 += TOK) {
            if (tok0 < p.M) {
              // 128 threads x float4 = four 512-byte token rows per step (coalesced).  All contributors' loads are issued BEFORE the first add
              // (independent L2 round trips in flight together); the sum itself runs in contributor order -> bit-deterministic.
              const float* base0 = p.ws + (size_t)tok0 * STK_BF;
              size_t coff[8];
#pragma unroll
              for (int cc = 0; cc < 8; ++cc) {
                const int c = c_first + min(cc, n_contrib - 1);
                const int slot = (stk_cut(U, c, G) >= lo) ? 0 : 1;   // tile t is c's first segment unless c's range starts before it
                coff[cc] = ((size_t)c * 2 + slot) * (size_t)(STK_BF * MT);
              }
#pragma unroll 1
              for (int idx = et; idx < TOK * (STK_BF / 4); idx += 128) {
                const int tok = idx / (STK_BF / 4), grp = idx % (STK_BF / 4);
                const size_t eoff = (size_t)tok * STK_BF + grp * 4;
                float4 v[8];
#pragma unroll
                for (int cc = 0; cc < 8; ++cc)
                  if (cc < n_contrib) v[cc] = __ldcg(reinterpret_cast<const float4*>(base0 + coff[cc] + eoff));
                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int cc = 0; cc < 8; ++cc)
                  if (cc < n_contrib) { acc.x += v[cc].x; acc.y += v[cc].y; acc.z += v[cc].z; acc.w += v[cc].w; }
                for (int cc = 8; cc < n_contrib; ++cc) {            // more than 8 contributors (tiny N, huge K): plain loop
                  const int c = c_first + cc;
                  const int slot = (stk_cut(U, c, G) >= lo) ? 0 : 1;
                  const float4 w = __ldcg(reinterpret_cast<const float4*>(base0 + ((size_t)c * 2 + slot) * (size_t)(STK_BF * MT) + eoff));
                  acc.x += w.x; acc.y += w.y; acc.z += w.z; acc.w += w.w;
                }
                if (EPI == STK_EPI_BIAS) {
                  const int f0 = t * STK_BF + grp * 4;
                  if (f0 + 3 < p.N) {
                    acc.x += __bfloat162float(p.bias[f0]); acc.y += __bfloat162float(p.bias[f0 + 1]);
                    acc.z += __bfloat162float(p.bias[f0 + 2]); acc.w += __bfloat162float(p.bias[f0 + 3]);
                  }
                }
                uint2 o;
                o.x = pack_bf16(acc.x, acc.y);
                o.y = pack_bf16(acc.z, acc.w);
                *reinterpret_cast<uint2*>(sOut + tok * STK_BF + grp * 4) = o;
              }
            }
            store_chunk(t, tok0);
          }
        }
      }
      u += b - a;
    }
    if (et == 0) tma_store_wait0();                       // all output bytes are globally visible before the kernel ends
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

// ---------------------------------------------------------------------------------------------------------------------------------
struct StkWorkspace { float* ws = nullptr; int* ctr = nullptr; };
constexpr int STK_MAX_TILES = 4096;

static StkWorkspace stk_workspace(cudaStream_t stream) {
  // one workspace per (device, stream): two GEMMs on different streams (two-batch overlap) must not share partials / counters.
  // Allocated at first eager use; a CUDA-graph capture reuses the entry its stream created during the warm-up steps.
  static std::mutex mu;
  static std::unordered_map<unsigned long long, StkWorkspace> pool;
  int dev = 0;
  SGLB_CUDA_CHECK(cudaGetDevice(&dev));
  const unsigned long long key = ((unsigned long long)dev << 56) ^ (unsigned long long)(uintptr_t)stream;
  std::lock_guard<std::mutex> lk(mu);
  auto it = pool.find(key);
  if (it != pool.end()) return it->second;
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  SGLB_CUDA_CHECK(cudaStreamIsCapturing(stream, &st));
  if (st != cudaStreamCaptureStatusNone)
    throw std::runtime_error("sglb gemm_streamk: first use on this stream is inside a CUDA-graph capture (run an eager warm-up step first)");
  StkWorkspace w;
  SGLB_CUDA_CHECK(cudaMalloc(&w.ws, (size_t)STK_MAX_CTAS * 2 * STK_BF * 256 * sizeof(float)));
  SGLB_CUDA_CHECK(cudaMalloc(&w.ctr, STK_MAX_TILES * sizeof(int)));
  SGLB_CUDA_CHECK(cudaMemset(w.ctr, 0, STK_MAX_TILES * sizeof(int)));
  pool[key] = w;
  return w;
}

template <int MT, int EPI>
static void launch_streamk(const void* A, const void* W, void* C, const void* bias, int M, int N, int K, int lda, int ldw, int ldc,
                           int n_half, int num_sms, cudaStream_t stream) {
  using Cfg = StkCfg<MT>;
  constexpr bool SILU = (EPI == STK_EPI_SILU_SPLIT);
  const int w_rows = SILU ? 2 * n_half : N;
  const int f_out = SILU ? STK_BF / 2 : STK_BF;
  uint64_t dW[2] = {(uint64_t)K, (uint64_t)w_rows}, sWs[1] = {(uint64_t)ldw * 2};
  uint64_t dA[2] = {(uint64_t)K, (uint64_t)M}, sAs[1] = {(uint64_t)lda * 2};
  uint64_t dC[2] = {(uint64_t)N, (uint64_t)M}, sCs[1] = {(uint64_t)ldc * 2};
  uint32_t bW[2] = {STK_BK, (uint32_t)f_out}, bA[2] = {STK_BK, (uint32_t)MT}, bC[2] = {(uint32_t)f_out, (uint32_t)Cfg::TOK};
  CUtensorMap tmW = make_tmap(W, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dW, sWs, bW, CU_TENSOR_MAP_SWIZZLE_128B);
  CUtensorMap tmA = make_tmap(A, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dA, sAs, bA, CU_TENSOR_MAP_SWIZZLE_128B);
  CUtensorMap tmC = make_tmap(C, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dC, sCs, bC, CU_TENSOR_MAP_SWIZZLE_NONE);
  auto kern = gemm_streamk_kernel<MT, EPI>;
  static bool configured = false;
  if (!configured) {
    SGLB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    configured = true;
  }
  StkParams p;
  p.bias = (const __nv_bfloat16*)bias;
  p.M = M; p.N = N; p.n_half = n_half;
  p.n_tiles = (N + f_out - 1) / f_out;
  p.num_kb = (K + STK_BK - 1) / STK_BK;
  if (p.n_tiles > STK_MAX_TILES) throw std::runtime_error("sglb gemm_streamk: too many output tiles");
  const StkWorkspace w = stk_workspace(stream);
  p.ws = w.ws; p.tile_ctr = w.ctr;
  const long long U = (long long)p.n_tiles * p.num_kb;
  const int grid = (int)std::min<long long>(std::min(num_sms, STK_MAX_CTAS), U);
  kern<<<grid, STK_THREADS, Cfg::SMEM_BYTES, stream>>>(tmW, tmA, tmC, p);
  SGLB_CUDA_CHECK(cudaGetLastError());
}

template <int EPI>
static void stk_dispatch(int M, const void* A, const void* W, void* C, const void* bias, int N, int K, int lda, int ldw, int ldc, int n_half,
                         int num_sms, cudaStream_t stream) {
  if (M <= 16) launch_streamk<16, EPI>(A, W, C, bias, M, N, K, lda, ldw, ldc, n_half, num_sms, stream);
  else if (M <= 32) launch_streamk<32, EPI>(A, W, C, bias, M, N, K, lda, ldw, ldc, n_half, num_sms, stream);
  else if (M <= 64) launch_streamk<64, EPI>(A, W, C, bias, M, N, K, lda, ldw, ldc, n_half, num_sms, stream);
  else if (M <= 128) launch_streamk<128, EPI>(A, W, C, bias, M, N, K, lda, ldw, ldc, n_half, num_sms, stream);
  else launch_streamk<256, EPI>(A, W, C, bias, M, N, K, lda, ldw, ldc, n_half, num_sms, stream);
}

// epilogue: 0 none, 1 bias, 4 split-silu (W = [gate; up] with N = intermediate size, C is [M, N])
void gemm_bf16_streamk(const void* A, const void* W, void* C, const void* bias, int M, int N, int K, int lda, int ldw, int ldc, int epilogue,
                       int num_sms, cudaStream_t stream) {
  if (M <= 0 || N <= 0) return;
  if (M > 256) throw std::runtime_error("sglb gemm_streamk: M must be <= 256");
  if (K % 8 || lda % 8 || ldw % 8 || ldc % 8 || N % 8) throw std::runtime_error("sglb gemm_streamk: K, N, lda, ldw, ldc must be multiples of 8");
  switch (epilogue) {
    case STK_EPI_NONE: stk_dispatch<STK_EPI_NONE>(M, A, W, C, bias, N, K, lda, ldw, ldc, 0, num_sms, stream); break;
    case STK_EPI_BIAS: stk_dispatch<STK_EPI_BIAS>(M, A, W, C, bias, N, K, lda, ldw, ldc, 0, num_sms, stream); break;
    case STK_EPI_SILU_SPLIT: stk_dispatch<STK_EPI_SILU_SPLIT>(M, A, W, C, bias, N, K, lda, ldw, ldc, N, num_sms, stream); break;
    default: throw std::runtime_error("sglb gemm_streamk: bad epilogue");
  }
}

}  // namespace sglb
