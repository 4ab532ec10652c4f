// Paged RadixAttention forward for sm_100a: ONE tcgen05/TMEM kernel family for ragged prefill (extend)
// and decode, reading K/V only through the page table with TMA page gathers.
//
//   grid  = (q tiles, kv_heads * kv_splits, requests)        block = 192 threads
//   warp 0    : TMA producer — Q tile once, then K/V tiles page by page (page table -> TMA coords)
//   warp 1    : MMA issuer   — S = Q K^T (SS, K-major) into TMEM, O += P V (TS: P from TMEM, V MN-major)
//   warps 2-5 : softmax      — tcgen05.ld S, online softmax with lazy O rescale, P (bf16) back to TMEM,
//                              final normalise + store (or fp32 partial + LSE for split-KV)
//   Q rows pack GQA: row r of the 128-row tile = (token r / G, head r % G) so one KV tile serves all
//   G query heads of its KV head (decode: G valid rows, only one softmax warp does work).
//
// Stands in for the reference's trtllm-gen / FA4 attention (python/sglang/srt/layers/attention/
// trtllm_mha_backend.py:1113,1470; kernels/ops/attention/flash_attn/cute/flash_fwd_sm100.py) and the Triton
// extend/decode kernels (kernels/ops/attention/extend_attention.py:305, decode_attention.py:472).
#include <math.h>

#include "common.cuh"
#include "tma_host.h"

#include <algorithm>

namespace sglb {

constexpr int ATT_D = 128;          // head dim
// kv tile KT is a template parameter: 128 for prefill (compute-bound, 1 CTA/SM), 64 for decode (latency-bound: 256
// TMEM columns + 96 KB smem so two CTAs share an SM and overlap their load -> MMA -> softmax chains)
constexpr int ATT_THREADS = 192;
constexpr int ATT_TILE_BYTES = 128 * ATT_D * 2;  // 32 KB (Q, K or V tile)
constexpr int ATT_HALF_BYTES = ATT_TILE_BYTES / 2;
template <int KT> constexpr int att_smem_bytes() { return ATT_TILE_BYTES + 4 * KT * ATT_D * 2 + 1024 + 256; }
constexpr float ATT_RESCALE_THRESHOLD = 8.0f;  // log2 units; lazy rescale (FA4-style)

struct AttnParams {
  __nv_bfloat16* out;
  float* part_o;      // [splits, T, Hq, D]
  float* part_lse;    // [splits, T, Hq]
  float* lse;         // optional [T, Hq]
  const int* page_table;
  const int* req_slots;
  const int* seq_lens;
  const int* cu_q;    // nullptr => decode (one query token per request)
  int pt_stride, Hq, Hkv, G, page_size, num_splits, T;
  float scale_log2, logit_cap;
  int causal, window;
  // persistent variant: work items = the (x, y, z) cells of the classic grid, handed out by an atomic counter
  int gx, gy, total_items;
  int* sched;         // [0] next item, [1] CTAs finished (the last one resets both: graph-replay safe)
  const unsigned* tree_anc;   // speculative tree verify: [T] bit k of tree_anc[t] = query token t may see draft key `prefix + k` (classic launch)
  int legacy_epilogue_wait;   // FAULT INJECTION ONLY (SGLB_ATTN_LEGACY_EPILOGUE=1): re-creates the round-1 epilogue hand-off (wait on o_done by
                              // tile parity) so tools/attn_stress.py can reproduce the premature-pass race / the TP>=4 deadlock it caused
};

template <int KT, bool HAS_CAP>
__global__ void __launch_bounds__(ATT_THREADS, (KT <= 64) ? 2 : 1)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  const int b = blockIdx.z;
  const int kvh = blockIdx.y / p.num_splits;
  const int split = blockIdx.y % p.num_splits;
  const int G = p.G;
  const int TQ = 128 / G;
  const int q_start = p.cu_q ? p.cu_q[b] : b;
  const int q_len = p.cu_q ? (p.cu_q[b + 1] - q_start) : 1;
  const int seq_len = p.seq_lens[b];
  const int prefix = seq_len - q_len;
  const int t0 = blockIdx.x * TQ;
  if (t0 >= q_len) return;
  const int t_hi = min(t0 + TQ, q_len);                       // tokens [t0, t_hi) of this request
  const int kv_hi = p.causal ? min(seq_len, prefix + t_hi) : seq_len;
  int kv_lo = 0;
  if (p.window > 0) kv_lo = max(0, prefix + t0 - p.window + 1);
  constexpr int KV_TILE_BYTES = KT * ATT_D * 2;      // K or V tile
  constexpr int KV_HALF_BYTES = KV_TILE_BYTES / 2;   // one 64-wide d-block of the tile
  constexpr int TMEM_COLS = (KT <= 64) ? 256 : 512;  // S0 | S1 | O(128)
  const int tile_lo_all = kv_lo / KT;
  const int tile_hi_all = (kv_hi + KT - 1) / KT;
  const int per_split = (tile_hi_all - tile_lo_all + p.num_splits - 1) / p.num_splits;
  const int tile_begin = tile_lo_all + split * per_split;
  const int tile_end = min(tile_hi_all, tile_begin + per_split);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (tile_begin >= tile_end) {
    // empty split: publish a neutral partial so the merge kernel can ignore it
    if (p.num_splits > 1 && warp >= 2) {
      const int r = ((warp & 3) << 5) + lane;
      const int tok = t0 + r / G;
      if (tok < q_len) {
        const size_t row = ((size_t)split * p.T + q_start + tok) * p.Hq + kvh * G + (r % G);
        p.part_lse[row] = -INFINITY;
        float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        float4* dst = reinterpret_cast<float4*>(p.part_o + row * ATT_D);
        for (int i = 0; i < ATT_D / 4; ++i) dst[i] = z;
      }
    }
    return;
  }
  const int n_tiles = tile_end - tile_begin;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = smem + ATT_TILE_BYTES;                       // 2 stages
  uint8_t* sV = smem + ATT_TILE_BYTES + 2 * KV_TILE_BYTES;   // 2 stages
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + ATT_TILE_BYTES + 4 * KV_TILE_BYTES);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;    // [2]
  uint64_t* k_empty = bars + 3;   // [2]
  uint64_t* v_full = bars + 5;    // [2]
  uint64_t* v_empty = bars + 7;   // [2]
  uint64_t* s_full = bars + 9;    // [2]
  uint64_t* p_full = bars + 11;   // [2]
  uint64_t* o_done = bars + 13;    // every PV tile (lazy-rescale hand-off: exactly one phase behind the waiter)
  uint64_t* o_final = bars + 14;   // the LAST PV tile only (epilogue hand-off).  Waiting on o_done by tile parity in the epilogue was a
                                   // protocol flaw: a warp that had not observed phase L-1 passes try_wait.parity(L&1) while the barrier
                                   // is still two phases behind (PV(L-1) in flight) and reads O early.
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 15);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1);
      mbar_init(&s_full[i], 1); mbar_init(&p_full[i], 128);
    }
    mbar_init(o_done, 1); mbar_init(o_final, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t tmem_S0 = tmem_base;           // S buffers: cols [0,KT) and [KT,2KT); P aliases the first KT/2 cols of each
  const uint32_t tmem_O = tmem_base + 2 * KT;   // O: 128 cols

  if (warp == 0 && lane == 0) {
    // ------------------------------------------------------------ producer
    mbar_expect_tx(q_full, ATT_TILE_BYTES);
    tma_load_4d(sQ, &tmQ, q_full, 0, 0, kvh, q_start + t0);
    tma_load_4d(sQ + ATT_HALF_BYTES, &tmQ, q_full, 64, 0, kvh, q_start + t0);
    const int ps = p.page_size;
    const int pages_per_tile = KT / ps;
    const int* pt = p.page_table + (size_t)p.req_slots[b] * p.pt_stride;
    for (int j = 0; j < n_tiles; ++j) {
      const int st = j & 1;
      const uint32_t ph = (j >> 1) & 1;
      const int pg0 = (tile_begin + j) * pages_per_tile;
      mbar_wait(&k_empty[st], ph ^ 1);
      mbar_expect_tx(&k_full[st], KV_TILE_BYTES);
      for (int i = 0; i < pages_per_tile; ++i) {
        const int page = pt[min(pg0 + i, p.pt_stride - 1)];
        uint8_t* dst = sK + st * KV_TILE_BYTES + i * ps * 128;
        tma_load_3d(dst, &tmK, &k_full[st], 0, kvh, page * ps);
        tma_load_3d(dst + KV_HALF_BYTES, &tmK, &k_full[st], 64, kvh, page * ps);
      }
      mbar_wait(&v_empty[st], ph ^ 1);
      mbar_expect_tx(&v_full[st], KV_TILE_BYTES);
      for (int i = 0; i < pages_per_tile; ++i) {
        const int page = pt[min(pg0 + i, p.pt_stride - 1)];
        uint8_t* dst = sV + st * KV_TILE_BYTES + i * ps * 128;
        tma_load_3d(dst, &tmV, &v_full[st], 0, kvh, page * ps);
        tma_load_3d(dst + KV_HALF_BYTES, &tmV, &v_full[st], 64, kvh, page * ps);
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc_s = make_idesc(UMMA_BF16, UMMA_BF16, 128, KT, 0, 0);
    constexpr uint32_t idesc_o = make_idesc(UMMA_BF16, UMMA_BF16, 128, ATT_D, 0, 1);   // B (V) is MN-major
    const uint32_t q_addr = smem_u32(sQ);
    auto issue_S = [&](int j) {
      const int st = j & 1;
      mbar_wait(&k_full[st], (j >> 1) & 1);
      tc_fence_after();
      const uint32_t k_addr = smem_u32(sK + st * KV_TILE_BYTES);
#pragma unroll
      for (int k = 0; k < ATT_D / 16; ++k) {
        const uint32_t qoff = (k >> 2) * ATT_HALF_BYTES + (k & 3) * 32;
        const uint32_t koff = (k >> 2) * KV_HALF_BYTES + (k & 3) * 32;
        umma_f16_ss(tmem_S0 + st * KT, make_smem_desc_sw128(q_addr + qoff, 0, 1024),
                    make_smem_desc_sw128(k_addr + koff, 0, 1024), idesc_s, k != 0);
      }
      umma_commit(&k_empty[st]);
      umma_commit(&s_full[st]);
    };
    mbar_wait(q_full, 0);
    issue_S(0);
    for (int j = 0; j < n_tiles; ++j) {
      const int st = j & 1;
      const uint32_t ph = (j >> 1) & 1;
      if (j + 1 < n_tiles) issue_S(j + 1);
      mbar_wait(&p_full[st], ph);
      mbar_wait(&v_full[st], ph);
      tc_fence_after();
      const uint32_t v_addr = smem_u32(sV + st * KV_TILE_BYTES);
#pragma unroll
      for (int k = 0; k < KT / 16; ++k) {
        // A = P (bf16 pairs in TMEM: 16 K-elements = 8 columns), B = V rows [16k, 16k+16) (MN-major)
        umma_f16_ts(tmem_O, tmem_S0 + st * KT + k * 8, make_smem_desc_sw128(v_addr + k * 2048, KV_HALF_BYTES, 1024),
                    idesc_o, (j | k) != 0);
      }
      umma_commit(&v_empty[st]);
      umma_commit(o_done);
      if (j + 1 == n_tiles) umma_commit(o_final);
    }
  } else if (warp >= 2) {
    // ------------------------------------------------------------ softmax + epilogue
    const int q = warp & 3;                       // TMEM lane quadrant
    const int r = (q << 5) + lane;                // row in the 128-row tile
    const int tok = t0 + r / G;                   // token index inside this request's queries
    const bool row_valid = tok < q_len;
    const bool warp_valid = (t0 + (q << 5) / G) < q_len;
    const int q_pos = prefix + tok;               // absolute position of this query token
    const uint32_t lane_addr = (uint32_t)(q << 5) << 16;
    float m_i = -INFINITY, l_i = 0.f;
    // loop-invariant uniforms hoisted into registers
    const float scale_log2 = p.scale_log2;
    const float cap = p.logit_cap, inv_cap = HAS_CAP ? 1.f / p.logit_cap : 0.f;
    const float scale_nat = scale_log2 * 0.6931471805599453f, inv_scale_nat = 1.f / scale_nat;
    const int kv_last = p.causal ? min(seq_len - 1, q_pos) : seq_len - 1;
    const int kv_first = p.window > 0 ? q_pos - p.window + 1 : -(1 << 30);

    for (int j = 0; j < n_tiles; ++j) {
      const int st = j & 1;
      const uint32_t ph = (j >> 1) & 1;
      mbar_wait(&s_full[st], ph);
      tc_fence_after();
      if (warp_valid) {
        constexpr int NC = KT / 32;
        const uint32_t s_addr = tmem_S0 + lane_addr + st * KT;
        uint32_t sr[NC][32];
#pragma unroll
        for (int c = 0; c < NC; ++c) tmem_ld_32x32b_x32(s_addr + c * 32, sr[c]);
        tmem_ld_wait();
        const int kv0 = (tile_begin + j) * KT;
        const bool need_mask = (kv0 + KT > kv_hi) || (p.causal && kv0 + KT - 1 > prefix + t0) ||
                               (p.window > 0 && kv0 < prefix + t_hi - p.window);
        // The inner loops are kept branch-free and tiny (the 128-wide body is fully unrolled: any per-element branch or
        // constant-bank reload multiplies into thousands of instructions and thrashes the instruction cache).
        if (HAS_CAP) {   // Gemma-2 style soft-capping (separate template instance): rewrite raw so raw*scale_log2 is the capped logit
#pragma unroll
          for (int c = 0; c < NC; ++c)
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const float x = __uint_as_float(sr[c][i]) * scale_nat;
              sr[c][i] = __float_as_uint(cap * tanhf(x * inv_cap) * inv_scale_nat);
            }
        }
        float mx_raw = -INFINITY;
        if (need_mask) {
          const int hi = kv_last - kv0;     // columns > hi are masked   (per-thread: causal / sequence end)
          const int lo = kv_first - kv0;    // columns < lo are masked   (sliding window)
#pragma unroll
          for (int c = 0; c < NC; ++c) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const int col = c * 32 + i;
              const float v = (col > hi || col < lo) ? -INFINITY : __uint_as_float(sr[c][i]);
              sr[c][i] = __float_as_uint(v);
              mx_raw = fmaxf(mx_raw, v);
            }
          }
        } else {
#pragma unroll
          for (int c = 0; c < NC; ++c) {
#pragma unroll
            for (int i = 0; i < 32; ++i) mx_raw = fmaxf(mx_raw, __uint_as_float(sr[c][i]));
          }
        }
        if (p.tree_anc != nullptr && kv0 + KT > prefix) {
          // token-tree verification: keys below `prefix` are the committed context (always visible); draft key prefix + k is visible to a
          // query node iff k is one of its ancestors or itself.  Separate pass: the common path above stays untouched.
          const unsigned anc = row_valid ? p.tree_anc[q_start + tok] : 0xffffffffu;
          const int tree_lo = prefix - kv0;
          mx_raw = -INFINITY;
#pragma unroll
          for (int c = 0; c < NC; ++c) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const int rel = c * 32 + i - tree_lo;
              const bool hidden = rel >= 0 && ((anc >> (rel & 31)) & 1u) == 0u;
              const float v = hidden ? -INFINITY : __uint_as_float(sr[c][i]);
              sr[c][i] = __float_as_uint(v);
              mx_raw = fmaxf(mx_raw, v);
            }
          }
        }
        const float mx = mx_raw * scale_log2;   // scale_log2 > 0: max commutes with the scaling
        const float m_new = fmaxf(m_i, mx);
        const bool first = (j == 0);
        const bool need = (!first) && (m_new > m_i + ATT_RESCALE_THRESHOLD);
        float m_use = m_i;
        if (first) {
          m_use = m_new;
        } else if (__any_sync(0xffffffffu, need)) {
          // lazy rescale of the O row in TMEM (warp-collective): wait for PV(j-1), scale, write back
          const float f = (m_new == -INFINITY) ? 1.f : exp2f(m_i - m_new);
          mbar_wait(o_done, (uint32_t)((j - 1) & 1));
          tc_fence_after();
#pragma unroll 1
          for (int c = 0; c < 4; ++c) {
            uint32_t orow[32];
            tmem_ld_32x32b_x32(tmem_O + lane_addr + c * 32, orow);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) orow[i] = __float_as_uint(__uint_as_float(orow[i]) * f);
            tmem_st_32x32b_x32(tmem_O + lane_addr + c * 32, orow);
          }
          tmem_st_wait();
          l_i *= f;
          m_use = m_new;
        }
        m_i = m_use;
        const float m_sub = (m_i == -INFINITY) ? 0.f : m_i;
        float psum = 0.f;
        const uint32_t p_addr = tmem_S0 + lane_addr + st * KT;
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            // p = 2^(raw*scale - m): one FFMA + one MUFU.EX2 per element
            const float p0 = ex2_approx(fmaf(__uint_as_float(sr[c][i]), scale_log2, -m_sub));
            const float p1 = ex2_approx(fmaf(__uint_as_float(sr[c][i + 1]), scale_log2, -m_sub));
            psum += p0 + p1;
            pk[i >> 1] = pack_bf16(p0, p1);
          }
          tmem_st_32x32b_x16(p_addr + c * 16, pk);
        }
        l_i += psum;
        tmem_st_wait();
      }
      tc_fence_before();
      mbar_arrive(&p_full[st]);
    }

    // ---- epilogue: wait for the last PV, normalise, store
    if (p.legacy_epilogue_wait) mbar_wait(o_done, (uint32_t)((n_tiles - 1) & 1));
    else mbar_wait(o_final, 0);
    tc_fence_after();
    if (warp_valid) {
      const float inv_l = (l_i > 0.f) ? 1.f / l_i : 0.f;
      const float lse = (l_i > 0.f) ? (m_i + log2f(l_i)) * 0.6931471805599453f : -INFINITY;
      const int g = r % G;
      const size_t out_row = ((size_t)(q_start + tok)) * p.Hq + kvh * G + g;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t orow[32];
        tmem_ld_32x32b_x32(tmem_O + lane_addr + c * 32, orow);
        tmem_ld_wait();
        if (row_valid) {
          if (p.num_splits == 1) {
            __nv_bfloat16* dst = p.out + out_row * ATT_D + c * 32;
#pragma unroll
            for (int v = 0; v < 4; ++v) {
              U4 o;
              o.x = pack_bf16(__uint_as_float(orow[v * 8 + 0]) * inv_l, __uint_as_float(orow[v * 8 + 1]) * inv_l);
              o.y = pack_bf16(__uint_as_float(orow[v * 8 + 2]) * inv_l, __uint_as_float(orow[v * 8 + 3]) * inv_l);
              o.z = pack_bf16(__uint_as_float(orow[v * 8 + 4]) * inv_l, __uint_as_float(orow[v * 8 + 5]) * inv_l);
              o.w = pack_bf16(__uint_as_float(orow[v * 8 + 6]) * inv_l, __uint_as_float(orow[v * 8 + 7]) * inv_l);
              st128(dst + v * 8, o);
            }
          } else {
            float* dst = p.part_o + (((size_t)split * p.T + q_start + tok) * p.Hq + kvh * G + g) * ATT_D + c * 32;
#pragma unroll
            for (int v = 0; v < 8; ++v) {
              float4 o = make_float4(__uint_as_float(orow[v * 4 + 0]) * inv_l, __uint_as_float(orow[v * 4 + 1]) * inv_l,
                                     __uint_as_float(orow[v * 4 + 2]) * inv_l, __uint_as_float(orow[v * 4 + 3]) * inv_l);
              *reinterpret_cast<float4*>(dst + v * 4) = o;
            }
          }
        }
      }
      if (row_valid) {
        if (p.num_splits == 1) {
          if (p.lse) p.lse[out_row] = lse;
        } else {
          p.part_lse[((size_t)split * p.T + q_start + tok) * p.Hq + kvh * G + g] = lse;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ---------------------------------------------------------------------------------------------------------------------
// Persistent variant.  The classic kernel above pays barrier init + TMEM alloc + first Q/K/V round trip + epilogue drain once
// per (request, kv head) — ~25-30 % of a decode CTA's life (8 KV tiles).  Here a CTA stays resident (2 per SM for KT=64), takes
// work items from an atomic counter (ragged lengths balance themselves) and pipelines ACROSS items: the producer warp is
// already streaming the next item's Q/K/V while the softmax warps normalise and store the previous one; TMEM, barriers and
// tensor maps are set up once.  Extra barriers: item ring (item_full/item_empty), q_empty (Q buffer reuse), o_free (O
// accumulator reuse).  All ring barriers use running counters so phases carry across items.
template <int KT, bool HAS_CAP>
__global__ void __launch_bounds__(ATT_THREADS, (KT <= 64) ? 2 : 1)
attn_fwd_tc_persist_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                           const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  const int G = p.G;
  const int TQ = 128 / G;
  constexpr int KV_TILE_BYTES = KT * ATT_D * 2;
  constexpr int KV_HALF_BYTES = KV_TILE_BYTES / 2;
  constexpr int TMEM_COLS = (KT <= 64) ? 256 : 512;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = smem + ATT_TILE_BYTES;
  uint8_t* sV = smem + ATT_TILE_BYTES + 2 * KV_TILE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + ATT_TILE_BYTES + 4 * KV_TILE_BYTES);
  uint64_t* q_full = bars;
  uint64_t* q_empty = bars + 1;
  uint64_t* k_full = bars + 2;    // [2]
  uint64_t* k_empty = bars + 4;   // [2]
  uint64_t* v_full = bars + 6;    // [2]
  uint64_t* v_empty = bars + 8;   // [2]
  uint64_t* s_full = bars + 10;   // [2]
  uint64_t* p_full = bars + 12;   // [2]
  uint64_t* o_done = bars + 14;
  uint64_t* o_free = bars + 15;
  uint64_t* item_full = bars + 16;   // [2]
  uint64_t* item_empty = bars + 18;  // [2]
  uint64_t* o_final = bars + 20;     // last PV of an item (phase = item ordinal): the epilogue hand-off, see the classic kernel
  volatile int* s_item = reinterpret_cast<volatile int*>(bars + 21);   // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 22);
  volatile uint32_t* dbg = reinterpret_cast<volatile uint32_t*>(bars + 23);   // [6] per-warp role state, dumped on a wait timeout
#define SGLB_ATT_DBG(code_, a_, b_) dbg[warp] = ((uint32_t)(code_) << 28) | (((uint32_t)(a_) & 0x3fff) << 14) | ((uint32_t)(b_) & 0x3fff)
#define SGLB_ATT_WAIT(bar_, par_) mbar_wait((bar_), (par_), dbg, 6)

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1); mbar_init(q_empty, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1);
      mbar_init(&s_full[i], 1); mbar_init(&p_full[i], 128);
      mbar_init(&item_full[i], 1); mbar_init(&item_empty[i], 5);   // MMA thread + 4 softmax warps
    }
    mbar_init(o_done, 1); mbar_init(o_free, 4); mbar_init(o_final, 1);
    fence_barrier_init();
  }
  if (lane == 0) dbg[warp] = 0;
  if (warp == 1) {
    tmem_alloc(tmem_ptr, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t tmem_S0 = tmem_base;
  const uint32_t tmem_O = tmem_base + 2 * KT;

  // geometry of one work item (identical code in every role: pure function of the item id)
  struct Item { int b, kvh, split, t0, t_hi, q_start, q_len, seq_len, prefix, kv_hi, tile_begin, n_tiles; bool valid; };
  auto decode_item = [&](int idx) {
    Item it;
    const int x = p.gx - 1 - (idx % p.gx);      // longest (latest) q tiles of a request first
    const int y = (idx / p.gx) % p.gy;
    it.b = idx / (p.gx * p.gy);
    it.kvh = y / p.num_splits; it.split = y % p.num_splits;
    it.q_start = p.cu_q ? p.cu_q[it.b] : it.b;
    it.q_len = p.cu_q ? (p.cu_q[it.b + 1] - it.q_start) : 1;
    it.seq_len = p.seq_lens[it.b];
    it.prefix = it.seq_len - it.q_len;
    it.t0 = x * TQ;
    it.valid = it.t0 < it.q_len;
    it.t_hi = min(it.t0 + TQ, it.q_len);
    it.kv_hi = p.causal ? min(it.seq_len, it.prefix + it.t_hi) : it.seq_len;
    int kv_lo = 0;
    if (p.window > 0) kv_lo = max(0, it.prefix + it.t0 - p.window + 1);
    const int tile_lo_all = kv_lo / KT, tile_hi_all = (it.kv_hi + KT - 1) / KT;
    const int per_split = (tile_hi_all - tile_lo_all + p.num_splits - 1) / p.num_splits;
    it.tile_begin = tile_lo_all + it.split * per_split;
    const int tile_end = min(tile_hi_all, it.tile_begin + per_split);
    it.n_tiles = it.valid ? max(0, tile_end - it.tile_begin) : 0;
    return it;
  };

  if (warp == 0 && lane == 0) {
    // ------------------------------------------------------------ scheduler + producer
    uint32_t n_it = 0, n_q = 0, g = 0;
    for (;;) {
      const int slot = n_it & 1;
      SGLB_ATT_WAIT(&item_empty[slot], ((n_it >> 1) & 1) ^ 1);
      const int idx = atomicAdd(p.sched, 1);
      SGLB_ATT_DBG(1, n_it, g);
      s_item[slot] = idx;
      mbar_arrive(&item_full[slot]);
      ++n_it;
      if (idx >= p.total_items) break;
      const Item it = decode_item(idx);
      if (it.n_tiles == 0) continue;
      SGLB_ATT_WAIT(q_empty, (n_q & 1) ^ 1);
      mbar_expect_tx(q_full, ATT_TILE_BYTES);
      tma_load_4d(sQ, &tmQ, q_full, 0, 0, it.kvh, it.q_start + it.t0);
      tma_load_4d(sQ + ATT_HALF_BYTES, &tmQ, q_full, 64, 0, it.kvh, it.q_start + it.t0);
      ++n_q;
      const int ps = p.page_size;
      const int pages_per_tile = KT / ps;
      const int* pt = p.page_table + (size_t)p.req_slots[it.b] * p.pt_stride;
      for (int j = 0; j < it.n_tiles; ++j, ++g) {
        const int st = g & 1;
        const uint32_t ph = (g >> 1) & 1;
        const int pg0 = (it.tile_begin + j) * pages_per_tile;
        SGLB_ATT_WAIT(&k_empty[st], ph ^ 1);
        mbar_expect_tx(&k_full[st], KV_TILE_BYTES);
        for (int i = 0; i < pages_per_tile; ++i) {
          const int page = pt[min(pg0 + i, p.pt_stride - 1)];
          uint8_t* dst = sK + st * KV_TILE_BYTES + i * ps * 128;
          tma_load_3d(dst, &tmK, &k_full[st], 0, it.kvh, page * ps);
          tma_load_3d(dst + KV_HALF_BYTES, &tmK, &k_full[st], 64, it.kvh, page * ps);
        }
        SGLB_ATT_WAIT(&v_empty[st], ph ^ 1);
        mbar_expect_tx(&v_full[st], KV_TILE_BYTES);
        for (int i = 0; i < pages_per_tile; ++i) {
          const int page = pt[min(pg0 + i, p.pt_stride - 1)];
          uint8_t* dst = sV + st * KV_TILE_BYTES + i * ps * 128;
          tma_load_3d(dst, &tmV, &v_full[st], 0, it.kvh, page * ps);
          tma_load_3d(dst + KV_HALF_BYTES, &tmV, &v_full[st], 64, it.kvh, page * ps);
        }
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc_s = make_idesc(UMMA_BF16, UMMA_BF16, 128, KT, 0, 0);
    constexpr uint32_t idesc_o = make_idesc(UMMA_BF16, UMMA_BF16, 128, ATT_D, 0, 1);
    const uint32_t q_addr = smem_u32(sQ);
    uint32_t n_it = 0, n_q = 0, g = 0;
    for (;;) {
      const int slot = n_it & 1;
      SGLB_ATT_WAIT(&item_full[slot], (n_it >> 1) & 1);
      const int idx = s_item[slot];
      mbar_arrive(&item_empty[slot]);
      ++n_it;
      if (idx >= p.total_items) break;
      const Item it = decode_item(idx);
      if (it.n_tiles == 0) continue;
      const int n_tiles = it.n_tiles;
      SGLB_ATT_DBG(2, n_q, g);
      auto issue_S = [&](uint32_t gt, bool last) {
        const int st = gt & 1;
        SGLB_ATT_WAIT(&k_full[st], (gt >> 1) & 1);
        tc_fence_after();
        const uint32_t k_addr = smem_u32(sK + st * KV_TILE_BYTES);
#pragma unroll
        for (int k = 0; k < ATT_D / 16; ++k) {
          const uint32_t qoff = (k >> 2) * ATT_HALF_BYTES + (k & 3) * 32;
          const uint32_t koff = (k >> 2) * KV_HALF_BYTES + (k & 3) * 32;
          umma_f16_ss(tmem_S0 + st * KT, make_smem_desc_sw128(q_addr + qoff, 0, 1024),
                      make_smem_desc_sw128(k_addr + koff, 0, 1024), idesc_s, k != 0);
        }
        umma_commit(&k_empty[st]);
        umma_commit(&s_full[st]);
        if (last) umma_commit(q_empty);            // the Q tile is dead once the item's last S has been computed
      };
      SGLB_ATT_WAIT(q_full, n_q & 1);
      issue_S(g, n_tiles == 1);
      for (int j = 0; j < n_tiles; ++j) {
        const uint32_t gt = g + j;
        const int st = gt & 1;
        const uint32_t ph = (gt >> 1) & 1;
        if (j + 1 < n_tiles) issue_S(gt + 1, j + 2 == n_tiles);
        SGLB_ATT_WAIT(&p_full[st], ph);
        SGLB_ATT_WAIT(&v_full[st], ph);
        if (j == 0) SGLB_ATT_WAIT(o_free, (n_q & 1) ^ 1);   // previous item's epilogue has read O
        tc_fence_after();
        const uint32_t v_addr = smem_u32(sV + st * KV_TILE_BYTES);
#pragma unroll
        for (int k = 0; k < KT / 16; ++k)
          umma_f16_ts(tmem_O, tmem_S0 + st * KT + k * 8, make_smem_desc_sw128(v_addr + k * 2048, KV_HALF_BYTES, 1024), idesc_o,
                      (j | k) != 0);
        umma_commit(&v_empty[st]);
        umma_commit(o_done);
        if (j + 1 == n_tiles) umma_commit(o_final);
      }
      g += n_tiles;
      ++n_q;
    }
  } else if (warp >= 2) {
    // ------------------------------------------------------------ softmax + epilogue
    const int q = warp & 3;
    const int r = (q << 5) + lane;
    const uint32_t lane_addr = (uint32_t)(q << 5) << 16;
    const float scale_log2 = p.scale_log2;
    const float cap = p.logit_cap, inv_cap = HAS_CAP ? 1.f / p.logit_cap : 0.f;
    const float scale_nat = scale_log2 * 0.6931471805599453f, inv_scale_nat = 1.f / scale_nat;
    uint32_t n_it = 0, n_q = 0, g = 0;
    for (;;) {
      const int slot = n_it & 1;
      SGLB_ATT_WAIT(&item_full[slot], (n_it >> 1) & 1);
      const int idx = s_item[slot];
      __syncwarp();
      if (lane == 0) mbar_arrive(&item_empty[slot]);
      ++n_it;
      if (idx >= p.total_items) break;
      const Item it = decode_item(idx);
      if (!it.valid) continue;
      const int tok = it.t0 + r / G;
      const bool row_valid = tok < it.q_len;
      const bool warp_valid = (it.t0 + (q << 5) / G) < it.q_len;
      if (it.n_tiles == 0) {
        // empty split: publish a neutral partial so the merge kernel can ignore it
        if (p.num_splits > 1 && row_valid) {
          const size_t row = ((size_t)it.split * p.T + it.q_start + tok) * p.Hq + it.kvh * G + (r % G);
          p.part_lse[row] = -INFINITY;
          float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
          float4* dst = reinterpret_cast<float4*>(p.part_o + row * ATT_D);
          for (int i = 0; i < ATT_D / 4; ++i) dst[i] = z;
        }
        continue;
      }
      const int n_tiles = it.n_tiles;
      if (lane == 0) SGLB_ATT_DBG(3, n_q, g);
      const int q_pos = it.prefix + tok;
      float m_i = -INFINITY, l_i = 0.f;
      const int kv_last = p.causal ? min(it.seq_len - 1, q_pos) : it.seq_len - 1;
      const int kv_first = p.window > 0 ? q_pos - p.window + 1 : -(1 << 30);
      for (int j = 0; j < n_tiles; ++j) {
        const uint32_t gt = g + j;
        const int st = gt & 1;
        const uint32_t ph = (gt >> 1) & 1;
        SGLB_ATT_WAIT(&s_full[st], ph);
        tc_fence_after();
        if (warp_valid) {
          constexpr int NC = KT / 32;
          const uint32_t s_addr = tmem_S0 + lane_addr + st * KT;
          uint32_t sr[NC][32];
#pragma unroll
          for (int c = 0; c < NC; ++c) tmem_ld_32x32b_x32(s_addr + c * 32, sr[c]);
          tmem_ld_wait();
          const int kv0 = (it.tile_begin + j) * KT;
          const bool need_mask = (kv0 + KT > it.kv_hi) || (p.causal && kv0 + KT - 1 > it.prefix + it.t0) ||
                                 (p.window > 0 && kv0 < it.prefix + it.t_hi - p.window);
          if (HAS_CAP) {
#pragma unroll
            for (int c = 0; c < NC; ++c)
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                const float x = __uint_as_float(sr[c][i]) * scale_nat;
                sr[c][i] = __float_as_uint(cap * tanhf(x * inv_cap) * inv_scale_nat);
              }
          }
          float mx_raw = -INFINITY;
          if (need_mask) {
            const int hi = kv_last - kv0;
            const int lo = kv_first - kv0;
#pragma unroll
            for (int c = 0; c < NC; ++c) {
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                const int col = c * 32 + i;
                const float v = (col > hi || col < lo) ? -INFINITY : __uint_as_float(sr[c][i]);
                sr[c][i] = __float_as_uint(v);
                mx_raw = fmaxf(mx_raw, v);
              }
            }
          } else {
#pragma unroll
            for (int c = 0; c < NC; ++c) {
#pragma unroll
              for (int i = 0; i < 32; ++i) mx_raw = fmaxf(mx_raw, __uint_as_float(sr[c][i]));
            }
          }
          const float mx = mx_raw * scale_log2;
          const float m_new = fmaxf(m_i, mx);
          const bool first = (j == 0);
          const bool need = (!first) && (m_new > m_i + ATT_RESCALE_THRESHOLD);
          float m_use = m_i;
          if (first) {
            m_use = m_new;
          } else if (__any_sync(0xffffffffu, need)) {
            const float f = (m_new == -INFINITY) ? 1.f : exp2f(m_i - m_new);
            SGLB_ATT_WAIT(o_done, (gt - 1) & 1);
            tc_fence_after();
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
              uint32_t orow[32];
              tmem_ld_32x32b_x32(tmem_O + lane_addr + c * 32, orow);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) orow[i] = __float_as_uint(__uint_as_float(orow[i]) * f);
              tmem_st_32x32b_x32(tmem_O + lane_addr + c * 32, orow);
            }
            tmem_st_wait();
            l_i *= f;
            m_use = m_new;
          }
          m_i = m_use;
          const float m_sub = (m_i == -INFINITY) ? 0.f : m_i;
          float psum = 0.f;
          const uint32_t p_addr = tmem_S0 + lane_addr + st * KT;
#pragma unroll
          for (int c = 0; c < NC; ++c) {
            uint32_t pk[16];
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              const float p0 = ex2_approx(fmaf(__uint_as_float(sr[c][i]), scale_log2, -m_sub));
              const float p1 = ex2_approx(fmaf(__uint_as_float(sr[c][i + 1]), scale_log2, -m_sub));
              psum += p0 + p1;
              pk[i >> 1] = pack_bf16(p0, p1);
            }
            tmem_st_32x32b_x16(p_addr + c * 16, pk);
          }
          l_i += psum;
          tmem_st_wait();
        }
        tc_fence_before();
        mbar_arrive(&p_full[st]);
      }
      g += n_tiles;

      // ---- epilogue: wait for the item's last PV, normalise, store, then hand O back to the MMA warp
      if (lane == 0) SGLB_ATT_DBG(4, n_q, g);
      if (p.legacy_epilogue_wait) SGLB_ATT_WAIT(o_done, (g - 1) & 1);
      else SGLB_ATT_WAIT(o_final, n_q & 1);
      ++n_q;
      tc_fence_after();
      if (warp_valid) {
        const float inv_l = (l_i > 0.f) ? 1.f / l_i : 0.f;
        const float lse = (l_i > 0.f) ? (m_i + log2f(l_i)) * 0.6931471805599453f : -INFINITY;
        const int gq = r % G;
        const size_t out_row = ((size_t)(it.q_start + tok)) * p.Hq + it.kvh * G + gq;
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t orow[32];
          tmem_ld_32x32b_x32(tmem_O + lane_addr + c * 32, orow);
          tmem_ld_wait();
          if (row_valid) {
            if (p.num_splits == 1) {
              __nv_bfloat16* dst = p.out + out_row * ATT_D + c * 32;
#pragma unroll
              for (int v = 0; v < 4; ++v) {
                U4 o;
                o.x = pack_bf16(__uint_as_float(orow[v * 8 + 0]) * inv_l, __uint_as_float(orow[v * 8 + 1]) * inv_l);
                o.y = pack_bf16(__uint_as_float(orow[v * 8 + 2]) * inv_l, __uint_as_float(orow[v * 8 + 3]) * inv_l);
                o.z = pack_bf16(__uint_as_float(orow[v * 8 + 4]) * inv_l, __uint_as_float(orow[v * 8 + 5]) * inv_l);
                o.w = pack_bf16(__uint_as_float(orow[v * 8 + 6]) * inv_l, __uint_as_float(orow[v * 8 + 7]) * inv_l);
                st128(dst + v * 8, o);
              }
            } else {
              float* dst = p.part_o + (((size_t)it.split * p.T + it.q_start + tok) * p.Hq + it.kvh * G + gq) * ATT_D + c * 32;
#pragma unroll
              for (int v = 0; v < 8; ++v) {
                float4 o = make_float4(__uint_as_float(orow[v * 4 + 0]) * inv_l, __uint_as_float(orow[v * 4 + 1]) * inv_l,
                                       __uint_as_float(orow[v * 4 + 2]) * inv_l, __uint_as_float(orow[v * 4 + 3]) * inv_l);
                *reinterpret_cast<float4*>(dst + v * 4) = o;
              }
            }
          }
        }
        if (row_valid) {
          if (p.num_splits == 1) {
            if (p.lse) p.lse[out_row] = lse;
          } else {
            p.part_lse[((size_t)it.split * p.T + it.q_start + tok) * p.Hq + it.kvh * G + gq] = lse;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(o_free);
    }
  }

#undef SGLB_ATT_DBG
#undef SGLB_ATT_WAIT
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
  if (threadIdx.x == 0) {
    __threadfence();
    const int done = atomicAdd(p.sched + 1, 1);
    if (done == (int)gridDim.x - 1) { p.sched[0] = 0; p.sched[1] = 0; __threadfence(); }
  }
}

// merge split-KV partials: one warp per (token, head) row
__global__ void attn_merge_splits_kernel(const float* __restrict__ part_o, const float* __restrict__ part_lse,
                                         __nv_bfloat16* __restrict__ out, float* __restrict__ lse_out, int rows,
                                         int num_splits) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  float m = -INFINITY;
  for (int s = 0; s < num_splits; ++s) m = fmaxf(m, part_lse[(size_t)s * rows + row]);
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  float wsum = 0.f;
  for (int s = 0; s < num_splits; ++s) {
    const float l = part_lse[(size_t)s * rows + row];
    const float w = (l == -INFINITY) ? 0.f : __expf(l - m);
    wsum += w;
    const float4 v = *reinterpret_cast<const float4*>(part_o + ((size_t)s * rows + row) * ATT_D + lane * 4);
    acc[0] += w * v.x; acc[1] += w * v.y; acc[2] += w * v.z; acc[3] += w * v.w;
  }
  const float inv = wsum > 0.f ? 1.f / wsum : 0.f;
  uint2 o;
  o.x = pack_bf16(acc[0] * inv, acc[1] * inv);
  o.y = pack_bf16(acc[2] * inv, acc[3] * inv);
  *reinterpret_cast<uint2*>(out + (size_t)row * ATT_D + lane * 4) = o;
  if (lse_out && lane == 0) lse_out[row] = wsum > 0.f ? m + __logf(wsum) : -INFINITY;
}

void attn_fwd_tc(const void* q, const void* k_cache, const void* v_cache, void* out, float* lse, float* part_o,
                 float* part_lse, const int* page_table, int pt_stride, const int* req_slots, const int* seq_lens,
                 const int* cu_q, int B, int T, int q_stride, int max_q_len, int Hq, int Hkv, int D, int num_slots, int page_size,
                 int num_splits, float scale, int causal, int window, float logit_cap, int* sched, int num_sms,
                 cudaStream_t stream, const unsigned* tree_anc) {
  if (D != ATT_D) throw std::runtime_error("sglb attn_tc: head_dim must be 128");
  const int G = Hq / Hkv;
  if (G < 1 || G > 128 || (128 % G) != 0) throw std::runtime_error("sglb attn_tc: GQA group must divide 128");
  if (page_size < 8 || page_size > 128 || (page_size & (page_size - 1))) throw std::runtime_error("sglb attn_tc: page_size must be 8..128 (pow2)");
  const int TQ = 128 / G;
  // Q viewed as [T][Hkv][G][D]
  uint64_t dq[4] = {(uint64_t)D, (uint64_t)G, (uint64_t)Hkv, (uint64_t)T};
  uint64_t sq[3] = {(uint64_t)D * 2, (uint64_t)G * D * 2, (uint64_t)q_stride * 2};
  uint32_t bq[4] = {64, (uint32_t)G, 1, (uint32_t)TQ};
  CUtensorMap tmQ = make_tmap(q, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, dq, sq, bq, CU_TENSOR_MAP_SWIZZLE_128B);
  uint64_t dk[3] = {(uint64_t)D, (uint64_t)Hkv, (uint64_t)num_slots};
  uint64_t sk[2] = {(uint64_t)D * 2, (uint64_t)Hkv * D * 2};
  uint32_t bk[3] = {64, 1, (uint32_t)page_size};
  CUtensorMap tmK = make_tmap(k_cache, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, dk, sk, bk, CU_TENSOR_MAP_SWIZZLE_128B);
  CUtensorMap tmV = make_tmap(v_cache, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, dk, sk, bk, CU_TENSOR_MAP_SWIZZLE_128B);
  AttnParams p;
  p.out = (__nv_bfloat16*)out; p.part_o = part_o; p.part_lse = part_lse; p.lse = lse;
  p.page_table = page_table; p.req_slots = req_slots; p.seq_lens = seq_lens; p.cu_q = cu_q;
  p.pt_stride = pt_stride; p.Hq = Hq; p.Hkv = Hkv; p.G = G; p.page_size = page_size; p.num_splits = num_splits; p.T = T;
  p.scale_log2 = scale * 1.4426950408889634f; p.logit_cap = logit_cap; p.causal = causal; p.window = window;
  p.tree_anc = tree_anc;
  if (tree_anc != nullptr) {
    if (max_q_len > 32 || num_splits != 1 || window > 0) throw std::runtime_error("sglb attn_tc: tree masks cover <= 32 draft nodes, no split-KV / window");
    p.causal = 0;        // every draft key is a candidate; the ancestor bits decide
  }
  dim3 grid((max_q_len + TQ - 1) / TQ, Hkv * num_splits, B);
  p.gx = (int)grid.x; p.gy = (int)grid.y; p.total_items = (int)(grid.x * grid.y * grid.z); p.sched = sched;
  static const int legacy = [] { const char* e = getenv("SGLB_ATTN_LEGACY_EPILOGUE"); return (e && e[0] == '1') ? 1 : 0; }();
  p.legacy_epilogue_wait = legacy;
  // decode (one query token per request) is latency-bound: small KV tiles, two CTAs per SM
  const bool small = (cu_q == nullptr) && page_size <= 64;
  const bool capped = logit_cap > 0.f;
#define SGLB_ATT_LAUNCH(KT_, CAP_)                                                                                         \
  do {                                                                                                                    \
    static bool configured = false;                                                                                       \
    if (!configured) {                                                                                                    \
      SGLB_CUDA_CHECK(cudaFuncSetAttribute(attn_fwd_tc_kernel<KT_, CAP_>, cudaFuncAttributeMaxDynamicSharedMemorySize,    \
                                           att_smem_bytes<KT_>()));                                                       \
      configured = true;                                                                                                  \
    }                                                                                                                     \
    attn_fwd_tc_kernel<KT_, CAP_><<<grid, ATT_THREADS, att_smem_bytes<KT_>(), stream>>>(tmQ, tmK, tmV, p);                \
  } while (0)
#define SGLB_ATT_LAUNCH_P(KT_, CAP_)                                                                                       \
  do {                                                                                                                    \
    static bool configured = false;                                                                                       \
    if (!configured) {                                                                                                    \
      SGLB_CUDA_CHECK(cudaFuncSetAttribute(attn_fwd_tc_persist_kernel<KT_, CAP_>,                                         \
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, att_smem_bytes<KT_>()));          \
      configured = true;                                                                                                  \
    }                                                                                                                     \
    const int ctas = std::min(p.total_items, num_sms * ((KT_) <= 64 ? 2 : 1));                                            \
    attn_fwd_tc_persist_kernel<KT_, CAP_><<<ctas, ATT_THREADS, att_smem_bytes<KT_>(), stream>>>(tmQ, tmK, tmV, p);        \
  } while (0)
  // persistent scheduling pays off once there are more items than resident CTAs (otherwise the classic launch is identical)
  const bool persist = sched != nullptr && tree_anc == nullptr && p.total_items > num_sms * (small ? 2 : 1);
  if (persist) {
    if (small) { if (capped) SGLB_ATT_LAUNCH_P(64, true); else SGLB_ATT_LAUNCH_P(64, false); }
    else { if (capped) SGLB_ATT_LAUNCH_P(128, true); else SGLB_ATT_LAUNCH_P(128, false); }
  } else {
    if (small) { if (capped) SGLB_ATT_LAUNCH(64, true); else SGLB_ATT_LAUNCH(64, false); }
    else { if (capped) SGLB_ATT_LAUNCH(128, true); else SGLB_ATT_LAUNCH(128, false); }
  }
#undef SGLB_ATT_LAUNCH
#undef SGLB_ATT_LAUNCH_P
  SGLB_CUDA_CHECK(cudaGetLastError());
  if (num_splits > 1) {
    const int rows = T * Hq;
    attn_merge_splits_kernel<<<(rows + 7) / 8, 256, 0, stream>>>(part_o, part_lse, (__nv_bfloat16*)out, lse, rows, num_splits);
    SGLB_CUDA_CHECK(cudaGetLastError());
  }
}

}  // namespace sglb
