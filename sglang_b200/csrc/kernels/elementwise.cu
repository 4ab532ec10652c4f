// Memory-bound elementwise / normalisation kernels for sm_100a (128-bit vector accesses, one pass over HBM).
// Reference counterparts: fused_add_rms_norm_kernel.cu / jit rmsnorm.cuh:26-186, activation.cu:98,
// pos_enc.cu:181 + rope.cuh:177 (RoPE fused with the KV-cache write), kvcache.cuh:189, copy.cu.
#include "common.cuh"
#include "tma_host.h"

namespace sglb {

// ---------------------------------------------------------------- RMSNorm (+ fused residual add)
// one CTA per token row; row cached in registers (H <= 256 threads * MAXV * 8 elements)
template <int MAXV, bool FUSED_ADD, bool GEMMA>
__global__ void __launch_bounds__(256) rmsnorm_kernel(__nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ residual,
                                                      const __nv_bfloat16* __restrict__ w, __nv_bfloat16* __restrict__ out,
                                                      int H, int x_stride, int out_stride, float eps) {
  const int row = blockIdx.x;
  __nv_bfloat16* xr = x + (size_t)row * x_stride;
  __nv_bfloat16* rr = FUSED_ADD ? residual + (size_t)row * H : nullptr;
  __nv_bfloat16* orow = out + (size_t)row * out_stride;
  const int nvec = H >> 3;
  float v[MAXV][8];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int vi = threadIdx.x + i * 256;
    if (vi < nvec) {
      U4 a = ld128(xr + vi * 8);
      float f[8] = {bf16_lo(a.x), bf16_hi(a.x), bf16_lo(a.y), bf16_hi(a.y), bf16_lo(a.z), bf16_hi(a.z), bf16_lo(a.w), bf16_hi(a.w)};
      if (FUSED_ADD) {
        U4 b = ld128(rr + vi * 8);
        float g[8] = {bf16_lo(b.x), bf16_hi(b.x), bf16_lo(b.y), bf16_hi(b.y), bf16_lo(b.z), bf16_hi(b.z), bf16_lo(b.w), bf16_hi(b.w)};
#pragma unroll
        for (int k = 0; k < 8; ++k) f[k] += g[k];
        U4 s = {pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7])};
        st128(rr + vi * 8, s);
        // normalise the bf16-rounded residual (matches the unfused two-kernel numerics)
        f[0] = bf16_lo(s.x); f[1] = bf16_hi(s.x); f[2] = bf16_lo(s.y); f[3] = bf16_hi(s.y);
        f[4] = bf16_lo(s.z); f[5] = bf16_hi(s.z); f[6] = bf16_lo(s.w); f[7] = bf16_hi(s.w);
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) { v[i][k] = f[k]; ss += f[k] * f[k]; }
    }
  }
  __shared__ float red[8];
  ss = warp_sum(ss);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) tot += red[i];
  const float rs = rsqrtf(tot / H + eps);
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int vi = threadIdx.x + i * 256;
    if (vi < nvec) {
      U4 ww = ldg128(w + vi * 8);
      float g[8] = {bf16_lo(ww.x), bf16_hi(ww.x), bf16_lo(ww.y), bf16_hi(ww.y), bf16_lo(ww.z), bf16_hi(ww.z), bf16_lo(ww.w), bf16_hi(ww.w)};
      float o[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) o[k] = v[i][k] * rs * (GEMMA ? (1.f + g[k]) : g[k]);
      st128(orow + vi * 8, U4{pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]), pack_bf16(o[4], o[5]), pack_bf16(o[6], o[7])});
    }
  }
}

void rmsnorm(void* x, void* residual, const void* w, void* out, int T, int H, int x_stride, int out_stride, float eps,
             bool gemma, cudaStream_t stream) {
  if (T <= 0) return;
  if (H % 8 || H > 256 * 8 * 8) throw std::runtime_error("sglb rmsnorm: H must be a multiple of 8 and <= 16384");
  auto X = (__nv_bfloat16*)x; auto R = (__nv_bfloat16*)residual; auto W = (const __nv_bfloat16*)w; auto O = (__nv_bfloat16*)out;
#define LAUNCH(MAXV)                                                                                          \
  do {                                                                                                        \
    if (residual) {                                                                                           \
      if (gemma) rmsnorm_kernel<MAXV, true, true><<<T, 256, 0, stream>>>(X, R, W, O, H, x_stride, out_stride, eps);   \
      else rmsnorm_kernel<MAXV, true, false><<<T, 256, 0, stream>>>(X, R, W, O, H, x_stride, out_stride, eps);        \
    } else {                                                                                                  \
      if (gemma) rmsnorm_kernel<MAXV, false, true><<<T, 256, 0, stream>>>(X, R, W, O, H, x_stride, out_stride, eps);  \
      else rmsnorm_kernel<MAXV, false, false><<<T, 256, 0, stream>>>(X, R, W, O, H, x_stride, out_stride, eps);       \
    }                                                                                                         \
  } while (0)
  const int nvec = H / 8;
  if (nvec <= 256) LAUNCH(1);
  else if (nvec <= 512) LAUNCH(2);
  else if (nvec <= 1024) LAUNCH(4);
  else LAUNCH(8);
#undef LAUNCH
  SGLB_CUDA_CHECK(cudaGetLastError());
}

// ---------------------------------------------------------------- act(x[:, :I]) * x[:, I:]
template <int ACT>  // 0 silu, 1 gelu(erf), 2 gelu(tanh)
__global__ void act_and_mul_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ out, int T, int I,
                                   int x_stride) {
  const int nvec = I >> 3;
  const size_t total = (size_t)T * nvec;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int t = idx / nvec, vi = idx % nvec;
    const __nv_bfloat16* row = x + (size_t)t * x_stride;
    U4 a = ld128(row + vi * 8), b = ld128(row + I + vi * 8);
    float g[8] = {bf16_lo(a.x), bf16_hi(a.x), bf16_lo(a.y), bf16_hi(a.y), bf16_lo(a.z), bf16_hi(a.z), bf16_lo(a.w), bf16_hi(a.w)};
    float u[8] = {bf16_lo(b.x), bf16_hi(b.x), bf16_lo(b.y), bf16_hi(b.y), bf16_lo(b.z), bf16_hi(b.z), bf16_lo(b.w), bf16_hi(b.w)};
    float o[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float a_;
      if (ACT == 0) a_ = g[k] / (1.f + __expf(-g[k]));
      else if (ACT == 1) a_ = 0.5f * g[k] * (1.f + erff(g[k] * 0.7071067811865476f));
      else { const float c = 0.7978845608028654f * (g[k] + 0.044715f * g[k] * g[k] * g[k]); a_ = 0.5f * g[k] * (1.f + tanhf(c)); }
      o[k] = a_ * u[k];
    }
    st128(out + (size_t)t * I + vi * 8, U4{pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]), pack_bf16(o[4], o[5]), pack_bf16(o[6], o[7])});
  }
}

void act_and_mul(const void* x, void* out, int T, int I, int x_stride, int act, cudaStream_t stream) {
  if (T <= 0) return;
  if (I % 8) throw std::runtime_error("sglb act_and_mul: I must be a multiple of 8");
  const size_t total = (size_t)T * (I / 8);
  size_t gq = (total + 255) / 256; const int grid = (int)(gq < (size_t)(148 * 16) ? gq : (size_t)(148 * 16));
  auto X = (const __nv_bfloat16*)x; auto O = (__nv_bfloat16*)out;
  if (act == 0) act_and_mul_kernel<0><<<grid, 256, 0, stream>>>(X, O, T, I, x_stride);
  else if (act == 1) act_and_mul_kernel<1><<<grid, 256, 0, stream>>>(X, O, T, I, x_stride);
  else act_and_mul_kernel<2><<<grid, 256, 0, stream>>>(X, O, T, I, x_stride);
  SGLB_CUDA_CHECK(cudaGetLastError());
}

// ---------------------------------------------------------------- RoPE (neox / gptj) fused with the paged KV write
// one warp per (token, head); head in [0,Hq) -> rotate q in place; [Hq,Hq+Hkv) -> rotate k, write cache;
// [Hq+Hkv, Hq+2Hkv) -> copy v to cache.  D in {64, 128, 256}; rotary_dim <= D.
template <typename CacheT>
__device__ __forceinline__ void store_cache2(CacheT* dst, float a, float b, float inv_scale);
template <>
__device__ __forceinline__ void store_cache2<__nv_bfloat16>(__nv_bfloat16* dst, float a, float b, float) {
  *reinterpret_cast<uint32_t*>(dst) = pack_bf16(a, b);
}
template <>
__device__ __forceinline__ void store_cache2<__nv_fp8_e4m3>(__nv_fp8_e4m3* dst, float a, float b, float inv_scale) {
  __nv_fp8_e4m3 x(a * inv_scale), y(b * inv_scale);
  dst[0] = x; dst[1] = y;
}

// bf16 pair load that also folds split-K partial slabs (gemm_skinny.cu): value = sum_s p[s * slab_stride]
__device__ __forceinline__ void ld2_slabs(const __nv_bfloat16* p, int nslabs, long long slab_stride, float& a, float& b) {
  uint32_t w = *reinterpret_cast<const uint32_t*>(p);
  a = bf16_lo(w); b = bf16_hi(w);
  for (int s = 1; s < nslabs; ++s) {
    w = *reinterpret_cast<const uint32_t*>(p + (size_t)s * slab_stride);
    a += bf16_lo(w); b += bf16_hi(w);
  }
}

template <typename CacheT, bool NEOX>
__global__ void __launch_bounds__(256) rope_kv_write_kernel(__nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ k,
                                                            const __nv_bfloat16* __restrict__ v, const int* __restrict__ positions,
                                                            const float* __restrict__ cos_sin, CacheT* __restrict__ k_cache,
                                                            CacheT* __restrict__ v_cache, const int* __restrict__ out_loc, int T,
                                                            int Hq, int Hkv, int D, int rot, int q_stride, int k_stride,
                                                            int v_stride, float k_inv_scale, float v_inv_scale, int nslabs,
                                                            long long slab_stride) {
  const int heads = Hq + 2 * Hkv;
  const long gw = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (gw >= (long)T * heads) return;
  const int t = gw / heads, h = gw % heads;
  const int lane = threadIdx.x & 31;
  const int half = rot >> 1;
  if (h < Hq + Hkv) {
    const bool is_q = h < Hq;
    const __nv_bfloat16* src = is_q ? q + (size_t)t * q_stride + h * D : k + (size_t)t * k_stride + (h - Hq) * D;
    const int pos = positions[t];
    const float* cs = cos_sin + (size_t)pos * rot;
    CacheT* cdst = is_q ? nullptr : k_cache + ((size_t)out_loc[t] * Hkv + (h - Hq)) * D;
    __nv_bfloat16* qdst = is_q ? q + (size_t)t * q_stride + h * D : nullptr;
    // rotary part: each lane handles pairs
    for (int i = lane * 2; i < half; i += 64) {
      float x1a, x1b, x2a, x2b, c0, c1, s0, s1;
      if (NEOX) {
        ld2_slabs(src + i, nslabs, slab_stride, x1a, x1b);
        ld2_slabs(src + half + i, nslabs, slab_stride, x2a, x2b);
        c0 = cs[i]; c1 = cs[i + 1]; s0 = cs[half + i]; s1 = cs[half + i + 1];
        const float o1a = x1a * c0 - x2a * s0, o1b = x1b * c1 - x2b * s1;
        const float o2a = x2a * c0 + x1a * s0, o2b = x2b * c1 + x1b * s1;
        if (is_q) {
          *reinterpret_cast<uint32_t*>(qdst + i) = pack_bf16(o1a, o1b);
          *reinterpret_cast<uint32_t*>(qdst + half + i) = pack_bf16(o2a, o2b);
        } else {
          store_cache2<CacheT>(cdst + i, o1a, o1b, k_inv_scale);
          store_cache2<CacheT>(cdst + half + i, o2a, o2b, k_inv_scale);
        }
      } else {
        // gptj / interleaved: pairs (2j, 2j+1); this lane handles j = i, i+1
        ld2_slabs(src + 2 * i, nslabs, slab_stride, x1a, x2a);
        ld2_slabs(src + 2 * i + 2, nslabs, slab_stride, x1b, x2b);
        c0 = cs[i]; c1 = cs[i + 1]; s0 = cs[half + i]; s1 = cs[half + i + 1];
        const float oa0 = x1a * c0 - x2a * s0, oa1 = x2a * c0 + x1a * s0;
        const float ob0 = x1b * c1 - x2b * s1, ob1 = x2b * c1 + x1b * s1;
        if (is_q) {
          *reinterpret_cast<uint32_t*>(qdst + 2 * i) = pack_bf16(oa0, oa1);
          *reinterpret_cast<uint32_t*>(qdst + 2 * i + 2) = pack_bf16(ob0, ob1);
        } else {
          store_cache2<CacheT>(cdst + 2 * i, oa0, oa1, k_inv_scale);
          store_cache2<CacheT>(cdst + 2 * i + 2, ob0, ob1, k_inv_scale);
        }
      }
    }
    // pass-through part (rot < D)
    if (!is_q) {
      for (int i = rot + lane * 2; i < D; i += 64) {
        float a, b;
        ld2_slabs(src + i, nslabs, slab_stride, a, b);
        store_cache2<CacheT>(cdst + i, a, b, k_inv_scale);
      }
    } else if (nslabs > 1) {
      for (int i = rot + lane * 2; i < D; i += 64) {
        float a, b;
        ld2_slabs(src + i, nslabs, slab_stride, a, b);
        *reinterpret_cast<uint32_t*>(qdst + i) = pack_bf16(a, b);
      }
    }
  } else {
    const int hv = h - Hq - Hkv;
    const __nv_bfloat16* src = v + (size_t)t * v_stride + hv * D;
    CacheT* cdst = v_cache + ((size_t)out_loc[t] * Hkv + hv) * D;
    for (int i = lane * 2; i < D; i += 64) {
      float a, b;
      ld2_slabs(src + i, nslabs, slab_stride, a, b);
      store_cache2<CacheT>(cdst + i, a, b, v_inv_scale);
    }
  }
}

void rope_kv_write(void* q, const void* k, const void* v, const int* positions, const float* cos_sin, void* k_cache,
                   void* v_cache, const int* out_loc, int T, int Hq, int Hkv, int D, int rot, int q_stride, int k_stride,
                   int v_stride, bool neox, bool fp8_cache, float k_scale, float v_scale, int nslabs, long long slab_stride,
                   cudaStream_t stream) {
  if (T <= 0) return;
  if (rot % 4 || D % 2) throw std::runtime_error("sglb rope: rotary_dim must be a multiple of 4");
  const long warps = (long)T * (Hq + 2 * Hkv);
  const int grid = (int)((warps + 7) / 8);
  auto Q = (__nv_bfloat16*)q; auto K = (const __nv_bfloat16*)k; auto V = (const __nv_bfloat16*)v;
#define RL(CT, NX)                                                                                                  \
  rope_kv_write_kernel<CT, NX><<<grid, 256, 0, stream>>>(Q, K, V, positions, cos_sin, (CT*)k_cache, (CT*)v_cache, out_loc, T, \
                                                         Hq, Hkv, D, rot, q_stride, k_stride, v_stride, 1.f / k_scale, 1.f / v_scale, \
                                                         nslabs, slab_stride)
  if (fp8_cache) { if (neox) RL(__nv_fp8_e4m3, true); else RL(__nv_fp8_e4m3, false); }
  else { if (neox) RL(__nv_bfloat16, true); else RL(__nv_bfloat16, false); }
#undef RL
  SGLB_CUDA_CHECK(cudaGetLastError());
}

// ---------------------------------------------------------------- small index kernels
__global__ void apply_pt_delta_kernel(int* __restrict__ page_table, int stride, const int* __restrict__ rows,
                                      const int* __restrict__ cols, const int* __restrict__ vals, const int* __restrict__ n_ptr,
                                      int n_max) {
  const int n = n_ptr ? min(*n_ptr, n_max) : n_max;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    page_table[(size_t)rows[i] * stride + cols[i]] = vals[i];
}
void apply_pt_delta(int* page_table, int stride, const int* rows, const int* cols, const int* vals, const int* n_ptr,
                    int n_max, cudaStream_t stream) {
  if (n_max <= 0) return;
  apply_pt_delta_kernel<<<min((n_max + 255) / 256, 64), 256, 0, stream>>>(page_table, stride, rows, cols, vals, n_ptr, n_max);
  SGLB_CUDA_CHECK(cudaGetLastError());
}

// out[t, :] = weight[ids[t], :]   with ids optionally gathered through an index (decode: ids = next_token_buf[req_slots])
__global__ void embedding_kernel(const __nv_bfloat16* __restrict__ weight, const int* __restrict__ ids,
                                 const int* __restrict__ gather_idx, __nv_bfloat16* __restrict__ out, int T, int H, int vocab,
                                 int vocab_start) {
  const int t = blockIdx.x;
  int id = gather_idx ? ids[gather_idx[t]] : ids[t];
  id -= vocab_start;
  const bool ok = id >= 0 && id < vocab;
  const __nv_bfloat16* src = weight + (size_t)(ok ? id : 0) * H;
  __nv_bfloat16* dst = out + (size_t)t * H;
  for (int i = threadIdx.x; i < (H >> 3); i += blockDim.x) {
    U4 v = ok ? ldg128(src + i * 8) : U4{0, 0, 0, 0};
    st128(dst + i * 8, v);
  }
}
void embedding(const void* weight, const int* ids, const int* gather_idx, void* out, int T, int H, int vocab, int vocab_start,
               cudaStream_t stream) {
  if (T <= 0) return;
  embedding_kernel<<<T, 128, 0, stream>>>((const __nv_bfloat16*)weight, ids, gather_idx, (__nv_bfloat16*)out, T, H, vocab,
                                          vocab_start);
  SGLB_CUDA_CHECK(cudaGetLastError());
}

// rows gather: out[i, :] = x[idx[i], :]  (last-token hidden-state gather before the LM head)
__global__ void gather_rows_kernel(const __nv_bfloat16* __restrict__ x, const int* __restrict__ idx, __nv_bfloat16* __restrict__ out,
                                   int H, int x_stride) {
  const __nv_bfloat16* src = x + (size_t)idx[blockIdx.x] * x_stride;
  __nv_bfloat16* dst = out + (size_t)blockIdx.x * H;
  for (int i = threadIdx.x; i < (H >> 3); i += blockDim.x) st128(dst + i * 8, ld128(src + i * 8));
}
void gather_rows(const void* x, const int* idx, void* out, int n, int H, int x_stride, cudaStream_t stream) {
  if (n <= 0) return;
  gather_rows_kernel<<<n, 128, 0, stream>>>((const __nv_bfloat16*)x, idx, (__nv_bfloat16*)out, H, x_stride);
  SGLB_CUDA_CHECK(cudaGetLastError());
}

// ---------------------------------------------------------------- reference-grade paged attention (SIMT)
// One warp per (query token, head): lanes split head_dim, keys visited sequentially with an online softmax.
// Purpose: numerics cross-check for attention_tc.cu on the GPU and an emergency path (SGLB_ATTN_IMPL=simt).
template <typename CacheT>
__device__ __forceinline__ float cache_to_float(CacheT v);
template <> __device__ __forceinline__ float cache_to_float<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float cache_to_float<__nv_fp8_e4m3>(__nv_fp8_e4m3 v) { return float(v); }

template <typename CacheT, int D>
__global__ void __launch_bounds__(128) attn_simt_kernel(const __nv_bfloat16* __restrict__ q, const CacheT* __restrict__ k_cache,
                                                        const CacheT* __restrict__ v_cache, __nv_bfloat16* __restrict__ out,
                                                        float* __restrict__ lse_out, const int* __restrict__ page_table,
                                                        int pt_stride, const int* __restrict__ req_slots,
                                                        const int* __restrict__ seq_lens, const int* __restrict__ cu_q,
                                                        const int* __restrict__ tok2req, int T, int q_stride, int Hq, int Hkv,
                                                        int page_size, float scale, int causal, int window, float logit_cap,
                                                        float k_scale, float v_scale) {
  constexpr int PER = D / 32;
  const long gw = (long)blockIdx.x * 4 + (threadIdx.x >> 5);
  if (gw >= (long)T * Hq) return;
  const int t = gw / Hq, h = gw % Hq, lane = threadIdx.x & 31;
  const int b = tok2req ? tok2req[t] : t;
  const int q_start = cu_q ? cu_q[b] : b;
  const int q_len = cu_q ? cu_q[b + 1] - q_start : 1;
  const int seq_len = seq_lens[b];
  const int q_pos = seq_len - q_len + (t - q_start);
  const int kvh = h / (Hq / Hkv);
  const int* pt = page_table + (size_t)req_slots[b] * pt_stride;
  float qv[PER], acc[PER];
#pragma unroll
  for (int i = 0; i < PER; ++i) { qv[i] = __bfloat162float(q[(size_t)t * q_stride + h * D + lane * PER + i]) * scale; acc[i] = 0.f; }
  float m = -INFINITY, l = 0.f;
  const int kv_end = causal ? min(seq_len, q_pos + 1) : seq_len;
  const int kv_begin = window > 0 ? max(0, q_pos - window + 1) : 0;
  for (int kv = kv_begin; kv < kv_end; ++kv) {
    const size_t slot = (size_t)pt[kv / page_size] * page_size + kv % page_size;
    const CacheT* kr = k_cache + (slot * Hkv + kvh) * D + lane * PER;
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < PER; ++i) s += qv[i] * cache_to_float<CacheT>(kr[i]);
    s = warp_sum(s) * k_scale;
    if (logit_cap > 0.f) s = logit_cap * tanhf(s / logit_cap);
    const float m_new = fmaxf(m, s);
    const float f = __expf(m - m_new), pexp = __expf(s - m_new);
    const CacheT* vr = v_cache + (slot * Hkv + kvh) * D + lane * PER;
#pragma unroll
    for (int i = 0; i < PER; ++i) acc[i] = acc[i] * f + pexp * cache_to_float<CacheT>(vr[i]);
    l = l * f + pexp;
    m = m_new;
  }
  const float inv = l > 0.f ? v_scale / l : 0.f;
#pragma unroll
  for (int i = 0; i < PER; ++i) out[((size_t)t * Hq + h) * D + lane * PER + i] = __float2bfloat16(acc[i] * inv);
  if (lse_out && lane == 0) lse_out[(size_t)t * Hq + h] = l > 0.f ? m + __logf(l) : -INFINITY;
}

void attn_fwd_simt(const void* q, const void* k_cache, const void* v_cache, void* out, float* lse, const int* page_table,
                   int pt_stride, const int* req_slots, const int* seq_lens, const int* cu_q, const int* tok2req, int T,
                   int q_stride, int Hq, int Hkv, int D, int page_size, float scale, int causal, int window, float logit_cap,
                   bool fp8_cache, float k_scale, float v_scale, cudaStream_t stream) {
  if (T <= 0) return;
  const long warps = (long)T * Hq;
  const int grid = (int)((warps + 3) / 4);
#define AL(CT, DD)                                                                                                     \
  attn_simt_kernel<CT, DD><<<grid, 128, 0, stream>>>((const __nv_bfloat16*)q, (const CT*)k_cache, (const CT*)v_cache,   \
                                                     (__nv_bfloat16*)out, lse, page_table, pt_stride, req_slots, seq_lens, \
                                                     cu_q, tok2req, T, q_stride, Hq, Hkv, page_size, scale, causal, window, \
                                                     logit_cap, k_scale, v_scale)
  if (fp8_cache) {
    if (D == 128) AL(__nv_fp8_e4m3, 128); else if (D == 64) AL(__nv_fp8_e4m3, 64); else if (D == 256) AL(__nv_fp8_e4m3, 256);
    else throw std::runtime_error("sglb attn_simt: head_dim must be 64/128/256");
  } else {
    if (D == 128) AL(__nv_bfloat16, 128); else if (D == 64) AL(__nv_bfloat16, 64); else if (D == 256) AL(__nv_bfloat16, 256);
    else if (D == 32) AL(__nv_bfloat16, 32);
    else throw std::runtime_error("sglb attn_simt: head_dim must be 32/64/128/256");
  }
#undef AL
  SGLB_CUDA_CHECK(cudaGetLastError());
}

}  // namespace sglb
