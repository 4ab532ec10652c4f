// State-space (Mamba / Mamba-2) kernels: causal depthwise conv1d over ragged batches with a per-request rolling state, and the
// single-token selective state update of decode.
//
//   causal_conv1d (prefill / extend / decode):  u[t, c] = act(bias[c] + sum_{j<k} w[c, j] * x[t - (k-1) + j, c]),   x[<0] taken from conv_state
//        conv_state[slot, c, :] <- the last k-1 inputs of the request (shifted by the new tokens)
//   selective_state_update (decode):            S[b,h,p,:] = S * exp(dt[b,h] * A[h]) + dt[b,h] * x[b,h,p] * B[b,g,:]
//                                               y[b,h,p]   = <S[b,h,p,:], C[b,g,:]> + D[h] * x[b,h,p]
//
// Both are bandwidth kernels: the conv reads x once (threads = channels: coalesced rows), the state update reads and writes the fp32 state
// exactly once (one warp per (b, h, p) row, 16-byte lanes) instead of the 4–5 passes of the PyTorch expression.
//
// Reference counterparts: kernels/aot/csrc/mamba/causal_conv1d.cu (fwd + update), layers/attention/mamba/ops/mamba_ssm.py
// (selective_state_update Triton kernel); states indexed by request slot as in mem_cache/memory_pool.py MambaPool.
#include "common.cuh"
#include "tma_host.h"

namespace sglb {

constexpr int CONV_MAX_K = 8;
constexpr int CONV_CHUNK = 128;      // tokens per block along a request

template <typename T> __device__ __forceinline__ float ld_f(const T* p);
template <> __device__ __forceinline__ float ld_f<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float ld_f<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
template <typename T> __device__ __forceinline__ void st_f(T* p, float v);
template <> __device__ __forceinline__ void st_f<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void st_f<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16(v); }

// grid (ceil(C / 256), chunks, B); thread = channel.  cu_q == nullptr: decode (one token per request, row = request index).
template <typename T>
__global__ void __launch_bounds__(256) causal_conv1d_fwd_kernel(const T* __restrict__ x, int ldx, const T* __restrict__ w, const T* __restrict__ bias,
                                                                const T* __restrict__ conv_state, const int* __restrict__ cu_q,
                                                                const int* __restrict__ req_slots, const int* __restrict__ seq_lens, int C, int K,
                                                                int silu, float* __restrict__ u, int ldu) {
  const int c = blockIdx.x * 256 + threadIdx.x;
  const int b = blockIdx.z;
  if (c >= C) return;
  const int q0 = cu_q ? cu_q[b] : b;
  const int n = cu_q ? cu_q[b + 1] - q0 : 1;
  const int t0 = blockIdx.y * CONV_CHUNK;
  if (t0 >= n) return;
  const int t1 = min(n, t0 + CONV_CHUNK);
  const bool fresh = seq_lens[b] - n == 0;            // first tokens of the request: the state is all zeros, whatever the slot holds
  const T* st = conv_state + ((size_t)req_slots[b] * C + c) * (K - 1);
  float wk[CONV_MAX_K], win[CONV_MAX_K];
#pragma unroll
  for (int j = 0; j < CONV_MAX_K; ++j) wk[j] = j < K ? ld_f<T>(w + (size_t)c * K + j) : 0.f;
  const float bs = bias ? ld_f<T>(bias + c) : 0.f;
  float w_last = 0.f;
#pragma unroll
  for (int j = 0; j < CONV_MAX_K; ++j)
    if (j == K - 1) w_last = wk[j];
  // window = inputs t0-(K-1) .. t0-1: from x when inside this call's tokens, else from the rolling state (position relative to its end)
#pragma unroll
  for (int j = 0; j < CONV_MAX_K - 1; ++j) {
    if (j < K - 1) {
      const int t = t0 - (K - 1) + j;
      float v = 0.f;
      if (t >= 0) v = ld_f<T>(x + (size_t)(q0 + t) * ldx + c);
      else if (!fresh) v = ld_f<T>(st + (K - 1) + t);
      win[j] = v;
    }
  }
  for (int t = t0; t < t1; ++t) {
    const float xv = ld_f<T>(x + (size_t)(q0 + t) * ldx + c);
    float acc = bs + w_last * xv;
#pragma unroll
    for (int j = 0; j < CONV_MAX_K - 1; ++j)
      if (j < K - 1) acc += wk[j] * win[j];
#pragma unroll
    for (int j = 0; j < CONV_MAX_K - 1; ++j) {
      if (j < K - 2) win[j] = win[j + 1];
      else if (j == K - 2) win[j] = xv;
    }
    if (silu) acc = acc / (1.f + __expf(-acc));
    u[(size_t)(q0 + t) * ldu + c] = acc;
  }
}

// second pass (stream-ordered after the forward, which still reads the OLD state): roll the state forward by the request's new tokens
template <typename T>
__global__ void __launch_bounds__(256) causal_conv1d_state_kernel(const T* __restrict__ x, int ldx, T* __restrict__ conv_state, const int* __restrict__ cu_q,
                                                                  const int* __restrict__ req_slots, const int* __restrict__ seq_lens, int C, int K) {
  const int c = blockIdx.x * 256 + threadIdx.x;
  const int b = blockIdx.y;
  if (c >= C) return;
  const int q0 = cu_q ? cu_q[b] : b;
  const int n = cu_q ? cu_q[b + 1] - q0 : 1;
  if (n <= 0) return;
  const bool fresh = seq_lens[b] - n == 0;
  T* st = conv_state + ((size_t)req_slots[b] * C + c) * (K - 1);
  float nw[CONV_MAX_K];
#pragma unroll
  for (int j = 0; j < CONV_MAX_K - 1; ++j) {
    if (j < K - 1) {
      const int t = n - (K - 1) + j;                 // input index (relative to this call) that ends up in state column j
      float v = 0.f;
      if (t >= 0) v = ld_f<T>(x + (size_t)(q0 + t) * ldx + c);
      else if (!fresh) v = ld_f<T>(st + (K - 1) + t);
      nw[j] = v;
    }
  }
#pragma unroll
  for (int j = 0; j < CONV_MAX_K - 1; ++j)
    if (j < K - 1) st_f<T>(st + j, nw[j]);
}

void causal_conv1d(const void* x, int ldx, const void* w, const void* bias, void* conv_state, const int* cu_q, const int* req_slots,
                   const int* seq_lens, int B, int max_q_len, int C, int K, bool silu, bool bf16, float* u, int ldu, cudaStream_t stream) {
  if (B <= 0) return;
  if (K < 2 || K > CONV_MAX_K) throw std::runtime_error("sglb causal_conv1d: kernel width must be 2..8");
  dim3 g1((C + 255) / 256, (std::max(max_q_len, 1) + CONV_CHUNK - 1) / CONV_CHUNK, B), g2((C + 255) / 256, B);
  if (bf16) {
    causal_conv1d_fwd_kernel<__nv_bfloat16><<<g1, 256, 0, stream>>>((const __nv_bfloat16*)x, ldx, (const __nv_bfloat16*)w, (const __nv_bfloat16*)bias,
                                                                    (const __nv_bfloat16*)conv_state, cu_q, req_slots, seq_lens, C, K, silu ? 1 : 0, u, ldu);
    causal_conv1d_state_kernel<__nv_bfloat16><<<g2, 256, 0, stream>>>((const __nv_bfloat16*)x, ldx, (__nv_bfloat16*)conv_state, cu_q, req_slots, seq_lens, C, K);
  } else {
    causal_conv1d_fwd_kernel<float><<<g1, 256, 0, stream>>>((const float*)x, ldx, (const float*)w, (const float*)bias, (const float*)conv_state, cu_q,
                                                            req_slots, seq_lens, C, K, silu ? 1 : 0, u, ldu);
    causal_conv1d_state_kernel<float><<<g2, 256, 0, stream>>>((const float*)x, ldx, (float*)conv_state, cu_q, req_slots, seq_lens, C, K);
  }
  SGLB_CUDA_CHECK(cudaGetLastError());
}

// one warp per (b, h, p) state row of N floats; x / B / C / dt fp32 (already activated), state fp32 [slots, H, P, N]
__global__ void __launch_bounds__(256) selective_state_update_kernel(float* __restrict__ state, const int* __restrict__ req_slots,
                                                                     const float* __restrict__ x, int ldx, const float* __restrict__ dt, const float* __restrict__ A,
                                                                     const float* __restrict__ Bm, const float* __restrict__ Cm, int ldbc,
                                                                     const float* __restrict__ Dv, int B, int H, int P, int N, int G, float* __restrict__ y) {
  const int warp = (blockIdx.x * 256 + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int rows = B * H * P;
  if (warp >= rows) return;
  const int b = warp / (H * P), hp = warp - b * H * P, h = hp / P, p = hp - h * P;
  const int g = h / (H / G);
  const float dtv = dt[(size_t)b * H + h];
  const float dA = __expf(dtv * A[h]);
  const float xv = x[(size_t)b * ldx + h * P + p];
  const float dbx = dtv * xv;
  float* s = state + (((size_t)req_slots[b] * H + h) * P + p) * N;
  const float* bb = Bm + (size_t)b * ldbc + (size_t)g * N;
  const float* cc = Cm + (size_t)b * ldbc + (size_t)g * N;
  float acc = 0.f;
  for (int i = lane * 4; i < N; i += 128) {
    float4 sv = *reinterpret_cast<const float4*>(s + i);
    const float4 bv = *reinterpret_cast<const float4*>(bb + i);
    const float4 cv = *reinterpret_cast<const float4*>(cc + i);
    sv.x = sv.x * dA + dbx * bv.x; sv.y = sv.y * dA + dbx * bv.y; sv.z = sv.z * dA + dbx * bv.z; sv.w = sv.w * dA + dbx * bv.w;
    *reinterpret_cast<float4*>(s + i) = sv;
    acc += sv.x * cv.x + sv.y * cv.y + sv.z * cv.z + sv.w * cv.w;
  }
  acc = warp_sum(acc);
  if (lane == 0) y[(size_t)b * H * P + hp] = acc + (Dv ? Dv[h] * xv : 0.f);
}

void selective_state_update(float* state, const int* req_slots, const float* x, int ldx, const float* dt, const float* A, const float* Bm,
                            const float* Cm, int ldbc, const float* Dv, int B, int H, int P, int N, int G, float* y, cudaStream_t stream) {
  if (B <= 0) return;
  if (N % 4 || H % G) throw std::runtime_error("sglb selective_state_update: d_state must be a multiple of 4, heads a multiple of groups");
  const long long warps = (long long)B * H * P;
  selective_state_update_kernel<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, stream>>>(state, req_slots, x, ldx, dt, A, Bm, Cm, ldbc, Dv, B, H, P, N, G, y);
  SGLB_CUDA_CHECK(cudaGetLastError());
}

// ---------------------------------------------------------------- gated delta rule, one token per request (Qwen3-Next / GDN linear attention)
//   S <- S * exp(g);  mem = S^T k;  delta = (v - mem) * beta;  S <- S + k delta^T;  o = S^T q        (S fp32 [Dk, Dv] per (request, value head))
// q and k arrive un-normalised from the conv (+SiLU) output: l2-normalised here, q also scaled by Dk^-1/2; value head h uses key head h / r.
// One block per (value head, request), thread = value column: the Dk-long column lives in registers between the two dependent passes,
// so the 64 KB state is read once and written once (the PyTorch form makes five passes).  Reference: fla / sglang fused_recurrent_gated_delta_rule.
// Threads = (DK / 64) key slices x Dv value columns: a thread keeps 64 state values of one column in registers (128 would spill).
template <int DK>
__global__ void __launch_bounds__(256) gated_delta_update_kernel(float* __restrict__ state, const int* __restrict__ req_slots, const float* __restrict__ u,
                                                                 int ldu, int k_off, int v_off, const float* __restrict__ g, const float* __restrict__ beta,
                                                                 int Hv, int r, int Dv, int dv_pad, float* __restrict__ out) {
  constexpr int NS = DK / 64;
  __shared__ float sq[DK], sk[DK], red[32], part[NS][256];
  const int h = blockIdx.x, b = blockIdx.y, hk = h / r;
  const float* row = u + (size_t)b * ldu;
  float qs = 0.f, ks = 0.f;
  for (int i = threadIdx.x; i < DK; i += blockDim.x) {
    const float qv = row[hk * DK + i], kv = row[k_off + hk * DK + i];
    sq[i] = qv; sk[i] = kv;
    qs += qv * qv; ks += kv * kv;
  }
  qs = warp_sum(qs); ks = warp_sum(ks);
  const int w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if ((threadIdx.x & 31) == 0) { red[w] = qs; red[16 + w] = ks; }
  __syncthreads();
  float tq = 0.f, tk = 0.f;
  for (int i = 0; i < nw; ++i) { tq += red[i]; tk += red[16 + i]; }
  const float qn = rsqrtf(tq + 1e-6f) * rsqrtf((float)DK), kn = rsqrtf(tk + 1e-6f);
  const int slice = threadIdx.x / dv_pad, v = threadIdx.x - slice * dv_pad;
  const bool live = v < Dv;
  const float decay = __expf(g[(size_t)b * Hv + h]);
  float* S = state + (((size_t)req_slots[b] * Hv + h) * DK + slice * 64) * Dv + v;
  float col[64];
  float mem = 0.f;
  if (live) {
#pragma unroll
    for (int k = 0; k < 64; ++k) {
      col[k] = S[(size_t)k * Dv] * decay;
      mem += col[k] * (sk[slice * 64 + k] * kn);
    }
  }
  part[slice][v] = mem;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int i = 0; i < NS; ++i) tot += part[i][v];
  __syncthreads();
  float o = 0.f;
  if (live) {
    const float delta = (row[v_off + h * Dv + v] - tot) * beta[(size_t)b * Hv + h];
#pragma unroll
    for (int k = 0; k < 64; ++k) {
      const float s2 = col[k] + (sk[slice * 64 + k] * kn) * delta;
      S[(size_t)k * Dv] = s2;
      o += s2 * (sq[slice * 64 + k] * qn);
    }
  }
  part[slice][v] = o;
  __syncthreads();
  if (slice == 0 && live) {
    float oo = 0.f;
#pragma unroll
    for (int i = 0; i < NS; ++i) oo += part[i][v];
    out[((size_t)b * Hv + h) * Dv + v] = oo;
  }
}

bool gated_delta_update(float* state, const int* req_slots, const float* u, int ldu, int k_off, int v_off, const float* g, const float* beta, int B,
                        int Hv, int r, int Dk, int Dv, float* out, cudaStream_t stream) {
  if (B <= 0) return true;
  if ((Dk != 64 && Dk != 128) || Dv > (Dk == 64 ? 256 : 128)) return false;      // <= 256 threads per block
  dim3 grid(Hv, B);
  const int dv_pad = ((Dv + 31) / 32) * 32;
  if (Dk == 64) gated_delta_update_kernel<64><<<grid, dv_pad, 0, stream>>>(state, req_slots, u, ldu, k_off, v_off, g, beta, Hv, r, Dv, dv_pad, out);
  else gated_delta_update_kernel<128><<<grid, 2 * dv_pad, 0, stream>>>(state, req_slots, u, ldu, k_off, v_off, g, beta, Hv, r, Dv, dv_pad, out);
  SGLB_CUDA_CHECK(cudaGetLastError());
  return true;
}

}  // namespace sglb
