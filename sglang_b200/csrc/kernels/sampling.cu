// Sampling kernels: greedy argmax, temperature + top-k / top-p / min-p sampling, grammar bitmask, penalties.
// One CTA (1024 threads) per batch row; the row stays L2-resident across passes (V = 128 K -> 256 KB bf16).
//
// Reference counterparts: python/sglang/srt/layers/sampler.py:98-320 (softmax + FlashInfer
// top_k_top_p_sampling_from_probs / min_p), kernels/aot/csrc/grammar/apply_token_bitmask_inplace_cuda.cu:157,
// sampling/penaltylib.  Design difference: a single fused kernel produces the token (and its logprob) straight
// from logits and scatters it into the per-request-slot `next_token_buf` the next decode step gathers from,
// so sampled ids never round-trip through the host on the critical path.
#include "common.cuh"
#include "tma_host.h"

namespace sglb {

constexpr int SAMP_THREADS = 1024;
constexpr int SAMP_BINS = 2048;

__device__ __forceinline__ float block_reduce_max(float v, float* red) {
  v = warp_max(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : -INFINITY;
  r = warp_max(r);
  r = __shfl_sync(0xffffffffu, r, 0);
  // broadcast from warp 0 to everyone
  __syncthreads();
  if (threadIdx.x == 0) red[0] = r;
  __syncthreads();
  return red[0];
}
__device__ __forceinline__ float block_reduce_sum(float v, float* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.f;
  r = warp_sum(r);
  __syncthreads();
  if (threadIdx.x == 0) red[0] = r;
  __syncthreads();
  return red[0];
}

template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

// counter-based RNG (splitmix64 finaliser over (seed, row, step)) -> uniform in [0,1)
__device__ __forceinline__ float rng_uniform(unsigned long long seed, unsigned long long a, unsigned long long b) {
  unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (a + 1) + 0xD1B54A32D192ED03ull * (b + 1);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  return (float)(z >> 40) * (1.0f / 16777216.0f);
}

struct SampleParams {
  const float* temperature;   // [B]
  const int* top_k;           // [B]  (1 => greedy)
  const float* top_p;         // [B]
  const float* min_p;         // [B]
  const long long* seeds;     // [B, 2] (per-request seed, per-request output position) or null; position < 0 -> this row uses the global stream
  const int* out_slots;       // [B] scatter index into next_token_buf (or null)
  int* next_token_buf;        // [max_reqs+1]
  int* out_tokens;            // [B]
  float* out_logprobs;        // [B] or null
  unsigned long long global_seed;
  const long long* step_ptr;  // device counter (graph-replay safe) or null
  long long step;
  int V, stride;
};

template <typename T>
__global__ void __launch_bounds__(SAMP_THREADS) sample_kernel(const T* __restrict__ logits, SampleParams p) {
  __shared__ float red[32];
  __shared__ float s_chunk[SAMP_THREADS];
  __shared__ int s_arg;
  __shared__ int h_cnt[SAMP_BINS];
  __shared__ float h_mass[SAMP_BINS];
  __shared__ int s_bin;
  __shared__ float s_bc, s_bm;
  const int row = blockIdx.x;
  const T* lr = logits + (size_t)row * p.stride;
  const int V = p.V;
  const int tid = threadIdx.x;
  const int topk = p.top_k[row];
  // ---- pass 1: max (+ argmax)
  float mx = -INFINITY; int am = 0x7fffffff;
  for (int i = tid; i < V; i += SAMP_THREADS) {
    const float v = to_f<T>(lr[i]);
    if (v > mx) { mx = v; am = i; }
  }
  const float gmax = block_reduce_max(mx, red);
  if (tid == 0) s_arg = 0x7fffffff;
  __syncthreads();
  if (mx == gmax) atomicMin(&s_arg, am);   // smallest index among ties (matches torch.argmax)
  __syncthreads();
  const int arg = s_arg;
  const float temp = fmaxf(p.temperature[row], 1e-6f);
  const float inv_t = 1.f / temp;
  // ---- pass 2: partition function at temperature
  float z = 0.f;
  for (int i = tid; i < V; i += SAMP_THREADS) z += __expf((to_f<T>(lr[i]) - gmax) * inv_t);
  const float Z = block_reduce_sum(z, red);
  int token = arg;
  if (topk != 1) {
    const float top_p = p.top_p[row];
    const float min_p = p.min_p[row];
    // threshold in "relative probability" space r = p / p_max in (0, 1]; keep {r >= thr}
    float thr = min_p;   // min-p: p >= min_p * p_max
    const bool need_k = topk > 0 && topk < V;
    const bool need_p = top_p < 1.f;
    // Threshold search by a two-level histogram over d = (max - logit) / T  (r = e^-d): 2048 bins of width 1/64 cover r down to e^-32, a
    // second 2048-bin histogram refines the one bin that holds the boundary (resolution 7.6e-6 in d).  5 passes over the row instead of
    // the 2 x 24 bisection passes of the first version (882 us at 256 x 128 K, profiles/microbench_r1.md).
    float d_max = INFINITY;              // keep { d <= d_max }
    float base_mass = Z;
    auto fill_hist = [&](float d_lo, float inv_w, float d_cap) {
      for (int i = tid; i < SAMP_BINS; i += SAMP_THREADS) { h_cnt[i] = 0; h_mass[i] = 0.f; }
      __syncthreads();
      for (int i = tid; i < V; i += SAMP_THREADS) {
        const float d = (gmax - to_f<T>(lr[i])) * inv_t;
        if (d >= d_lo && d <= d_cap) {
          const int b = min(SAMP_BINS - 1, (int)((d - d_lo) * inv_w));
          atomicAdd(&h_cnt[b], 1);
          atomicAdd(&h_mass[b], __expf(-d));
        }
      }
      __syncthreads();
    };
    // first bin index whose inclusive prefix reaches the target (count or mass); also returns the exclusive prefix before that bin
    auto find_bin = [&](bool by_count, float target, float& before_cnt, float& before_mass) -> int {
      // serial scan by one warp over 2048 bins (64 per lane) — 2048 smem reads, negligible next to a pass over V
      if (tid < 32) {
        float c = 0.f, m = 0.f;
        for (int i = tid * (SAMP_BINS / 32); i < (tid + 1) * (SAMP_BINS / 32); ++i) { c += (float)h_cnt[i]; m += h_mass[i]; }
        float pc = c, pm = m;            // inclusive scan across lanes
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const float tc = __shfl_up_sync(0xffffffffu, pc, o), tm = __shfl_up_sync(0xffffffffu, pm, o);
          if (tid >= o) { pc += tc; pm += tm; }
        }
        const float ec = pc - c, em = pm - m;          // exclusive prefix of this lane's segment
        const bool mine = by_count ? (ec <= target && pc >= target) : (em <= target && pm >= target);      // lowest such lane wins below
        const unsigned ball = __ballot_sync(0xffffffffu, mine);
        if (ball == 0) {
          if (tid == 31) { s_bin = SAMP_BINS; s_bc = pc; s_bm = pm; }      // target beyond the histogram: keep everything in range
        } else if (tid == __ffs(ball) - 1) {
          float rc = ec, rm = em;
          int b = tid * (SAMP_BINS / 32);
          for (;; ++b) {
            const float nc = rc + (float)h_cnt[b], nm = rm + h_mass[b];
            if ((by_count ? nc : nm) >= target) break;
            rc = nc; rm = nm;
          }
          s_bin = b; s_bc = rc; s_bm = rm;
        }
      }
      __syncthreads();
      before_cnt = s_bc; before_mass = s_bm;
      const int b = s_bin;
      __syncthreads();
      return b;
    };
    constexpr float W1 = 1.f / 64.f;                   // level-1 bin width (2048 bins -> d in [0, 32))
    if (need_k || need_p) {
      fill_hist(0.f, 1.f / W1, INFINITY);
      float bc, bm;
      if (need_k) {
        const int b1 = find_bin(true, (float)topk, bc, bm);
        if (b1 < SAMP_BINS) {
          const float lo1 = b1 * W1;
          float c_before = bc, m_before = bm;
          fill_hist(lo1, (float)SAMP_BINS / W1, (b1 == SAMP_BINS - 1) ? INFINITY : lo1 + W1);
          const int b2 = find_bin(true, (float)topk - c_before, bc, bm);
          d_max = (b2 >= SAMP_BINS - 1 && b1 == SAMP_BINS - 1) ? INFINITY : lo1 + (float)(min(b2, SAMP_BINS - 1) + 1) * (W1 / SAMP_BINS);
          base_mass = m_before + bm + ((b2 < SAMP_BINS) ? h_mass[min(b2, SAMP_BINS - 1)] : 0.f);
          __syncthreads();
          if (need_p) fill_hist(0.f, 1.f / W1, d_max);          // level-1 again, restricted to the top-k set
        }
      }
      if (need_p) {
        const float target = top_p * base_mass;
        const int b1 = find_bin(false, target, bc, bm);
        if (b1 < SAMP_BINS) {
          const float lo1 = b1 * W1, m_before = bm;
          fill_hist(lo1, (float)SAMP_BINS / W1, fminf(d_max, (b1 == SAMP_BINS - 1) ? INFINITY : lo1 + W1));
          const int b2 = find_bin(false, target - m_before, bc, bm);
          if (!(b2 >= SAMP_BINS - 1 && b1 == SAMP_BINS - 1)) d_max = fminf(d_max, lo1 + (float)(min(b2, SAMP_BINS - 1) + 1) * (W1 / SAMP_BINS));
        }
      }
    }
    if (d_max < INFINITY) thr = fmaxf(thr, __expf(-d_max));
    // ---- pass 3: kept mass per contiguous chunk, block scan, pick
    const int chunk = (V + SAMP_THREADS - 1) / SAMP_THREADS;
    const int c0 = tid * chunk, c1 = min(V, c0 + chunk);
    float cm = 0.f;
    for (int i = c0; i < c1; ++i) {
      const float r = __expf((to_f<T>(lr[i]) - gmax) * inv_t);
      if (r >= thr) cm += r;
    }
    s_chunk[tid] = cm;
    const float total = block_reduce_sum(cm, red);
    // a seeded request draws from (its seed, ITS OWN output position): the stream does not depend on which step / row / batch the scheduler
    // happened to put it in (the first version mixed the global step counter in, so seeded output changed with scheduling)
    const bool own = p.seeds && p.seeds[2 * row + 1] >= 0;
    const long long step = p.step_ptr ? *p.step_ptr : p.step;
    const float u = (own ? rng_uniform((unsigned long long)p.seeds[2 * row], 0ull, (unsigned long long)p.seeds[2 * row + 1])
                         : rng_uniform(p.global_seed, (unsigned long long)row, (unsigned long long)step)) * total;
    if (tid == 0) {
      float run = 0.f; int pick = SAMP_THREADS - 1;
      for (int t = 0; t < SAMP_THREADS; ++t) {
        if (run + s_chunk[t] > u) { pick = t; break; }
        run += s_chunk[t];
      }
      // walk inside the chunk
      int tok = arg;
      const int a0 = pick * chunk, a1 = min(V, a0 + chunk);
      for (int i = a0; i < a1; ++i) {
        const float r = __expf((to_f<T>(lr[i]) - gmax) * inv_t);
        if (r >= thr) { run += r; tok = i; if (run > u) break; }
      }
      s_arg = tok;
    }
    __syncthreads();
    token = s_arg;
  }
  if (tid == 0) {
    p.out_tokens[row] = token;
    if (p.next_token_buf && p.out_slots) p.next_token_buf[p.out_slots[row]] = token;
    if (p.out_logprobs) p.out_logprobs[row] = (to_f<T>(lr[token]) - gmax) * inv_t - __logf(Z);
  }
}

void sample(const void* logits, bool bf16, int B, int V, int stride, const float* temperature, const int* top_k,
            const float* top_p, const float* min_p, const long long* seeds, const int* out_slots, int* next_token_buf,
            int* out_tokens, float* out_logprobs, unsigned long long global_seed, const long long* step_ptr, long long step,
            cudaStream_t stream) {
  if (B <= 0) return;
  SampleParams p{temperature, top_k, top_p, min_p, seeds, out_slots, next_token_buf, out_tokens, out_logprobs, global_seed,
                 step_ptr, step, V, stride};
  if (bf16) sample_kernel<__nv_bfloat16><<<B, SAMP_THREADS, 0, stream>>>((const __nv_bfloat16*)logits, p);
  else sample_kernel<float><<<B, SAMP_THREADS, 0, stream>>>((const float*)logits, p);
  SGLB_CUDA_CHECK(cudaGetLastError());
}

// ---------------------------------------------------------------- grammar bitmask: logits[b, v] = -inf where bit == 0
// ---- vocab-parallel greedy decoding: no [B, V] logits all-gather on the TP decode path.
// Every rank takes the argmax of its own vocabulary shard and publishes 4 bf16 words per row — (max logit, idx byte 0, 1, 2), all exactly
// representable — into its own [rank] plane of a [tp, Bpad, 4] buffer whose other planes it zeroes; the arena all-reduce (a sum with
// zeros: exact) then is an all-gather of 8 bytes per (rank, row), and greedy_merge_kernel picks the winner (ties -> lowest global index,
// as torch.argmax on the gathered row would).  Reference counterpart: the full-logits all-gather in layers/logits_processor.py:877-905.
template <typename T>
__global__ void __launch_bounds__(SAMP_THREADS) greedy_shard_kernel(const T* __restrict__ logits, int stride, int v_valid, int v_offset,
                                                                    int rank, int tp, int b_pad, __nv_bfloat16* __restrict__ buf) {
  __shared__ float red[32];
  __shared__ int s_arg;
  const int row = blockIdx.x, tid = threadIdx.x;
  const T* lr = logits + (size_t)row * stride;
  float mx = -INFINITY; int am = 0x7fffffff;
  for (int i = tid; i < v_valid; i += SAMP_THREADS) {
    const float v = to_f<T>(lr[i]);
    if (v > mx) { mx = v; am = i; }
  }
  const float gmax = block_reduce_max(mx, red);
  if (tid == 0) s_arg = 0x7fffffff;
  __syncthreads();
  if (mx == gmax && am != 0x7fffffff) atomicMin(&s_arg, am);
  __syncthreads();
  if (tid < tp) {
    __nv_bfloat16* dst = buf + ((size_t)tid * b_pad + row) * 4;
    if (tid == rank) {
      const int g = (s_arg == 0x7fffffff) ? 0 : (s_arg + v_offset);
      // a shard with no valid column (padding past the real vocabulary) publishes the most negative finite bf16
      const float val = (s_arg == 0x7fffffff) ? -3.3e38f : fmaxf(gmax, -3.3e38f);
      dst[0] = __float2bfloat16(val);
      dst[1] = __float2bfloat16((float)(g & 255));
      dst[2] = __float2bfloat16((float)((g >> 8) & 255));
      dst[3] = __float2bfloat16((float)((g >> 16) & 255));
    } else {
      dst[0] = dst[1] = dst[2] = dst[3] = __float2bfloat16(0.f);
    }
  }
}

__global__ void greedy_merge_kernel(const __nv_bfloat16* __restrict__ buf, int tp, int b_pad, int B, const int* __restrict__ out_slots,
                                    int* __restrict__ next_token_buf, int* __restrict__ out_tokens) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= B) return;
  float best = -INFINITY; int tok = 0;
  for (int r = 0; r < tp; ++r) {
    const __nv_bfloat16* src = buf + ((size_t)r * b_pad + row) * 4;
    const float v = __bfloat162float(src[0]);
    if (v > best) {          // strict: equal maxima keep the lower rank = the lower global index
      best = v;
      tok = (int)__bfloat162float(src[1]) | ((int)__bfloat162float(src[2]) << 8) | ((int)__bfloat162float(src[3]) << 16);
    }
  }
  out_tokens[row] = tok;
  if (next_token_buf != nullptr && out_slots != nullptr) next_token_buf[out_slots[row]] = tok;
}

void greedy_shard(const void* logits, bool bf16, int B, int stride, int v_valid, int v_offset, int rank, int tp, int b_pad, void* buf,
                  cudaStream_t stream) {
  if (B <= 0) return;
  if (tp > 32) throw std::runtime_error("sglb greedy_shard: tp <= 32");
  if (bf16)
    greedy_shard_kernel<__nv_bfloat16><<<B, SAMP_THREADS, 0, stream>>>((const __nv_bfloat16*)logits, stride, v_valid, v_offset, rank, tp, b_pad,
                                                                       (__nv_bfloat16*)buf);
  else
    greedy_shard_kernel<float><<<B, SAMP_THREADS, 0, stream>>>((const float*)logits, stride, v_valid, v_offset, rank, tp, b_pad, (__nv_bfloat16*)buf);
  SGLB_CUDA_CHECK(cudaGetLastError());
}

void greedy_merge(const void* buf, int tp, int b_pad, int B, const int* out_slots, int* next_token_buf, int* out_tokens, cudaStream_t stream) {
  if (B <= 0) return;
  greedy_merge_kernel<<<(B + 255) / 256, 256, 0, stream>>>((const __nv_bfloat16*)buf, tp, b_pad, B, out_slots, next_token_buf, out_tokens);
  SGLB_CUDA_CHECK(cudaGetLastError());
}

template <typename T>
__global__ void apply_bitmask_kernel(T* __restrict__ logits, const int* __restrict__ bitmask, const int* __restrict__ row_idx,
                                     int V, int stride, int words) {
  const int r = blockIdx.y;
  const int row = row_idx ? row_idx[r] : r;
  T* lr = logits + (size_t)row * stride;
  const int* bm = bitmask + (size_t)r * words;
  for (int w = blockIdx.x * blockDim.x + threadIdx.x; w < words; w += gridDim.x * blockDim.x) {
    const unsigned bits = (unsigned)bm[w];
    if (bits == 0xffffffffu) continue;
#pragma unroll 4
    for (int b = 0; b < 32; ++b) {
      const int v = w * 32 + b;
      if (v < V && !((bits >> b) & 1u)) lr[v] = T(-INFINITY);
    }
  }
}
void apply_token_bitmask(void* logits, bool bf16, const int* bitmask, const int* row_idx, int n_rows, int V, int stride,
                         cudaStream_t stream) {
  if (n_rows <= 0) return;
  const int words = (V + 31) / 32;
  dim3 grid(min((words + 255) / 256, 16), n_rows);
  if (bf16) apply_bitmask_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>((__nv_bfloat16*)logits, bitmask, row_idx, V, stride, words);
  else apply_bitmask_kernel<float><<<grid, 256, 0, stream>>>((float*)logits, bitmask, row_idx, V, stride, words);
  SGLB_CUDA_CHECK(cudaGetLastError());
}

// ---------------------------------------------------------------- penalties (frequency / presence / repetition) + logit bias
// counts: int32 [B, V] occurrences in the generated output; prompt_mask: uint8 [B, V] (optional)
__global__ void apply_penalties_kernel(float* __restrict__ logits, const int* __restrict__ counts,
                                       const unsigned char* __restrict__ prompt_mask, const float* __restrict__ freq,
                                       const float* __restrict__ pres, const float* __restrict__ rep, int V, int stride) {
  const int row = blockIdx.y;
  float* lr = logits + (size_t)row * stride;
  const int* cr = counts + (size_t)row * V;
  const float f = freq[row], pz = pres[row], rp = rep[row];
  for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < V; v += gridDim.x * blockDim.x) {
    const int c = cr[v];
    const bool seen = c > 0 || (prompt_mask && prompt_mask[(size_t)row * V + v]);
    if (!seen) continue;
    float x = lr[v];
    if (rp != 1.f) x = x > 0.f ? x / rp : x * rp;
    x -= f * (float)c + (c > 0 ? pz : 0.f);
    lr[v] = x;
  }
}
void apply_penalties(float* logits, const int* counts, const unsigned char* prompt_mask, const float* freq, const float* pres,
                     const float* rep, int B, int V, int stride, cudaStream_t stream) {
  if (B <= 0) return;
  dim3 grid(min((V + 255) / 256, 64), B);
  apply_penalties_kernel<<<grid, 256, 0, stream>>>(logits, counts, prompt_mask, freq, pres, rep, V, stride);
  SGLB_CUDA_CHECK(cudaGetLastError());
}

__global__ void bump_counts_kernel(int* __restrict__ counts, const int* __restrict__ tokens, int V, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B) atomicAdd(&counts[(size_t)b * V + tokens[b]], 1);
}
void bump_counts(int* counts, const int* tokens, int B, int V, cudaStream_t stream) {
  if (B <= 0) return;
  bump_counts_kernel<<<(B + 127) / 128, 128, 0, stream>>>(counts, tokens, V, B);
  SGLB_CUDA_CHECK(cudaGetLastError());
}

}  // namespace sglb
