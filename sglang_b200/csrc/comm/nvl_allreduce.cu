// NVLink-5 / NVSwitch collectives over a symmetric-memory arena (peer pointers + optional NVLS multicast alias).
//
//   * one-shot all-reduce (pull): every rank publishes its vector in its own arena slot, one flag round, then reads
//     all peers over NVLink and sums in fp32 — latency-optimal for decode-size messages.
//   * two-shot all-reduce: reduce-scatter by ownership (rank r reduces slice r from all peers) + all-gather of the
//     reduced slices — bandwidth-optimal for prefill-size messages; NVLS variant uses multimem.ld_reduce / multimem.st
//     so the switch does the reduction and the broadcast.
//   * FUSED all-reduce + residual add + RMSNorm: the consumer of a row-parallel GEMM sums the peers' partial rows,
//     adds the residual, normalises and writes the next layer's input in one pass (no standalone all-reduce kernel,
//     no extra HBM round trip of the [T,H] activations).
//
// Flag protocol: each rank owns a signal pad; `flags[channel][src_rank]` holds the last epoch published by src_rank.
// Epochs increase monotonically and live in device memory, so CUDA-graph replays need no host involvement.
// Arena slots are double-buffered by epoch parity (see sglang_b200/parallel/nvl_comm.py for the hazard argument).
//
// Reference counterparts: kernels/jit/csrc/distributed/custom_all_reduce.cuh:160-205,486 (1shot push/pull, 2shot,
// multimem), kimi_k3/comm/ar_fusion.cuh:1-32 (AR + residual + RMSNorm), layers/flashinfer_comm_fusion.py:775-867.
#include "../kernels/common.cuh"
#include "../kernels/tma_host.h"

namespace sglb {

constexpr int NVL_MAX_RANKS = 8;

struct NvlPeers {
  const void* buf[NVL_MAX_RANKS];    // peer arena base (this call's slot offset already applied)
  uint32_t* flags[NVL_MAX_RANKS];    // peer signal pads (uint32 flags[channel][NVL_MAX_RANKS])
  void* mc_buf;                      // multicast alias of the arena slot (nullptr if NVLS unavailable)
  int rank, world;
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ U4 ld_peer128(const void* p) {
  U4 r;
  asm volatile("ld.global.relaxed.sys.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
  return r;
}

// NVLS (NVSwitch multicast) primitives: one load returns the SUM over every rank's copy (the switch reduces in fp32), one store
// lands in every rank's copy.  `p` is an address inside the multicast alias of the symmetric arena.
__device__ __forceinline__ U4 mc_ld_reduce_bf16x8(const void* p) {
  U4 r;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.bf16x2 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
  return r;
}
__device__ __forceinline__ void mc_st_bf16x8(void* p, const U4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.bf16x2 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// Wall-clock bound for every cross-rank wait: ranks of an SPMD job may legitimately skew by seconds at trace boundaries (host-side work),
// so the bound is TIME (globaltimer, 45 s), not a spin count; a protocol bug still ends in a trap with a message instead of a hung box.
__device__ __forceinline__ bool nvl_wait_expired(uint32_t& spins, unsigned long long& t0) {
  if ((++spins & 0xfffu) != 0) return false;
  unsigned long long now;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
  if (t0 == 0) { t0 = now; return false; }
  return now - t0 > 45ull * 1000ull * 1000ull * 1000ull;
}

__device__ __forceinline__ uint32_t ld_relaxed_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void fence_acquire_sys() { asm volatile("fence.acq_rel.sys;" ::: "memory"); }

// Flag round: publish `epoch` to every peer on `channel`, then wait until every peer has published it to us.
// Called by all threads of every block.  Only block 0 PUBLISHES (the data being announced was produced by an earlier kernel on this
// stream, so one release store per peer is enough — measured: with every CTA issuing sys-scope fences + remote stores the flag
// round cost ~7 us on 2 x B200); every block polls its own rank's (local) flag words with relaxed loads and fences once.
__device__ __forceinline__ void nvl_barrier(const NvlPeers& P, int channel, uint32_t epoch) {
  __syncthreads();
  if (threadIdx.x < P.world) {
    if (blockIdx.x == 0) st_release_sys(P.flags[threadIdx.x] + channel * NVL_MAX_RANKS + P.rank, epoch);
    const uint32_t* mine = P.flags[P.rank] + channel * NVL_MAX_RANKS + threadIdx.x;
    uint32_t spins = 0;
    unsigned long long t0 = 0;
    while ((int32_t)(ld_relaxed_sys(mine) - epoch) < 0) {
      if (nvl_wait_expired(spins, t0)) {
        printf("sglb nvl: barrier timeout (45 s) rank=%d block=%d waiting on rank=%d channel=%d epoch=%u have=%u\n", P.rank, blockIdx.x,
               threadIdx.x, channel, epoch, ld_acquire_sys(mine));
        __trap();
      }
    }
    fence_acquire_sys();
  }
  __syncthreads();
}

__device__ __forceinline__ void acc_bf16x8(float (&a)[8], const U4& v) {
  a[0] += bf16_lo(v.x); a[1] += bf16_hi(v.x); a[2] += bf16_lo(v.y); a[3] += bf16_hi(v.y);
  a[4] += bf16_lo(v.z); a[5] += bf16_hi(v.z); a[6] += bf16_lo(v.w); a[7] += bf16_hi(v.w);
}
__device__ __forceinline__ U4 pack8(const float (&a)[8]) {
  return U4{pack_bf16(a[0], a[1]), pack_bf16(a[2], a[3]), pack_bf16(a[4], a[5]), pack_bf16(a[6], a[7])};
}

// epoch bookkeeping: every block reads *epoch_ptr at entry; the last block to leave bumps it (graph-replay safe)
__device__ __forceinline__ uint32_t epoch_begin(const uint32_t* epoch_ptr) { return *((volatile const uint32_t*)epoch_ptr) + 1; }
__device__ __forceinline__ void epoch_end(uint32_t* epoch_ptr, uint32_t* done_ctr, uint32_t epoch) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const uint32_t t = atomicAdd(done_ctr, 1);
    if (t == gridDim.x - 1) {
      *done_ctr = 0;
      __threadfence();
      *((volatile uint32_t*)epoch_ptr) = epoch;
    }
  }
}

// ---------------------------------------------------------------- one-shot all-reduce (bf16, n multiple of 8)
// WEAK: plain (weak) 128-bit loads after the acquire of the flag round instead of relaxed.sys loads — this kernel never read the
// peers' slots before the barrier, so no stale L1 line can exist; weak loads pipeline better over NVLink.
template <bool WEAK>
__global__ void __launch_bounds__(512) nvl_allreduce_oneshot_kernel(NvlPeers P, __nv_bfloat16* __restrict__ out, int n,
                                                                    uint32_t* epoch_ptr, uint32_t* done_ctr, int channel) {
  const uint32_t epoch = epoch_begin(epoch_ptr);
  nvl_barrier(P, channel, epoch);
  const int nvec = n >> 3;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += gridDim.x * blockDim.x) {
    float a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    U4 v[NVL_MAX_RANKS];
#pragma unroll
    for (int r = 0; r < NVL_MAX_RANKS; ++r)
      if (r < P.world) {
        const __nv_bfloat16* src = reinterpret_cast<const __nv_bfloat16*>(P.buf[r]) + (size_t)i * 8;
        v[r] = WEAK ? ld128(src) : ld_peer128(src);
      }
#pragma unroll
    for (int r = 0; r < NVL_MAX_RANKS; ++r)
      if (r < P.world) acc_bf16x8(a, v[r]);
    st128(out + (size_t)i * 8, pack8(a));
  }
  epoch_end(epoch_ptr, done_ctr, epoch);
}

// ---------------------------------------------------------------- two-shot all-reduce (reduce-scatter + all-gather)
// The reduced slice is written back into every peer's arena slot (push all-gather); a second flag round then
// releases the result, which each rank copies out of its own slot.
__global__ void __launch_bounds__(512) nvl_allreduce_twoshot_kernel(NvlPeers P, __nv_bfloat16* __restrict__ out, int n,
                                                                    uint32_t* epoch_ptr, uint32_t* done_ctr, int channel) {
  const uint32_t epoch = epoch_begin(epoch_ptr);
  nvl_barrier(P, channel, 2 * epoch - 1);
  const int nvec = n >> 3;
  const int per = (nvec + P.world - 1) / P.world;
  const int lo = P.rank * per, hi = min(nvec, lo + per);
  for (int i = lo + blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += gridDim.x * blockDim.x) {
    float a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int r = 0; r < NVL_MAX_RANKS; ++r)
      if (r < P.world) acc_bf16x8(a, ld_peer128(reinterpret_cast<const __nv_bfloat16*>(P.buf[r]) + (size_t)i * 8));
    const U4 v = pack8(a);
    // result region lives after the input region inside each slot: [n inputs | n results]
#pragma unroll
    for (int r = 0; r < NVL_MAX_RANKS; ++r)
      if (r < P.world)
        st128(const_cast<__nv_bfloat16*>(reinterpret_cast<const __nv_bfloat16*>(P.buf[r])) + (size_t)n + (size_t)i * 8, v);
  }
  // all writers of this grid must be done before the flag goes out: grid-wide completion via the done counter
  __syncthreads();
  __shared__ uint32_t s_last;
  if (threadIdx.x == 0) {
    __threadfence_system();
    s_last = (atomicAdd(done_ctr + 1, 1) == gridDim.x - 1);
  }
  __syncthreads();
  // every block waits until the whole grid has pushed (spin on the local counter), then the 2nd flag round
  if (threadIdx.x == 0) {
    uint32_t spins = 0;
    unsigned long long t0 = 0;
    while (*((volatile uint32_t*)(done_ctr + 1)) < gridDim.x) {
      if (nvl_wait_expired(spins, t0)) { printf("sglb nvl: two-shot grid counter timeout rank=%d block=%d\n", P.rank, blockIdx.x); __trap(); }
    }
  }
  nvl_barrier(P, channel, 2 * epoch);
  const __nv_bfloat16* res = reinterpret_cast<const __nv_bfloat16*>(P.buf[P.rank]) + (size_t)n;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += gridDim.x * blockDim.x)
    st128(out + (size_t)i * 8, ld128(res + (size_t)i * 8));
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const uint32_t t = atomicAdd(done_ctr, 1);
    if (t == gridDim.x - 1) {
      *done_ctr = 0;
      *(done_ctr + 1) = 0;
      __threadfence();
      *((volatile uint32_t*)epoch_ptr) = epoch;
    }
  }
}

// ---------------------------------------------------------------- NVLS one-shot: out[i] = multimem.ld_reduce(arena[i])
// One flag round, then every rank pulls the already-reduced vector through the switch: inbound traffic per rank is n bytes
// (not (world-1) * n as in the peer-pull one-shot) and there is ONE load instruction per 16 bytes regardless of world size.
__global__ void __launch_bounds__(512) nvl_allreduce_nvls_oneshot_kernel(NvlPeers P, __nv_bfloat16* __restrict__ out, int n,
                                                                         uint32_t* epoch_ptr, uint32_t* done_ctr, int channel) {
  const uint32_t epoch = epoch_begin(epoch_ptr);
  nvl_barrier(P, channel, epoch);
  const int nvec = n >> 3;
  const __nv_bfloat16* mc = reinterpret_cast<const __nv_bfloat16*>(P.mc_buf);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += gridDim.x * blockDim.x)
    st128(out + (size_t)i * 8, mc_ld_reduce_bf16x8(mc + (size_t)i * 8));
  epoch_end(epoch_ptr, done_ctr, epoch);
}

// ---------------------------------------------------------------- NVLS two-shot: reduce-scatter by ld_reduce + all-gather by multimem.st
// rank r reduces slice r through the switch and broadcasts it into the result region of EVERY rank's slot with one multicast
// store per 16 bytes; bandwidth-optimal for prefill-size messages (each byte crosses each link once in, once out).
__global__ void __launch_bounds__(512) nvl_allreduce_nvls_twoshot_kernel(NvlPeers P, __nv_bfloat16* __restrict__ out, int n,
                                                                         uint32_t* epoch_ptr, uint32_t* done_ctr, int channel) {
  const uint32_t epoch = epoch_begin(epoch_ptr);
  nvl_barrier(P, channel, 2 * epoch - 1);
  const int nvec = n >> 3;
  const int per = (nvec + P.world - 1) / P.world;
  const int lo = P.rank * per, hi = min(nvec, lo + per);
  __nv_bfloat16* mc = reinterpret_cast<__nv_bfloat16*>(P.mc_buf);
  for (int i = lo + blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += gridDim.x * blockDim.x)
    mc_st_bf16x8(mc + (size_t)n + (size_t)i * 8, mc_ld_reduce_bf16x8(mc + (size_t)i * 8));
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    atomicAdd(done_ctr + 1, 1);
    uint32_t spins = 0;
    unsigned long long t0 = 0;
    while (*((volatile uint32_t*)(done_ctr + 1)) < gridDim.x) {
      if (nvl_wait_expired(spins, t0)) { printf("sglb nvl: NVLS two-shot grid counter timeout rank=%d block=%d\n", P.rank, blockIdx.x); __trap(); }
    }
  }
  nvl_barrier(P, channel, 2 * epoch);
  const __nv_bfloat16* res = reinterpret_cast<const __nv_bfloat16*>(P.buf[P.rank]) + (size_t)n;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += gridDim.x * blockDim.x)
    st128(out + (size_t)i * 8, ld_peer128(res + (size_t)i * 8));
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const uint32_t t = atomicAdd(done_ctr, 1);
    if (t == gridDim.x - 1) {
      *done_ctr = 0;
      *(done_ctr + 1) = 0;
      __threadfence();
      *((volatile uint32_t*)epoch_ptr) = epoch;
    }
  }
}

// ---------------------------------------------------------------- fused all-reduce + residual add + RMSNorm
// partial rows of every rank sit in the arena slots; one CTA per token row.
//   residual[t] += sum_r partial_r[t] ; out[t] = rmsnorm(residual[t]) * w
template <int MAXV>
__global__ void __launch_bounds__(256) nvl_ar_add_rmsnorm_kernel(NvlPeers P, __nv_bfloat16* __restrict__ residual,
                                                                 const __nv_bfloat16* __restrict__ w, __nv_bfloat16* __restrict__ out,
                                                                 int T, int H, float eps, uint32_t* epoch_ptr, uint32_t* done_ctr,
                                                                 int channel, int gemma) {
  const uint32_t epoch = epoch_begin(epoch_ptr);
  nvl_barrier(P, channel, epoch);
  const int nvec = H >> 3;
  __shared__ float red[8];
  for (int row = blockIdx.x; row < T; row += gridDim.x) {
    float v[MAXV][8];
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int vi = threadIdx.x + i * 256;
      if (vi < nvec) {
        float a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        const size_t off = (size_t)row * H + (size_t)vi * 8;
        if (P.mc_buf) {      // NVLS: the switch returns the cross-rank sum in one load
          acc_bf16x8(a, mc_ld_reduce_bf16x8(reinterpret_cast<const __nv_bfloat16*>(P.mc_buf) + off));
        } else {
#pragma unroll
          for (int r = 0; r < NVL_MAX_RANKS; ++r)
            if (r < P.world) acc_bf16x8(a, ld_peer128(reinterpret_cast<const __nv_bfloat16*>(P.buf[r]) + off));
        }
        if (residual) {
          acc_bf16x8(a, ld128(residual + off));
          const U4 s = pack8(a);
          st128(residual + off, s);
          a[0] = bf16_lo(s.x); a[1] = bf16_hi(s.x); a[2] = bf16_lo(s.y); a[3] = bf16_hi(s.y);
          a[4] = bf16_lo(s.z); a[5] = bf16_hi(s.z); a[6] = bf16_lo(s.w); a[7] = bf16_hi(s.w);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) { v[i][k] = a[k]; ss += a[k] * a[k]; }
      }
    }
    ss = warp_sum(ss);
    __syncthreads();
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
        const U4 ww = ldg128(w + (size_t)vi * 8);
        const float g[8] = {bf16_lo(ww.x), bf16_hi(ww.x), bf16_lo(ww.y), bf16_hi(ww.y), bf16_lo(ww.z), bf16_hi(ww.z), bf16_lo(ww.w), bf16_hi(ww.w)};
        float o[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] = v[i][k] * rs * (gemma ? 1.f + g[k] : g[k]);
        st128(out + (size_t)row * H + (size_t)vi * 8, pack8(o));
      }
    }
  }
  epoch_end(epoch_ptr, done_ctr, epoch);
}

// ---------------------------------------------------------------- fused reduce-scatter + residual add + RMSNorm + all-gather (large T)
// The one-shot kernel above makes EVERY rank reduce EVERY row: per-rank NVLink traffic grows with world x bytes, which is what caps TP=4/8 at
// decode batches of a few hundred tokens (8 MB of partials at bs 1024).  Here rank r owns rows r, r + world, ...:
//   phase 1  (after flag round 2e-1)  reduced = multimem.ld_reduce(partials[row])  (or peer pulls without NVLS)
//            residual[row] += reduced          -- the residual stream stays ROW-SHARDED across the whole layer stack: only its owner
//                                                  ever reads or writes a row, and ownership depends on (row, world) only
//            out[row] = rmsnorm(residual[row]) * w   -> ONE multimem.st per 16 bytes lands it in the result region of every rank's slot
//   phase 2  grid-wide completion, flag round 2e: every rank's result region holds all T normalised rows; the next GEMM reads it in place.
// Bytes per rank: T*H*2 in (reduce) + T*H*2 out (gather), independent of world.
template <int MAXV>
__global__ void __launch_bounds__(256) nvl_rs_add_rmsnorm_ag_kernel(NvlPeers P, __nv_bfloat16* __restrict__ residual,
                                                                    const __nv_bfloat16* __restrict__ w, size_t out_off, int T, int H, float eps,
                                                                    uint32_t* epoch_ptr, uint32_t* done_ctr, int channel, int gemma) {
  const uint32_t epoch = epoch_begin(epoch_ptr);
  nvl_barrier(P, channel, 2 * epoch - 1);
  const int nvec = H >> 3;
  __shared__ float red[8];
  for (int row = blockIdx.x * P.world + P.rank; row < T; row += gridDim.x * P.world) {
    float v[MAXV][8];
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int vi = threadIdx.x + i * 256;
      if (vi < nvec) {
        float a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        const size_t off = (size_t)row * H + (size_t)vi * 8;
        if (P.mc_buf) {
          acc_bf16x8(a, mc_ld_reduce_bf16x8(reinterpret_cast<const __nv_bfloat16*>(P.mc_buf) + off));
        } else {
#pragma unroll
          for (int r = 0; r < NVL_MAX_RANKS; ++r)
            if (r < P.world) acc_bf16x8(a, ld_peer128(reinterpret_cast<const __nv_bfloat16*>(P.buf[r]) + off));
        }
        if (residual) {
          acc_bf16x8(a, ld128(residual + off));
          const U4 s = pack8(a);
          st128(residual + off, s);
          a[0] = bf16_lo(s.x); a[1] = bf16_hi(s.x); a[2] = bf16_lo(s.y); a[3] = bf16_hi(s.y);
          a[4] = bf16_lo(s.z); a[5] = bf16_hi(s.z); a[6] = bf16_lo(s.w); a[7] = bf16_hi(s.w);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) { v[i][k] = a[k]; ss += a[k] * a[k]; }
      }
    }
    ss = warp_sum(ss);
    __syncthreads();
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
        const U4 ww = ldg128(w + (size_t)vi * 8);
        const float g[8] = {bf16_lo(ww.x), bf16_hi(ww.x), bf16_lo(ww.y), bf16_hi(ww.y), bf16_lo(ww.z), bf16_hi(ww.z), bf16_lo(ww.w), bf16_hi(ww.w)};
        float o[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] = v[i][k] * rs * (gemma ? 1.f + g[k] : g[k]);
        const U4 ov = pack8(o);
        const size_t eoff = out_off + (size_t)row * H + (size_t)vi * 8;
        if (P.mc_buf) {
          mc_st_bf16x8(reinterpret_cast<__nv_bfloat16*>(P.mc_buf) + eoff, ov);
        } else {
#pragma unroll
          for (int r = 0; r < NVL_MAX_RANKS; ++r)
            if (r < P.world) st128(const_cast<__nv_bfloat16*>(reinterpret_cast<const __nv_bfloat16*>(P.buf[r])) + eoff, ov);
        }
      }
    }
  }
  // every store of this grid must be on its way before block 0 publishes the second flag
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    atomicAdd(done_ctr + 1, 1);
    if (blockIdx.x == 0) {
      uint32_t spins = 0;
      unsigned long long t0 = 0;
      while (*((volatile uint32_t*)(done_ctr + 1)) < gridDim.x) {
        if (nvl_wait_expired(spins, t0)) { printf("sglb nvl: rs+norm+ag grid counter timeout rank=%d\n", P.rank); __trap(); }
      }
      __threadfence_system();
    }
  }
  nvl_barrier(P, channel, 2 * epoch);
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const uint32_t t = atomicAdd(done_ctr, 1);
    if (t == gridDim.x - 1) {
      *done_ctr = 0;
      *(done_ctr + 1) = 0;
      __threadfence();
      *((volatile uint32_t*)epoch_ptr) = epoch;
    }
  }
}

// ---------------------------------------------------------------- consumer of the push-GEMM (gemm_tc.cu EPI_PUSH)
// All `world` partial copies of the [T,H] activation already sit in THIS rank's arena slot (layout [world][T][H]); wait
// for every source rank's flag, then sum locally (+ residual) and normalise.  No NVLink traffic in this kernel.
template <int MAXV>
__global__ void __launch_bounds__(256) nvl_local_reduce_add_rmsnorm_kernel(const __nv_bfloat16* __restrict__ slot, uint32_t* flags,
                                                                           __nv_bfloat16* __restrict__ residual,
                                                                           const __nv_bfloat16* __restrict__ w,
                                                                           __nv_bfloat16* __restrict__ out, int T, int H, int world,
                                                                           int nslabs, float eps, uint32_t* epoch_ptr, uint32_t* done_ctr,
                                                                           int gemma) {
  // flags == nullptr: the slabs were produced by a LOCAL split-K GEMM on this stream (no cross-rank wait, no epoch)
  const uint32_t epoch = epoch_ptr ? epoch_begin(epoch_ptr) : 0;
  if (flags && threadIdx.x < world) {
    uint32_t spins = 0;
    unsigned long long t0 = 0;
    while ((int32_t)(ld_relaxed_sys(flags + threadIdx.x) - epoch) < 0) {
      if (nvl_wait_expired(spins, t0)) {
        printf("sglb nvl: push-flag timeout (45 s) block=%d waiting on rank=%d epoch=%u have=%u\n", blockIdx.x, threadIdx.x, epoch,
               ld_acquire_sys(flags + threadIdx.x));
        __trap();
      }
    }
    fence_acquire_sys();
  }
  __syncthreads();
  const int nvec = H >> 3;
  const size_t stride = (size_t)T * H;
  __shared__ float red[8];
  for (int row = blockIdx.x; row < T; row += gridDim.x) {
    float v[MAXV][8];
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int vi = threadIdx.x + i * 256;
      if (vi < nvec) {
        float a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        const size_t off = (size_t)row * H + (size_t)vi * 8;
#pragma unroll 4
        for (int r = 0; r < nslabs; ++r) acc_bf16x8(a, ld_peer128(slot + r * stride + off));   // written by peers: bypass stale L1 lines
        if (residual) {
          acc_bf16x8(a, ld128(residual + off));
          const U4 s = pack8(a);
          st128(residual + off, s);
          a[0] = bf16_lo(s.x); a[1] = bf16_hi(s.x); a[2] = bf16_lo(s.y); a[3] = bf16_hi(s.y);
          a[4] = bf16_lo(s.z); a[5] = bf16_hi(s.z); a[6] = bf16_lo(s.w); a[7] = bf16_hi(s.w);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) { v[i][k] = a[k]; ss += a[k] * a[k]; }
      }
    }
    ss = warp_sum(ss);
    __syncthreads();
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
        const U4 ww = ldg128(w + (size_t)vi * 8);
        const float g[8] = {bf16_lo(ww.x), bf16_hi(ww.x), bf16_lo(ww.y), bf16_hi(ww.y), bf16_lo(ww.z), bf16_hi(ww.z), bf16_lo(ww.w), bf16_hi(ww.w)};
        float o[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] = v[i][k] * rs * (gemma ? 1.f + g[k] : g[k]);
        st128(out + (size_t)row * H + (size_t)vi * 8, pack8(o));
      }
    }
  }
  if (epoch_ptr) epoch_end(epoch_ptr, done_ctr, epoch);
}

void nvl_local_reduce_add_rmsnorm(const void* slot, uint32_t* flags, void* residual, const void* w, void* out, int T, int H, int world,
                                  int nslabs, float eps, bool gemma, uint32_t* epoch_ptr, uint32_t* done_ctr, int num_sms,
                                  cudaStream_t stream) {
  if (H % 8 || H > 256 * 8 * 8) throw std::runtime_error("sglb nvl reduce+norm: H must be a multiple of 8 and <= 16384");
  const int grid = max(1, min(T, num_sms * 4));
  const int nvec = H / 8;
  auto S = (const __nv_bfloat16*)slot; auto R = (__nv_bfloat16*)residual; auto W = (const __nv_bfloat16*)w; auto O = (__nv_bfloat16*)out;
#define SGLB_LRN(MV) nvl_local_reduce_add_rmsnorm_kernel<MV><<<grid, 256, 0, stream>>>(S, flags, R, W, O, T, H, world, nslabs, eps, epoch_ptr, done_ctr, gemma)
  if (nvec <= 256) SGLB_LRN(1); else if (nvec <= 512) SGLB_LRN(2); else if (nvec <= 1024) SGLB_LRN(4); else SGLB_LRN(8);
#undef SGLB_LRN
  SGLB_CUDA_CHECK(cudaGetLastError());
}

static NvlPeers make_peers(const int64_t* buf_ptrs, const int64_t* flag_ptrs, int64_t mc_ptr, int64_t slot_off, int rank, int world) {
  if (world > NVL_MAX_RANKS) throw std::runtime_error("sglb nvl: world size > 8");
  NvlPeers P;
  for (int r = 0; r < NVL_MAX_RANKS; ++r) {
    P.buf[r] = r < world ? reinterpret_cast<const void*>(buf_ptrs[r] + slot_off) : nullptr;
    P.flags[r] = r < world ? reinterpret_cast<uint32_t*>(flag_ptrs[r]) : nullptr;
  }
  P.mc_buf = mc_ptr ? reinterpret_cast<void*>(mc_ptr + slot_off) : nullptr;
  P.rank = rank; P.world = world;
  return P;
}

void nvl_allreduce(const int64_t* buf_ptrs, const int64_t* flag_ptrs, int64_t mc_ptr, int64_t slot_off, int rank, int world,
                   void* out, int n, uint32_t* epoch_ptr, uint32_t* done_ctr, int channel, int algo, int num_sms,
                   cudaStream_t stream) {
  if (n % 8) throw std::runtime_error("sglb nvl allreduce: element count must be a multiple of 8");
  NvlPeers P = make_peers(buf_ptrs, flag_ptrs, mc_ptr, slot_off, rank, world);
  const int nvec = n / 8;
  if (algo >= 2 && !P.mc_buf) throw std::runtime_error("sglb nvl allreduce: NVLS algorithms need a multicast mapping");
  if (algo == 2) {
    const int grid = max(1, min((nvec + 511) / 512, num_sms));
    nvl_allreduce_nvls_oneshot_kernel<<<grid, 512, 0, stream>>>(P, (__nv_bfloat16*)out, n, epoch_ptr, done_ctr, channel);
  } else if (algo == 3) {
    const int grid = max(1, min((nvec / world + 511) / 512, num_sms / 2));   // co-resident (spin on the local counter)
    nvl_allreduce_nvls_twoshot_kernel<<<grid, 512, 0, stream>>>(P, (__nv_bfloat16*)out, n, epoch_ptr, done_ctr, channel);
  } else if (algo == 0 || algo == 4) {
    const int grid = max(1, min((nvec + 511) / 512, num_sms));
    if (algo == 4) nvl_allreduce_oneshot_kernel<true><<<grid, 512, 0, stream>>>(P, (__nv_bfloat16*)out, n, epoch_ptr, done_ctr, channel);
    else nvl_allreduce_oneshot_kernel<false><<<grid, 512, 0, stream>>>(P, (__nv_bfloat16*)out, n, epoch_ptr, done_ctr, channel);
  } else {
    const int grid = max(1, min((nvec / world + 511) / 512, num_sms / 2));   // all blocks must be co-resident (spin)
    nvl_allreduce_twoshot_kernel<<<grid, 512, 0, stream>>>(P, (__nv_bfloat16*)out, n, epoch_ptr, done_ctr, channel);
  }
  SGLB_CUDA_CHECK(cudaGetLastError());
}

void nvl_ar_add_rmsnorm(const int64_t* buf_ptrs, const int64_t* flag_ptrs, int64_t mc_ptr, int64_t slot_off, int rank, int world,
                        void* residual, const void* w, void* out, int T, int H, float eps, bool gemma, uint32_t* epoch_ptr,
                        uint32_t* done_ctr, int channel, int num_sms, cudaStream_t stream) {
  if (H % 8 || H > 256 * 8 * 8) throw std::runtime_error("sglb nvl ar+norm: H must be a multiple of 8 and <= 16384");
  NvlPeers P = make_peers(buf_ptrs, flag_ptrs, mc_ptr, slot_off, rank, world);
  const int grid = max(1, min(T, num_sms * 4));
  const int nvec = H / 8;
  auto R = (__nv_bfloat16*)residual; auto W = (const __nv_bfloat16*)w; auto O = (__nv_bfloat16*)out;
  if (nvec <= 256) nvl_ar_add_rmsnorm_kernel<1><<<grid, 256, 0, stream>>>(P, R, W, O, T, H, eps, epoch_ptr, done_ctr, channel, gemma);
  else if (nvec <= 512) nvl_ar_add_rmsnorm_kernel<2><<<grid, 256, 0, stream>>>(P, R, W, O, T, H, eps, epoch_ptr, done_ctr, channel, gemma);
  else if (nvec <= 1024) nvl_ar_add_rmsnorm_kernel<4><<<grid, 256, 0, stream>>>(P, R, W, O, T, H, eps, epoch_ptr, done_ctr, channel, gemma);
  else nvl_ar_add_rmsnorm_kernel<8><<<grid, 256, 0, stream>>>(P, R, W, O, T, H, eps, epoch_ptr, done_ctr, channel, gemma);
  SGLB_CUDA_CHECK(cudaGetLastError());
}

void nvl_rs_add_rmsnorm_ag(const int64_t* buf_ptrs, const int64_t* flag_ptrs, int64_t mc_ptr, int64_t slot_off, int64_t out_off_bytes, int rank,
                           int world, void* residual, const void* w, int T, int H, float eps, bool gemma, uint32_t* epoch_ptr, uint32_t* done_ctr,
                           int channel, int num_sms, cudaStream_t stream) {
  if (H % 8 || H > 256 * 8 * 8) throw std::runtime_error("sglb nvl rs+norm+ag: H must be a multiple of 8 and <= 16384");
  if (out_off_bytes % 16) throw std::runtime_error("sglb nvl rs+norm+ag: result region must be 16-byte aligned");
  NvlPeers P = make_peers(buf_ptrs, flag_ptrs, mc_ptr, slot_off, rank, world);
  const int own = (T + world - 1) / world;
  const int grid = max(1, min(own, num_sms * 2));      // every block polls the second flag round: all of them must be resident
  const int nvec = H / 8;
  const size_t out_off = (size_t)out_off_bytes / 2;    // in bf16 elements, relative to the slot base the peers table already points at
  auto R = (__nv_bfloat16*)residual; auto W = (const __nv_bfloat16*)w;
  if (nvec <= 256) nvl_rs_add_rmsnorm_ag_kernel<1><<<grid, 256, 0, stream>>>(P, R, W, out_off, T, H, eps, epoch_ptr, done_ctr, channel, gemma);
  else if (nvec <= 512) nvl_rs_add_rmsnorm_ag_kernel<2><<<grid, 256, 0, stream>>>(P, R, W, out_off, T, H, eps, epoch_ptr, done_ctr, channel, gemma);
  else if (nvec <= 1024) nvl_rs_add_rmsnorm_ag_kernel<4><<<grid, 256, 0, stream>>>(P, R, W, out_off, T, H, eps, epoch_ptr, done_ctr, channel, gemma);
  else nvl_rs_add_rmsnorm_ag_kernel<8><<<grid, 256, 0, stream>>>(P, R, W, out_off, T, H, eps, epoch_ptr, done_ctr, channel, gemma);
  SGLB_CUDA_CHECK(cudaGetLastError());
}

}  // namespace sglb
