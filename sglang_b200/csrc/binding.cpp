// torch.ops.sglb.* — thin adapters from at::Tensor to the raw-pointer kernel API (kernels/api.h).
// Loaded with torch.ops.load_library(sglang_b200/_C.so).  Every adapter bumps a launch counter so bench.py can
// report how many of OUR kernels ran inside the timed region (CUDA-graph replays are accounted in Python).
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/library.h>
#include <torch/types.h>

#include <atomic>

#include "kernels/api.h"

namespace {

std::atomic<int64_t> g_launches{0};
inline void count(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }
inline cudaStream_t cur_stream() { return at::cuda::getCurrentCUDAStream().stream(); }
int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  }
  return n;
}
#define CHECK_CUDA(t) TORCH_CHECK((t).is_cuda(), #t " must be a CUDA tensor")
#define CHECK_BF16(t) TORCH_CHECK((t).scalar_type() == at::kBFloat16, #t " must be bfloat16")
#define CHECK_I32(t) TORCH_CHECK((t).scalar_type() == at::kInt, #t " must be int32")
#define CHECK_LAST_CONTIG(t) TORCH_CHECK((t).stride(-1) == 1, #t " must be contiguous in the last dim")
template <typename T>
const T* opt_ptr(const c10::optional<at::Tensor>& t) { return t.has_value() ? t->data_ptr<T>() : nullptr; }
template <typename T>
T* opt_mptr(const c10::optional<at::Tensor>& t) { return t.has_value() ? t->data_ptr<T>() : nullptr; }

int64_t launch_count() { return g_launches.load(); }
void reset_launch_count() { g_launches.store(0); }

void rmsnorm(at::Tensor x, c10::optional<at::Tensor> residual, const at::Tensor& w, at::Tensor out, double eps, bool gemma) {
  CHECK_CUDA(x); CHECK_BF16(x); CHECK_BF16(w); CHECK_BF16(out); CHECK_LAST_CONTIG(x); CHECK_LAST_CONTIG(out);
  TORCH_CHECK(x.dim() == 2 && out.dim() == 2);
  if (residual.has_value()) { CHECK_BF16(*residual); TORCH_CHECK(residual->is_contiguous()); }
  c10::cuda::CUDAGuard g(x.device());
  sglb::rmsnorm(x.data_ptr(), residual.has_value() ? residual->data_ptr() : nullptr, w.data_ptr(), out.data_ptr(),
                (int)x.size(0), (int)x.size(1), (int)x.stride(0), (int)out.stride(0), (float)eps, gemma, cur_stream());
  count();
}

void act_and_mul(const at::Tensor& x, at::Tensor out, int64_t act) {
  CHECK_CUDA(x); CHECK_BF16(x); CHECK_BF16(out); CHECK_LAST_CONTIG(x); TORCH_CHECK(out.is_contiguous());
  c10::cuda::CUDAGuard g(x.device());
  sglb::act_and_mul(x.data_ptr(), out.data_ptr(), (int)x.size(0), (int)out.size(1), (int)x.stride(0), (int)act, cur_stream());
  count();
}

void rope_kv_write(at::Tensor q, const at::Tensor& k, const at::Tensor& v, const at::Tensor& positions,
                   const at::Tensor& cos_sin, at::Tensor k_cache, at::Tensor v_cache, const at::Tensor& out_loc, bool neox,
                   double k_scale, double v_scale, int64_t nslabs, int64_t slab_stride) {
  CHECK_CUDA(q); CHECK_BF16(q); CHECK_BF16(k); CHECK_BF16(v); CHECK_I32(positions); CHECK_I32(out_loc);
  TORCH_CHECK(q.dim() == 3 && k.dim() == 3 && v.dim() == 3, "q/k/v must be [T, H, D]");
  TORCH_CHECK(cos_sin.scalar_type() == at::kFloat && cos_sin.is_contiguous());
  TORCH_CHECK(q.stride(2) == 1 && q.stride(1) == q.size(2) && k.stride(1) == k.size(2) && v.stride(1) == v.size(2));
  const bool fp8 = k_cache.scalar_type() == at::kFloat8_e4m3fn;
  TORCH_CHECK(fp8 || k_cache.scalar_type() == at::kBFloat16, "kv cache must be bf16 or fp8_e4m3");
  TORCH_CHECK(k_cache.is_contiguous() && v_cache.is_contiguous());
  c10::cuda::CUDAGuard g(q.device());
  sglb::rope_kv_write(q.data_ptr(), k.data_ptr(), v.data_ptr(), positions.data_ptr<int>(), cos_sin.data_ptr<float>(),
                      k_cache.data_ptr(), v_cache.data_ptr(), out_loc.data_ptr<int>(), (int)q.size(0), (int)q.size(1),
                      (int)k.size(1), (int)q.size(2), (int)cos_sin.size(1), (int)q.stride(0), (int)k.stride(0),
                      (int)v.stride(0), neox, fp8, (float)k_scale, (float)v_scale, (int)std::max<int64_t>(1, nslabs), (long long)slab_stride,
                      cur_stream());
  count();
}

void apply_pt_delta(at::Tensor page_table, const at::Tensor& rows, const at::Tensor& cols, const at::Tensor& vals,
                    c10::optional<at::Tensor> n, int64_t n_max) {
  CHECK_CUDA(page_table); CHECK_I32(page_table); CHECK_I32(rows); CHECK_I32(cols); CHECK_I32(vals);
  if (n_max <= 0) return;
  c10::cuda::CUDAGuard g(page_table.device());
  sglb::apply_pt_delta(page_table.data_ptr<int>(), (int)page_table.stride(0), rows.data_ptr<int>(), cols.data_ptr<int>(),
                       vals.data_ptr<int>(), opt_ptr<int>(n), (int)n_max, cur_stream());
  count();
}

void embedding(const at::Tensor& weight, const at::Tensor& ids, c10::optional<at::Tensor> gather_idx, at::Tensor out,
               int64_t vocab_start) {
  CHECK_CUDA(weight); CHECK_BF16(weight); CHECK_I32(ids); CHECK_BF16(out);
  TORCH_CHECK(weight.is_contiguous() && out.is_contiguous());
  c10::cuda::CUDAGuard g(weight.device());
  sglb::embedding(weight.data_ptr(), ids.data_ptr<int>(), opt_ptr<int>(gather_idx), out.data_ptr(), (int)out.size(0),
                  (int)weight.size(1), (int)weight.size(0), (int)vocab_start, cur_stream());
  count();
}

void gather_rows(const at::Tensor& x, const at::Tensor& idx, at::Tensor out) {
  CHECK_CUDA(x); CHECK_BF16(x); CHECK_I32(idx); CHECK_LAST_CONTIG(x); TORCH_CHECK(out.is_contiguous());
  c10::cuda::CUDAGuard g(x.device());
  sglb::gather_rows(x.data_ptr(), idx.data_ptr<int>(), out.data_ptr(), (int)idx.size(0), (int)x.size(1), (int)x.stride(0),
                    cur_stream());
  count();
}

// impl: 0 = tcgen05, 1 = SIMT reference
void attn_fwd(const at::Tensor& q, const at::Tensor& k_cache, const at::Tensor& v_cache, at::Tensor out,
              c10::optional<at::Tensor> lse, c10::optional<at::Tensor> part_o, c10::optional<at::Tensor> part_lse,
              const at::Tensor& page_table, const at::Tensor& req_slots, const at::Tensor& seq_lens,
              c10::optional<at::Tensor> cu_q, c10::optional<at::Tensor> tok2req, int64_t max_q_len, int64_t page_size,
              int64_t num_splits, double scale, bool causal, int64_t window, double logit_cap, double k_scale,
              double v_scale, int64_t impl, c10::optional<at::Tensor> tree_anc) {
  CHECK_CUDA(q); CHECK_BF16(q); CHECK_BF16(out); CHECK_I32(page_table); CHECK_I32(req_slots); CHECK_I32(seq_lens);
  TORCH_CHECK(q.dim() == 3 && q.stride(2) == 1 && q.stride(1) == q.size(2), "q must be [T, Hq, D] with packed heads");
  TORCH_CHECK(out.is_contiguous() && k_cache.is_contiguous() && v_cache.is_contiguous());
  TORCH_CHECK(k_cache.dim() == 3, "k_cache must be [slots, Hkv, D]");
  const int T = (int)q.size(0), Hq = (int)q.size(1), D = (int)q.size(2), Hkv = (int)k_cache.size(1);
  const int B = (int)req_slots.size(0);
  const bool fp8 = k_cache.scalar_type() == at::kFloat8_e4m3fn;
  c10::cuda::CUDAGuard g(q.device());
  if (impl == 0 || impl == 2) {
    TORCH_CHECK(!fp8, "tcgen05 attention: fp8 KV cache not yet supported (use impl=1)");
    if (num_splits > 1) TORCH_CHECK(part_o.has_value() && part_lse.has_value(), "split-KV needs workspaces");
    // persistent kernels take work items from an atomic counter (one zero-initialised [next, done] pair per device, re-armed by the kernel)
    static at::Tensor sched[16];
    const int di = q.get_device();
    TORCH_CHECK(di >= 0 && di < 16);
    if (!sched[di].defined()) {
      cudaStreamCaptureStatus cst = cudaStreamCaptureStatusNone;
      cudaStreamIsCapturing(cur_stream(), &cst);
      TORCH_CHECK(cst == cudaStreamCaptureStatusNone, "attn_fwd: first call on this device is inside a CUDA-graph capture (warm up eagerly first)");
      sched[di] = at::zeros({2}, q.options().dtype(at::kInt));
    }
    int* sched_ptr = sched[di].data_ptr<int>();
    // decode (one query token per request): keys-as-M kernel (attention_decode.cu) unless SGLB_ATTN_DECODE=v2 asks for the query-major one
    static const bool decode_v3 = [] { const char* e = getenv("SGLB_ATTN_DECODE"); return !(e && std::string(e) == "v2"); }();
    if (tree_anc.has_value())
      TORCH_CHECK(cu_q.has_value() && tree_anc->scalar_type() == at::kInt && tree_anc->is_contiguous() && tree_anc->numel() == T,
                  "tree_anc must be int32 [T] (ancestor bits per query token)");
    if (!cu_q.has_value() && decode_v3 && sglb::attn_decode_tc_supported(Hq, Hkv, D, (int)page_size)) {
      TORCH_CHECK(T == B, "decode attention expects one query token per request");
      sglb::attn_decode_tc(q.data_ptr(), k_cache.data_ptr(), v_cache.data_ptr(), out.data_ptr(), opt_mptr<float>(lse), opt_mptr<float>(part_o),
                           opt_mptr<float>(part_lse), page_table.data_ptr<int>(), (int)page_table.stride(0), req_slots.data_ptr<int>(),
                           seq_lens.data_ptr<int>(), B, (int)q.stride(0), Hq, Hkv, D, (int)k_cache.size(0), (int)page_size, (int)num_splits,
                           (float)scale, (int)window, (float)logit_cap, sched_ptr, num_sms(), cur_stream());
      count(num_splits > 1 ? 2 : 1);
      return;
    }
    if (impl != 2) sched_ptr = nullptr;      // impl 0: classic (non-persistent) launch of the query-major kernel
    sglb::attn_fwd_tc(q.data_ptr(), k_cache.data_ptr(), v_cache.data_ptr(), out.data_ptr(), opt_mptr<float>(lse),
                      opt_mptr<float>(part_o), opt_mptr<float>(part_lse), page_table.data_ptr<int>(),
                      (int)page_table.stride(0), req_slots.data_ptr<int>(), seq_lens.data_ptr<int>(), opt_ptr<int>(cu_q), B, T,
                      (int)q.stride(0), (int)max_q_len, Hq, Hkv, D, (int)k_cache.size(0), (int)page_size, (int)num_splits,
                      (float)scale, causal ? 1 : 0, (int)window, (float)logit_cap, sched_ptr, num_sms(), cur_stream(),
                      tree_anc.has_value() ? reinterpret_cast<const unsigned*>(tree_anc->data_ptr<int>()) : nullptr);
    count(num_splits > 1 ? 2 : 1);
  } else {
    TORCH_CHECK(!tree_anc.has_value(), "tree masks need the tcgen05 attention (impl 0 / 2)");
    sglb::attn_fwd_simt(q.data_ptr(), k_cache.data_ptr(), v_cache.data_ptr(), out.data_ptr(), opt_mptr<float>(lse),
                        page_table.data_ptr<int>(), (int)page_table.stride(0), req_slots.data_ptr<int>(),
                        seq_lens.data_ptr<int>(), opt_ptr<int>(cu_q), opt_ptr<int>(tok2req), T, (int)q.stride(0), Hq, Hkv, D,
                        (int)page_size, (float)scale, causal ? 1 : 0, (int)window, (float)logit_cap, fp8, (float)k_scale,
                        (float)v_scale, cur_stream());
    count();
  }
}

// Mamba: causal depthwise conv1d over a ragged batch with the rolling per-request state (decode when cu_q is None)
void causal_conv1d(const at::Tensor& x, const at::Tensor& w, c10::optional<at::Tensor> bias, at::Tensor conv_state, c10::optional<at::Tensor> cu_q,
                   const at::Tensor& req_slots, const at::Tensor& seq_lens, int64_t max_q_len, bool silu, at::Tensor u) {
  CHECK_CUDA(x); CHECK_CUDA(conv_state); CHECK_I32(req_slots); CHECK_I32(seq_lens);
  TORCH_CHECK(x.dim() == 2 && x.stride(1) == 1 && w.dim() == 2 && w.is_contiguous() && conv_state.dim() == 3 && conv_state.is_contiguous());
  const bool bf16 = x.scalar_type() == at::kBFloat16;
  TORCH_CHECK((bf16 || x.scalar_type() == at::kFloat) && w.scalar_type() == x.scalar_type() && conv_state.scalar_type() == x.scalar_type(),
              "causal_conv1d: x / weight / state must share bf16 or fp32");
  const int C = (int)x.size(1), K = (int)w.size(1);
  TORCH_CHECK(w.size(0) == C && conv_state.size(1) == C && conv_state.size(2) == K - 1 && u.scalar_type() == at::kFloat && u.stride(1) == 1 &&
              u.size(0) == x.size(0) && u.size(1) == C);
  if (bias.has_value()) TORCH_CHECK(bias->scalar_type() == x.scalar_type() && bias->numel() == C);
  c10::cuda::CUDAGuard g(x.device());
  sglb::causal_conv1d(x.data_ptr(), (int)x.stride(0), w.data_ptr(), bias.has_value() ? bias->data_ptr() : nullptr, conv_state.data_ptr(),
                      opt_ptr<int>(cu_q), req_slots.data_ptr<int>(), seq_lens.data_ptr<int>(), (int)req_slots.size(0), (int)max_q_len, C, K, silu, bf16,
                      u.data_ptr<float>(), (int)u.stride(0), cur_stream());
  count(2);
}

// Mamba-2 decode: S = S * exp(dt A) + dt x B ; y = <S, C> + D x   (state fp32 [slots, H, P, N], one token per request)
void selective_state_update(at::Tensor state, const at::Tensor& req_slots, const at::Tensor& x, const at::Tensor& dt, const at::Tensor& A,
                            const at::Tensor& Bm, const at::Tensor& Cm, c10::optional<at::Tensor> D, at::Tensor y) {
  CHECK_CUDA(state); CHECK_I32(req_slots);
  TORCH_CHECK(state.dim() == 4 && state.is_contiguous() && state.scalar_type() == at::kFloat);
  const int H = (int)state.size(1), P = (int)state.size(2), N = (int)state.size(3), B = (int)req_slots.size(0);
  for (const at::Tensor* t : std::initializer_list<const at::Tensor*>{&x, &dt, &A, &Bm, &Cm, &y})
    TORCH_CHECK(t->scalar_type() == at::kFloat && t->is_cuda(), "selective_state_update: fp32 CUDA tensors");
  TORCH_CHECK(x.dim() == 2 && x.stride(1) == 1 && x.size(0) == B && x.size(1) == H * P && dt.is_contiguous() && dt.numel() == B * H && A.numel() == H);
  TORCH_CHECK(Bm.dim() == 2 && Bm.stride(1) == 1 && Cm.dim() == 2 && Cm.stride(1) == 1 && Bm.stride(0) == Cm.stride(0) && Bm.size(1) % N == 0);
  TORCH_CHECK((reinterpret_cast<uintptr_t>(Bm.data_ptr()) & 15) == 0 && (reinterpret_cast<uintptr_t>(Cm.data_ptr()) & 15) == 0 && Bm.stride(0) % 4 == 0,
              "selective_state_update: B / C rows must be 16-byte aligned");
  const int G = (int)(Bm.size(1) / N);
  TORCH_CHECK(y.is_contiguous() && y.numel() == (int64_t)B * H * P);
  c10::cuda::CUDAGuard g(state.device());
  sglb::selective_state_update(state.data_ptr<float>(), req_slots.data_ptr<int>(), x.data_ptr<float>(), (int)x.stride(0), dt.data_ptr<float>(),
                               A.data_ptr<float>(), Bm.data_ptr<float>(), Cm.data_ptr<float>(), (int)Bm.stride(0),
                               D.has_value() ? D->data_ptr<float>() : nullptr, B, H, P, N, G, y.data_ptr<float>(), cur_stream());
  count();
}

// gated delta rule decode step (Qwen3-Next): u = conv output rows [q (Hk*Dk) | k (Hk*Dk) | v (Hv*Dv)], state fp32 [slots, Hv, Dk, Dv]
void gated_delta_update(at::Tensor state, const at::Tensor& req_slots, const at::Tensor& u, const at::Tensor& g, const at::Tensor& beta, int64_t Hk,
                        at::Tensor out) {
  CHECK_CUDA(state); CHECK_I32(req_slots);
  TORCH_CHECK(state.dim() == 4 && state.is_contiguous() && state.scalar_type() == at::kFloat && u.scalar_type() == at::kFloat && u.dim() == 2 &&
              u.stride(1) == 1 && g.scalar_type() == at::kFloat && beta.scalar_type() == at::kFloat && g.is_contiguous() && beta.is_contiguous() &&
              out.scalar_type() == at::kFloat && out.is_contiguous());
  const int Hv = (int)state.size(1), Dk = (int)state.size(2), Dv = (int)state.size(3), B = (int)req_slots.size(0);
  TORCH_CHECK(Hk > 0 && Hv % Hk == 0 && u.size(0) == B && u.size(1) == 2 * Hk * Dk + (int64_t)Hv * Dv && g.numel() == (int64_t)B * Hv &&
              out.numel() == (int64_t)B * Hv * Dv);
  c10::cuda::CUDAGuard guard(state.device());
  const bool ok = sglb::gated_delta_update(state.data_ptr<float>(), req_slots.data_ptr<int>(), u.data_ptr<float>(), (int)u.stride(0), (int)(Hk * Dk),
                                           (int)(2 * Hk * Dk), g.data_ptr<float>(), beta.data_ptr<float>(), B, Hv, (int)(Hv / Hk), Dk, Dv,
                                           out.data_ptr<float>(), cur_stream());
  TORCH_CHECK(ok, "gated_delta_update: key dim must be 64 or 128 and value dim <= 256");
  count();
}

// bench tool: per-role cycle accounting of the decode attention kernel into buf [ctas * 30] (int64); an empty tensor turns it off
void attn_decode_prof(const at::Tensor& buf) {
  if (buf.numel() == 0) { sglb::attn_decode_set_prof(nullptr); return; }
  CHECK_CUDA(buf); TORCH_CHECK(buf.scalar_type() == at::kLong && buf.is_contiguous() && buf.numel() >= 148 * 40);
  sglb::attn_decode_set_prof((long long*)buf.data_ptr<int64_t>());
}

// out[M, N(/2 for silu_mul)] = a[M, K] @ w[N, K]^T (+bias | silu*mul)
void gemm(const at::Tensor& a, const at::Tensor& w, at::Tensor out, c10::optional<at::Tensor> bias, int64_t epilogue,
          int64_t bn) {
  CHECK_CUDA(a); CHECK_BF16(a); CHECK_BF16(w); CHECK_BF16(out); CHECK_LAST_CONTIG(a); CHECK_LAST_CONTIG(w); CHECK_LAST_CONTIG(out);
  TORCH_CHECK(a.dim() == 2 && w.dim() == 2 && out.dim() == 2 && a.size(1) == w.size(1));
  TORCH_CHECK((reinterpret_cast<uintptr_t>(a.data_ptr()) & 15) == 0 && (reinterpret_cast<uintptr_t>(w.data_ptr()) & 15) == 0 &&
              (reinterpret_cast<uintptr_t>(out.data_ptr()) & 15) == 0, "gemm operands must be 16-byte aligned");
  c10::cuda::CUDAGuard g(a.device());
  sglb::gemm_bf16_tc(a.data_ptr(), w.data_ptr(), out.data_ptr(), bias.has_value() ? bias->data_ptr() : nullptr,
                     (int)a.size(0), (int)w.size(0), (int)a.size(1), (int)a.stride(0), (int)w.stride(0), (int)out.stride(0),
                     (int)epilogue, (int)bn, num_sms(), cur_stream());
  count();
}
int64_t gemm_pick_bn(int64_t M, int64_t N, int64_t K) { return sglb::gemm_pick_bn((int)M, (int)N, (int)K, num_sms()); }

void sample(const at::Tensor& logits, const at::Tensor& temperature, const at::Tensor& top_k, const at::Tensor& top_p,
            const at::Tensor& min_p, c10::optional<at::Tensor> seeds, c10::optional<at::Tensor> out_slots,
            c10::optional<at::Tensor> next_token_buf, at::Tensor out_tokens, c10::optional<at::Tensor> out_logprobs,
            int64_t global_seed, c10::optional<at::Tensor> step_ptr, int64_t step) {
  CHECK_CUDA(logits); CHECK_LAST_CONTIG(logits); CHECK_I32(top_k); CHECK_I32(out_tokens);
  TORCH_CHECK(logits.dim() == 2);
  const bool bf16 = logits.scalar_type() == at::kBFloat16;
  TORCH_CHECK(bf16 || logits.scalar_type() == at::kFloat, "logits must be bf16 or fp32");
  if (seeds.has_value())
    TORCH_CHECK(seeds->dim() == 2 && seeds->size(0) == logits.size(0) && seeds->size(1) == 2 && seeds->is_contiguous() &&
                seeds->scalar_type() == at::kLong, "seeds must be int64 [B, 2] = (seed, output position | -1)");
  c10::cuda::CUDAGuard g(logits.device());
  sglb::sample(logits.data_ptr(), bf16, (int)logits.size(0), (int)logits.size(1), (int)logits.stride(0),
               temperature.data_ptr<float>(), top_k.data_ptr<int>(), top_p.data_ptr<float>(), min_p.data_ptr<float>(),
               seeds.has_value() ? (const long long*)seeds->data_ptr<int64_t>() : nullptr, opt_ptr<int>(out_slots),
               opt_mptr<int>(next_token_buf), out_tokens.data_ptr<int>(), opt_mptr<float>(out_logprobs),
               (unsigned long long)global_seed, step_ptr.has_value() ? (const long long*)step_ptr->data_ptr<int64_t>() : nullptr,
               (long long)step, cur_stream());
  count();
}

void greedy_shard(const at::Tensor& logits, int64_t v_valid, int64_t v_offset, int64_t rank, int64_t tp, at::Tensor buf) {
  CHECK_CUDA(logits); CHECK_LAST_CONTIG(logits); CHECK_CUDA(buf);
  TORCH_CHECK(logits.dim() == 2 && buf.dim() == 3 && buf.size(0) == tp && buf.size(2) == 4 && buf.is_contiguous() &&
              buf.scalar_type() == at::kBFloat16 && buf.size(1) >= logits.size(0), "greedy_shard: buf must be bf16 [tp, Bpad >= B, 4]");
  const bool bf16 = logits.scalar_type() == at::kBFloat16;
  TORCH_CHECK(bf16 || logits.scalar_type() == at::kFloat, "logits must be bf16 or fp32");
  TORCH_CHECK(v_valid >= 0 && v_valid <= logits.size(1) && v_offset + v_valid < (1 << 24) && rank >= 0 && rank < tp);
  c10::cuda::CUDAGuard g(logits.device());
  sglb::greedy_shard(logits.data_ptr(), bf16, (int)logits.size(0), (int)logits.stride(0), (int)v_valid, (int)v_offset, (int)rank, (int)tp,
                     (int)buf.size(1), buf.data_ptr(), cur_stream());
  count();
}
void greedy_merge(const at::Tensor& buf, int64_t B, c10::optional<at::Tensor> out_slots, c10::optional<at::Tensor> next_token_buf,
                  at::Tensor out_tokens) {
  CHECK_CUDA(buf); CHECK_I32(out_tokens);
  TORCH_CHECK(buf.dim() == 3 && buf.size(2) == 4 && buf.is_contiguous() && buf.scalar_type() == at::kBFloat16 && B <= buf.size(1) &&
              out_tokens.numel() >= B);
  c10::cuda::CUDAGuard g(buf.device());
  sglb::greedy_merge(buf.data_ptr(), (int)buf.size(0), (int)buf.size(1), (int)B, opt_ptr<int>(out_slots), opt_mptr<int>(next_token_buf),
                     out_tokens.data_ptr<int>(), cur_stream());
  count();
}

void apply_token_bitmask(at::Tensor logits, const at::Tensor& bitmask, c10::optional<at::Tensor> row_idx) {
  CHECK_CUDA(logits); CHECK_I32(bitmask); TORCH_CHECK(bitmask.is_contiguous() && logits.dim() == 2);
  const bool bf16 = logits.scalar_type() == at::kBFloat16;
  c10::cuda::CUDAGuard g(logits.device());
  sglb::apply_token_bitmask(logits.data_ptr(), bf16, bitmask.data_ptr<int>(), opt_ptr<int>(row_idx), (int)bitmask.size(0),
                            (int)logits.size(1), (int)logits.stride(0), cur_stream());
  count();
}

void apply_penalties(at::Tensor logits, const at::Tensor& counts, c10::optional<at::Tensor> prompt_mask, const at::Tensor& freq,
                     const at::Tensor& pres, const at::Tensor& rep) {
  CHECK_CUDA(logits); TORCH_CHECK(logits.scalar_type() == at::kFloat && logits.dim() == 2); CHECK_I32(counts);
  c10::cuda::CUDAGuard g(logits.device());
  sglb::apply_penalties(logits.data_ptr<float>(), counts.data_ptr<int>(),
                        prompt_mask.has_value() ? prompt_mask->data_ptr<uint8_t>() : nullptr, freq.data_ptr<float>(),
                        pres.data_ptr<float>(), rep.data_ptr<float>(), (int)logits.size(0), (int)logits.size(1),
                        (int)logits.stride(0), cur_stream());
  count();
}

void bump_counts(at::Tensor counts, const at::Tensor& tokens) {
  CHECK_CUDA(counts); CHECK_I32(counts); CHECK_I32(tokens);
  c10::cuda::CUDAGuard g(counts.device());
  sglb::bump_counts(counts.data_ptr<int>(), tokens.data_ptr<int>(), (int)tokens.size(0), (int)counts.size(1), cur_stream());
  count();
}

void nvl_allreduce(const at::Tensor& buf_ptrs, const at::Tensor& flag_ptrs, int64_t mc_ptr, int64_t slot_off, int64_t rank,
                   int64_t world, at::Tensor out, at::Tensor epoch, at::Tensor done, int64_t channel, int64_t algo) {
  CHECK_CUDA(out); CHECK_BF16(out); TORCH_CHECK(out.is_contiguous() && !buf_ptrs.is_cuda() && !flag_ptrs.is_cuda());
  c10::cuda::CUDAGuard g(out.device());
  sglb::nvl_allreduce(buf_ptrs.data_ptr<int64_t>(), flag_ptrs.data_ptr<int64_t>(), mc_ptr, slot_off, (int)rank, (int)world,
                      out.data_ptr(), (int)out.numel(), reinterpret_cast<uint32_t*>(epoch.data_ptr<int>()),
                      reinterpret_cast<uint32_t*>(done.data_ptr<int>()), (int)channel, (int)algo, num_sms(), cur_stream());
  count();
}

void nvl_ar_add_rmsnorm(const at::Tensor& buf_ptrs, const at::Tensor& flag_ptrs, int64_t mc_ptr, int64_t slot_off, int64_t rank,
                        int64_t world, c10::optional<at::Tensor> residual, const at::Tensor& w, at::Tensor out, double eps, bool gemma,
                        at::Tensor epoch, at::Tensor done, int64_t channel) {
  CHECK_CUDA(out); CHECK_BF16(out); CHECK_BF16(w); TORCH_CHECK(out.is_contiguous() && out.dim() == 2);
  if (residual.has_value()) TORCH_CHECK(residual->is_contiguous());
  c10::cuda::CUDAGuard g(out.device());
  sglb::nvl_ar_add_rmsnorm(buf_ptrs.data_ptr<int64_t>(), flag_ptrs.data_ptr<int64_t>(), mc_ptr, slot_off, (int)rank, (int)world,
                           residual.has_value() ? residual->data_ptr() : nullptr, w.data_ptr(), out.data_ptr(), (int)out.size(0),
                           (int)out.size(1), (float)eps, gemma, reinterpret_cast<uint32_t*>(epoch.data_ptr<int>()),
                           reinterpret_cast<uint32_t*>(done.data_ptr<int>()), (int)channel, num_sms(), cur_stream());
  count();
}

// large-T variant: reduce-scatter + residual add + RMSNorm + all-gather; the normalised rows land in the result region (slot_off + out_off)
// of EVERY rank's arena, the residual stream stays row-sharded (row t belongs to rank t % world)
void nvl_rs_add_rmsnorm_ag(const at::Tensor& buf_ptrs, const at::Tensor& flag_ptrs, int64_t mc_ptr, int64_t slot_off, int64_t out_off, int64_t rank,
                           int64_t world, c10::optional<at::Tensor> residual, const at::Tensor& w, int64_t T, int64_t H, double eps, bool gemma,
                           at::Tensor epoch, at::Tensor done, int64_t channel) {
  CHECK_CUDA(w); CHECK_BF16(w);
  if (residual.has_value()) TORCH_CHECK(residual->is_contiguous() && residual->scalar_type() == at::kBFloat16 && residual->numel() == T * H);
  c10::cuda::CUDAGuard g(w.device());
  sglb::nvl_rs_add_rmsnorm_ag(buf_ptrs.data_ptr<int64_t>(), flag_ptrs.data_ptr<int64_t>(), mc_ptr, slot_off, out_off, (int)rank, (int)world,
                              residual.has_value() ? residual->data_ptr() : nullptr, w.data_ptr(), (int)T, (int)H, (float)eps, gemma,
                              reinterpret_cast<uint32_t*>(epoch.data_ptr<int>()), reinterpret_cast<uint32_t*>(done.data_ptr<int>()), (int)channel,
                              num_sms(), cur_stream());
  count();
}

// fused GEMM + all-reduce producer: a[M,K] @ w[N,K]^T pushed into slot [rank] of every peer's arena
void gemm_push(const at::Tensor& a, const at::Tensor& w, const at::Tensor& peer_slot_ptrs, const at::Tensor& peer_flag_ptrs,
               int64_t rank, int64_t world, at::Tensor epoch, at::Tensor done, int64_t bn) {
  CHECK_CUDA(a); CHECK_BF16(a); CHECK_BF16(w); CHECK_LAST_CONTIG(a); CHECK_LAST_CONTIG(w);
  TORCH_CHECK(a.dim() == 2 && w.dim() == 2 && a.size(1) == w.size(1) && !peer_slot_ptrs.is_cuda() && !peer_flag_ptrs.is_cuda());
  c10::cuda::CUDAGuard g(a.device());
  sglb::gemm_bf16_tc_push(a.data_ptr(), w.data_ptr(), (int)a.size(0), (int)w.size(0), (int)a.size(1), (int)a.stride(0), (int)w.stride(0),
                          peer_slot_ptrs.data_ptr<int64_t>(), peer_flag_ptrs.data_ptr<int64_t>(), (int)rank, (int)world,
                          reinterpret_cast<const uint32_t*>(epoch.data_ptr<int>()), reinterpret_cast<uint32_t*>(done.data_ptr<int>()),
                          (int)bn, num_sms(), cur_stream());
  count();
}

void nvl_local_reduce_add_rmsnorm(const at::Tensor& slot, int64_t flags_ptr, c10::optional<at::Tensor> residual, const at::Tensor& w,
                                  at::Tensor out, int64_t world, int64_t nslabs, double eps, bool gemma, at::Tensor epoch, at::Tensor done) {
  CHECK_CUDA(out); CHECK_BF16(out); CHECK_BF16(w); CHECK_BF16(slot); TORCH_CHECK(out.is_contiguous() && out.dim() == 2);
  c10::cuda::CUDAGuard g(out.device());
  sglb::nvl_local_reduce_add_rmsnorm(slot.data_ptr(), reinterpret_cast<uint32_t*>(flags_ptr),
                                     residual.has_value() ? residual->data_ptr() : nullptr, w.data_ptr(), out.data_ptr(),
                                     (int)out.size(0), (int)out.size(1), (int)world, (int)nslabs, (float)eps, gemma,
                                     reinterpret_cast<uint32_t*>(epoch.data_ptr<int>()), reinterpret_cast<uint32_t*>(done.data_ptr<int>()),
                                     num_sms(), cur_stream());
  count();
}

// HiCache page mover: dev_buf [planes, slots, ...] <-> pinned host_buf; page indices as int32 DEVICE tensors; strides in bytes
void kv_transfer_pages(at::Tensor dev_buf, at::Tensor host_buf, const at::Tensor& dev_pages, const at::Tensor& host_pages, int64_t planes,
                       int64_t dev_plane_stride, int64_t host_plane_stride, int64_t host_page_stride, int64_t chunk_bytes, bool to_host) {
  CHECK_CUDA(dev_buf); CHECK_I32(dev_pages); CHECK_I32(host_pages);
  TORCH_CHECK(!host_buf.is_cuda() && host_buf.is_pinned(), "kv_transfer_pages: the host pool must be pinned memory");
  TORCH_CHECK(dev_pages.is_cuda() && host_pages.is_cuda() && dev_pages.numel() == host_pages.numel());
  c10::cuda::CUDAGuard g(dev_buf.device());
  void* hptr = nullptr;
  TORCH_CHECK(cudaHostGetDevicePointer(&hptr, host_buf.data_ptr(), 0) == cudaSuccess, "kv_transfer_pages: host pool is not device-mapped");
  sglb::kv_transfer_pages(dev_buf.data_ptr(), hptr, dev_pages.data_ptr<int>(), host_pages.data_ptr<int>(), (int)dev_pages.numel(), (int)planes,
                          dev_plane_stride, host_plane_stride, host_page_stride, (int)chunk_bytes, to_host, cur_stream());
  count();
}

// y[perm[i], :] += B[slot] (A[slot] x[perm[i], :])  for the tokens listed in the tile table (see kernels/lora.cu)
void lora_sgmv(const at::Tensor& x, const at::Tensor& a_pool, const at::Tensor& b_pool, const at::Tensor& tile_slot,
               const at::Tensor& tile_start, const at::Tensor& tile_cnt, const at::Tensor& perm, at::Tensor h, at::Tensor y, int64_t b1, int64_t b2) {
  CHECK_CUDA(x); CHECK_BF16(x); CHECK_BF16(y); CHECK_BF16(a_pool); CHECK_BF16(b_pool); CHECK_LAST_CONTIG(x); CHECK_LAST_CONTIG(y);
  CHECK_I32(tile_slot); CHECK_I32(tile_start); CHECK_I32(tile_cnt); CHECK_I32(perm);
  TORCH_CHECK(a_pool.dim() == 4 && b_pool.dim() == 3 && a_pool.is_contiguous() && b_pool.is_contiguous() && x.dim() == 2 && y.dim() == 2);
  TORCH_CHECK(h.scalar_type() == at::kFloat && h.is_contiguous());
  const int n_sub = (int)a_pool.size(1), R = (int)a_pool.size(2), K = (int)a_pool.size(3), N = (int)b_pool.size(1);
  TORCH_CHECK(x.size(1) == K && y.size(1) == N && b_pool.size(2) == R && b_pool.size(0) == a_pool.size(0));
  TORCH_CHECK(h.numel() >= perm.numel() * n_sub * R, "lora_sgmv: scratch too small");
  c10::cuda::CUDAGuard g(x.device());
  sglb::lora_sgmv(x.data_ptr(), (int)x.stride(0), a_pool.data_ptr(), b_pool.data_ptr(), R, n_sub, K, N, (int)b1, (int)b2, tile_slot.data_ptr<int>(),
                  tile_start.data_ptr<int>(), tile_cnt.data_ptr<int>(), perm.data_ptr<int>(), (int)tile_slot.numel(), h.data_ptr<float>(),
                  y.data_ptr(), (int)y.stride(0), cur_stream());
  count(2);
}

// persistent stream-K decode GEMM (M <= 256): out[M, N] = a[M, K] @ w[N, K]^T (+bias) ; epilogue 4: w = [gate; up], out[M, N/2] = silu(g) * u
void gemm_streamk(const at::Tensor& a, const at::Tensor& w, at::Tensor out, c10::optional<at::Tensor> bias, int64_t epilogue) {
  CHECK_CUDA(a); CHECK_BF16(a); CHECK_BF16(w); CHECK_BF16(out); CHECK_LAST_CONTIG(a); CHECK_LAST_CONTIG(w); CHECK_LAST_CONTIG(out);
  TORCH_CHECK(a.dim() == 2 && w.dim() == 2 && out.dim() == 2 && a.size(1) == w.size(1));
  const int M = (int)a.size(0), K = (int)a.size(1);
  const int N = epilogue == 4 ? (int)w.size(0) / 2 : (int)w.size(0);
  TORCH_CHECK(out.size(1) == N && out.size(0) == M, "gemm_streamk: bad output shape");
  TORCH_CHECK((reinterpret_cast<uintptr_t>(a.data_ptr()) & 15) == 0 && (reinterpret_cast<uintptr_t>(out.data_ptr()) & 15) == 0 &&
              (reinterpret_cast<uintptr_t>(w.data_ptr()) & 15) == 0, "gemm_streamk: operands must be 16-byte aligned (TMA)");
  if (epilogue == 1) TORCH_CHECK(bias.has_value() && bias->numel() == N);
  c10::cuda::CUDAGuard g(a.device());
  sglb::gemm_bf16_streamk(a.data_ptr(), w.data_ptr(), out.data_ptr(), bias.has_value() ? bias->data_ptr() : nullptr, M, N, K,
                          (int)a.stride(0), (int)w.stride(0), (int)out.stride(0), (int)epilogue, num_sms(), cur_stream());
  count();
}

// decode-shape GEMM (M <= 256): out is [M, N] (splits == 1) or a slab workspace [splits, M, N] that the consumer sums
int64_t gemm_skinny_splits(int64_t M, int64_t N, int64_t K, int64_t max_splits) {
  return sglb::gemm_skinny_splits((int)M, (int)N, (int)K, (int)max_splits, num_sms());
}
int64_t gemm_skinny_norm_splits(int64_t K, int64_t s) { return sglb::gemm_skinny_norm_splits((int)K, (int)s); }
void gemm_skinny(const at::Tensor& a, const at::Tensor& w, at::Tensor out, c10::optional<at::Tensor> bias, int64_t epilogue,
                 int64_t splits) {
  CHECK_CUDA(a); CHECK_BF16(a); CHECK_BF16(w); CHECK_BF16(out); CHECK_LAST_CONTIG(a); CHECK_LAST_CONTIG(w); CHECK_LAST_CONTIG(out);
  TORCH_CHECK(a.dim() == 2 && w.dim() == 2 && a.size(1) == w.size(1));
  const int M = (int)a.size(0), K = (int)a.size(1);
  const int N = epilogue == 4 ? (int)w.size(0) / 2 : (int)w.size(0);
  TORCH_CHECK(splits >= 1 && (splits == 1 ? out.dim() == 2 : (out.dim() == 3 && out.size(0) >= splits && out.is_contiguous())));
  TORCH_CHECK(out.size(-1) == N && out.size(-2) == M, "gemm_skinny: bad output shape");
  c10::cuda::CUDAGuard g(a.device());
  sglb::gemm_bf16_skinny(a.data_ptr(), w.data_ptr(), out.data_ptr(), bias.has_value() ? bias->data_ptr() : nullptr, M, N, K,
                         (int)a.stride(0), (int)w.stride(0), (int)out.stride(-2), (int)epilogue, (int)splits,
                         splits > 1 ? (long long)out.stride(0) : 0, cur_stream());
  count();
}
void gemm_skinny_push(const at::Tensor& a, const at::Tensor& w, int64_t splits, const at::Tensor& peer_slot_ptrs,
                      const at::Tensor& peer_flag_ptrs, int64_t rank, int64_t world, at::Tensor epoch, at::Tensor done) {
  CHECK_CUDA(a); CHECK_BF16(a); CHECK_BF16(w); CHECK_LAST_CONTIG(a); CHECK_LAST_CONTIG(w);
  TORCH_CHECK(a.dim() == 2 && w.dim() == 2 && a.size(1) == w.size(1) && !peer_slot_ptrs.is_cuda() && !peer_flag_ptrs.is_cuda());
  c10::cuda::CUDAGuard g(a.device());
  sglb::gemm_bf16_skinny_push(a.data_ptr(), w.data_ptr(), (int)a.size(0), (int)w.size(0), (int)a.size(1), (int)a.stride(0),
                              (int)w.stride(0), (int)splits, peer_slot_ptrs.data_ptr<int64_t>(), peer_flag_ptrs.data_ptr<int64_t>(),
                              (int)rank, (int)world, reinterpret_cast<const uint32_t*>(epoch.data_ptr<int>()),
                              reinterpret_cast<uint32_t*>(done.data_ptr<int>()), cur_stream());
  count();
}
// out = rmsnorm(residual += sum_s slabs[s]) — the consumer of a LOCAL split-K GEMM (no flags, stream order is enough)
void slab_reduce_add_rmsnorm(const at::Tensor& slabs, c10::optional<at::Tensor> residual, const at::Tensor& w, at::Tensor out,
                             double eps, bool gemma) {
  CHECK_CUDA(out); CHECK_BF16(out); CHECK_BF16(w); CHECK_BF16(slabs);
  TORCH_CHECK(slabs.dim() == 3 && slabs.is_contiguous() && out.is_contiguous() && out.dim() == 2);
  TORCH_CHECK(slabs.size(1) == out.size(0) && slabs.size(2) == out.size(1) && slabs.size(0) <= 64, "slab_reduce: bad shapes");
  c10::cuda::CUDAGuard g(out.device());
  sglb::nvl_local_reduce_add_rmsnorm(slabs.data_ptr(), nullptr, residual.has_value() ? residual->data_ptr() : nullptr, w.data_ptr(),
                                     out.data_ptr(), (int)out.size(0), (int)out.size(1), 0, (int)slabs.size(0), (float)eps, gemma, nullptr,
                                     nullptr, num_sms(), cur_stream());
  count();
}

void per_token_group_quant_fp8(const at::Tensor& x, at::Tensor q, at::Tensor scales, int64_t group, bool ue8m0) {
  CHECK_CUDA(x); CHECK_BF16(x); CHECK_LAST_CONTIG(x);
  TORCH_CHECK(group == 128, "fp8 activation quant: group size must be 128");
  TORCH_CHECK(x.dim() == 2 && q.is_contiguous() && q.element_size() == 1 && scales.scalar_type() == at::kFloat && scales.is_contiguous());
  c10::cuda::CUDAGuard g(x.device());
  sglb::per_token_group_quant_fp8(x.data_ptr(), q.data_ptr(), scales.data_ptr<float>(), (int)x.size(0), (int)x.size(1), (int)x.stride(0),
                                  ue8m0, cur_stream());
  count();
}

// out[M,N] bf16 = block-scaled (xq[M,K] e4m3, xs[M,K/128]) @ (wq[N,K] e4m3, ws[N/128,K/128])^T
void gemm_fp8_block(const at::Tensor& xq, const at::Tensor& xs, const at::Tensor& wq, const at::Tensor& ws, at::Tensor out) {
  CHECK_CUDA(xq); CHECK_BF16(out); CHECK_LAST_CONTIG(xq); CHECK_LAST_CONTIG(wq); CHECK_LAST_CONTIG(out);
  TORCH_CHECK(xq.element_size() == 1 && wq.element_size() == 1 && xq.dim() == 2 && wq.dim() == 2 && xq.size(1) == wq.size(1));
  TORCH_CHECK(xs.scalar_type() == at::kFloat && ws.scalar_type() == at::kFloat && xs.stride(-1) == 1 && ws.stride(-1) == 1);
  const int64_t kg = (xq.size(1) + 127) / 128;
  TORCH_CHECK(xs.size(0) == xq.size(0) && xs.size(1) == kg && ws.size(0) == (wq.size(0) + 127) / 128 && ws.size(1) == kg,
              "fp8 block gemm: scale shapes must be [M, K/128] and [N/128, K/128]");
  c10::cuda::CUDAGuard g(xq.device());
  sglb::gemm_fp8_block(xq.data_ptr(), xs.data_ptr<float>(), wq.data_ptr(), ws.data_ptr<float>(), out.data_ptr(), (int)xq.size(0),
                       (int)wq.size(0), (int)xq.size(1), (int)xq.stride(0), (int)wq.stride(0), (int)out.stride(0), (int)xs.stride(0),
                       (int)ws.stride(0), num_sms(), cur_stream());
  count();
}

void moe_topk_softmax(const at::Tensor& logits, at::Tensor out_w, at::Tensor out_ids, bool renorm) {
  CHECK_CUDA(logits); CHECK_LAST_CONTIG(logits); CHECK_I32(out_ids);
  TORCH_CHECK(logits.dim() == 2 && out_w.scalar_type() == at::kFloat && out_w.is_contiguous() && out_ids.is_contiguous());
  const bool bf16 = logits.scalar_type() == at::kBFloat16;
  TORCH_CHECK(bf16 || logits.scalar_type() == at::kFloat, "router logits must be bf16 or fp32");
  c10::cuda::CUDAGuard g(logits.device());
  sglb::moe_topk_softmax(logits.data_ptr(), bf16, out_w.data_ptr<float>(), out_ids.data_ptr<int>(), (int)logits.size(0),
                         (int)logits.size(1), (int)out_ids.size(1), (int)logits.stride(0), renorm, cur_stream());
  count();
}

// Whole expert MLP block on expert-sorted rows: align -> gather -> grouped GEMM (silu(gate)*up fused) -> grouped GEMM (routing
// weight fused) -> combine.  With peer pointers the combine pushes its result into every peer's NVLink arena instead of `out`.
void moe_grouped_mlp_impl(const at::Tensor& x, const at::Tensor& w13, const at::Tensor& w2, const at::Tensor& topk_w,
                          const at::Tensor& topk_ids, c10::optional<at::Tensor> out, int64_t expert_offset,
                          const int64_t* peer_slots, const int64_t* peer_flags, int rank, int world, const uint32_t* epoch,
                          uint32_t* done) {
  CHECK_CUDA(x); CHECK_BF16(x); CHECK_BF16(w13); CHECK_BF16(w2); CHECK_I32(topk_ids); CHECK_LAST_CONTIG(x);
  TORCH_CHECK(x.dim() == 2 && w13.dim() == 3 && w2.dim() == 3 && w13.is_contiguous() && w2.is_contiguous());
  TORCH_CHECK(topk_w.scalar_type() == at::kFloat && topk_w.is_contiguous() && topk_ids.is_contiguous());
  const int T = (int)x.size(0), H = (int)x.size(1), K = (int)topk_ids.size(1), E = (int)w13.size(0), I = (int)w2.size(2);
  TORCH_CHECK(w13.size(1) == 2 * I && w13.size(2) == H && w2.size(1) == H && w2.size(0) == E);
  if (T == 0) return;
  c10::cuda::CUDAGuard g(x.device());
  auto stream = cur_stream();
  const int pairs = T * K;
  const int max_tiles = (pairs + 127) / 128 + E;
  const int cap = max_tiles * 128;
  auto i32 = x.options().dtype(at::kInt);
  auto meta = at::empty({cap + pairs + max_tiles + 1}, i32);
  int* sorted_row = meta.data_ptr<int>();
  int* inv_pos = sorted_row + cap;
  int* tile_expert = inv_pos + pairs;
  int* num_tiles = tile_expert + max_tiles;
  auto row_w = at::empty({cap}, x.options().dtype(at::kFloat));
  sglb::moe_align(topk_ids.data_ptr<int>(), topk_w.data_ptr<float>(), pairs, E, (int)expert_offset, sorted_row, inv_pos,
                  row_w.data_ptr<float>(), tile_expert, num_tiles, max_tiles, stream);
  auto xs = at::empty({cap, H}, x.options());
  sglb::moe_gather(x.data_ptr(), sorted_row, num_tiles, xs.data_ptr(), H, (int)x.stride(0), K, cap, num_sms(), stream);
  auto h = at::empty({cap, I}, x.options());
  sglb::gemm_bf16_tc_grouped(xs.data_ptr(), w13.data_ptr(), h.data_ptr(), nullptr, cap, I, H, H, I, E, tile_expert, num_tiles, true,
                             num_sms(), stream);
  auto& ys = xs;   // the gathered inputs are dead once the first GEMM has run (same stream): reuse as the [cap, H] output
  sglb::gemm_bf16_tc_grouped(h.data_ptr(), w2.data_ptr(), ys.data_ptr(), row_w.data_ptr<float>(), cap, H, I, I, H, E, tile_expert,
                             num_tiles, false, num_sms(), stream);
  if (peer_slots) {
    sglb::moe_combine_push(ys.data_ptr(), inv_pos, nullptr, T, H, K, peer_slots, peer_flags, rank, world, epoch, done, num_sms(), stream);
  } else {
    TORCH_CHECK(out.has_value() && out->is_contiguous());
    sglb::moe_combine(ys.data_ptr(), inv_pos, nullptr, out->data_ptr(), T, H, K, num_sms(), stream);
  }
  count(5);
}

void moe_grouped_mlp(const at::Tensor& x, const at::Tensor& w13, const at::Tensor& w2, const at::Tensor& topk_w,
                     const at::Tensor& topk_ids, at::Tensor out, int64_t expert_offset) {
  CHECK_BF16(out);
  moe_grouped_mlp_impl(x, w13, w2, topk_w, topk_ids, out, expert_offset, nullptr, nullptr, 0, 1, nullptr, nullptr);
}

void moe_grouped_mlp_push(const at::Tensor& x, const at::Tensor& w13, const at::Tensor& w2, const at::Tensor& topk_w,
                          const at::Tensor& topk_ids, int64_t expert_offset, const at::Tensor& peer_slot_ptrs,
                          const at::Tensor& peer_flag_ptrs, int64_t rank, int64_t world, at::Tensor epoch, at::Tensor done) {
  TORCH_CHECK(!peer_slot_ptrs.is_cuda() && !peer_flag_ptrs.is_cuda());
  moe_grouped_mlp_impl(x, w13, w2, topk_w, topk_ids, c10::nullopt, expert_offset, peer_slot_ptrs.data_ptr<int64_t>(),
                       peer_flag_ptrs.data_ptr<int64_t>(), (int)rank, (int)world, reinterpret_cast<const uint32_t*>(epoch.data_ptr<int>()),
                       reinterpret_cast<uint32_t*>(done.data_ptr<int>()));
}

}  // namespace

TORCH_LIBRARY(sglb, m) {
  m.def("gemm_streamk(Tensor a, Tensor w, Tensor(a!) out, Tensor? bias, int epilogue) -> ()", &gemm_streamk);
  m.def("kv_transfer_pages(Tensor(a!) dev_buf, Tensor(b!) host_buf, Tensor dev_pages, Tensor host_pages, int planes, int dev_plane_stride, "
        "int host_plane_stride, int host_page_stride, int chunk_bytes, bool to_host) -> ()", &kv_transfer_pages);
  m.def("lora_sgmv(Tensor x, Tensor a_pool, Tensor b_pool, Tensor tile_slot, Tensor tile_start, Tensor tile_cnt, Tensor perm, Tensor(a!) h, "
        "Tensor(b!) y, int b1, int b2) -> ()", &lora_sgmv);
  m.def("gemm_skinny_splits(int M, int N, int K, int max_splits) -> int", &gemm_skinny_splits);
  m.def("gemm_skinny_norm_splits(int K, int s) -> int", &gemm_skinny_norm_splits);
  m.def("gemm_skinny(Tensor a, Tensor w, Tensor(a!) out, Tensor? bias, int epilogue, int splits) -> ()", &gemm_skinny);
  m.def("gemm_skinny_push(Tensor a, Tensor w, int splits, Tensor peer_slot_ptrs, Tensor peer_flag_ptrs, int rank, int world, "
        "Tensor(a!) epoch, Tensor(b!) done) -> ()", &gemm_skinny_push);
  m.def("slab_reduce_add_rmsnorm(Tensor slabs, Tensor(a!)? residual, Tensor w, Tensor(b!) out, float eps, bool gemma) -> ()",
        &slab_reduce_add_rmsnorm);
  m.def("per_token_group_quant_fp8(Tensor x, Tensor(a!) q, Tensor(b!) scales, int group, bool ue8m0) -> ()", &per_token_group_quant_fp8);
  m.def("gemm_fp8_block(Tensor xq, Tensor xs, Tensor wq, Tensor ws, Tensor(a!) out) -> ()", &gemm_fp8_block);
  m.def("moe_topk_softmax(Tensor logits, Tensor(a!) out_w, Tensor(b!) out_ids, bool renorm) -> ()", &moe_topk_softmax);
  m.def("moe_grouped_mlp(Tensor x, Tensor w13, Tensor w2, Tensor topk_w, Tensor topk_ids, Tensor(a!) out, int expert_offset) -> ()",
        &moe_grouped_mlp);
  m.def("moe_grouped_mlp_push(Tensor x, Tensor w13, Tensor w2, Tensor topk_w, Tensor topk_ids, int expert_offset, Tensor peer_slot_ptrs, "
        "Tensor peer_flag_ptrs, int rank, int world, Tensor(a!) epoch, Tensor(b!) done) -> ()", &moe_grouped_mlp_push);
  m.def("gemm_push(Tensor a, Tensor w, Tensor peer_slot_ptrs, Tensor peer_flag_ptrs, int rank, int world, Tensor(a!) epoch, "
        "Tensor(b!) done, int bn) -> ()", &gemm_push);
  m.def("nvl_local_reduce_add_rmsnorm(Tensor slot, int flags_ptr, Tensor(a!)? residual, Tensor w, Tensor(b!) out, int world, int nslabs, float eps, "
        "bool gemma, Tensor(c!) epoch, Tensor(d!) done) -> ()", &nvl_local_reduce_add_rmsnorm);
  m.def("nvl_allreduce(Tensor buf_ptrs, Tensor flag_ptrs, int mc_ptr, int slot_off, int rank, int world, Tensor(a!) out, "
        "Tensor(b!) epoch, Tensor(c!) done, int channel, int algo) -> ()", &nvl_allreduce);
  m.def("nvl_ar_add_rmsnorm(Tensor buf_ptrs, Tensor flag_ptrs, int mc_ptr, int slot_off, int rank, int world, Tensor(a!)? residual, "
        "Tensor w, Tensor(b!) out, float eps, bool gemma, Tensor(c!) epoch, Tensor(d!) done, int channel) -> ()", &nvl_ar_add_rmsnorm);
  m.def("nvl_rs_add_rmsnorm_ag(Tensor buf_ptrs, Tensor flag_ptrs, int mc_ptr, int slot_off, int out_off, int rank, int world, "
        "Tensor(a!)? residual, Tensor w, int T, int H, float eps, bool gemma, Tensor(c!) epoch, Tensor(d!) done, int channel) -> ()",
        &nvl_rs_add_rmsnorm_ag);
  m.def("launch_count() -> int", &launch_count);
  m.def("reset_launch_count() -> ()", &reset_launch_count);
  m.def("gemm_pick_bn(int M, int N, int K) -> int", &gemm_pick_bn);
  m.def("rmsnorm(Tensor(a!) x, Tensor(b!)? residual, Tensor w, Tensor(c!) out, float eps, bool gemma) -> ()", &rmsnorm);
  m.def("act_and_mul(Tensor x, Tensor(a!) out, int act) -> ()", &act_and_mul);
  m.def("rope_kv_write(Tensor(a!) q, Tensor k, Tensor v, Tensor positions, Tensor cos_sin, Tensor(b!) k_cache, "
        "Tensor(c!) v_cache, Tensor out_loc, bool neox, float k_scale, float v_scale, int nslabs=1, int slab_stride=0) -> ()", &rope_kv_write);
  m.def("apply_pt_delta(Tensor(a!) page_table, Tensor rows, Tensor cols, Tensor vals, Tensor? n, int n_max) -> ()", &apply_pt_delta);
  m.def("embedding(Tensor weight, Tensor ids, Tensor? gather_idx, Tensor(a!) out, int vocab_start) -> ()", &embedding);
  m.def("gather_rows(Tensor x, Tensor idx, Tensor(a!) out) -> ()", &gather_rows);
  m.def("attn_fwd(Tensor q, Tensor k_cache, Tensor v_cache, Tensor(a!) out, Tensor(b!)? lse, Tensor(c!)? part_o, "
        "Tensor(d!)? part_lse, Tensor page_table, Tensor req_slots, Tensor seq_lens, Tensor? cu_q, Tensor? tok2req, "
        "int max_q_len, int page_size, int num_splits, float scale, bool causal, int window, float logit_cap, "
        "float k_scale, float v_scale, int impl, Tensor? tree_anc=None) -> ()", &attn_fwd);
  m.def("attn_decode_prof(Tensor buf) -> ()", &attn_decode_prof);
  m.def("green_ctx_streams(int device, int[] sm_counts) -> int[]", [](int64_t device, std::vector<int64_t> counts) {
    std::vector<long long> c(counts.begin(), counts.end());
    auto r = sglb::green_ctx_streams((int)device, c);
    return std::vector<int64_t>(r.begin(), r.end());
  });
  m.def("gated_delta_update(Tensor(a!) state, Tensor req_slots, Tensor u, Tensor g, Tensor beta, int Hk, Tensor(b!) out) -> ()", &gated_delta_update);
  m.def("causal_conv1d(Tensor x, Tensor w, Tensor? bias, Tensor(a!) conv_state, Tensor? cu_q, Tensor req_slots, Tensor seq_lens, int max_q_len, "
        "bool silu, Tensor(b!) u) -> ()", &causal_conv1d);
  m.def("selective_state_update(Tensor(a!) state, Tensor req_slots, Tensor x, Tensor dt, Tensor A, Tensor Bm, Tensor Cm, Tensor? D, Tensor(b!) y) -> ()",
        &selective_state_update);
  m.def("gemm(Tensor a, Tensor w, Tensor(a!) out, Tensor? bias, int epilogue, int bn) -> ()", &gemm);
  m.def("sample(Tensor logits, Tensor temperature, Tensor top_k, Tensor top_p, Tensor min_p, Tensor? seeds, "
        "Tensor? out_slots, Tensor(a!)? next_token_buf, Tensor(b!) out_tokens, Tensor(c!)? out_logprobs, int global_seed, "
        "Tensor? step_ptr, int step) -> ()", &sample);
  m.def("greedy_shard(Tensor logits, int v_valid, int v_offset, int rank, int tp, Tensor(a!) buf) -> ()", &greedy_shard);
  m.def("greedy_merge(Tensor buf, int B, Tensor? out_slots, Tensor(a!)? next_token_buf, Tensor(b!) out_tokens) -> ()", &greedy_merge);
  m.def("apply_token_bitmask(Tensor(a!) logits, Tensor bitmask, Tensor? row_idx) -> ()", &apply_token_bitmask);
  m.def("apply_penalties(Tensor(a!) logits, Tensor counts, Tensor? prompt_mask, Tensor freq, Tensor pres, Tensor rep) -> ()",
        &apply_penalties);
  m.def("bump_counts(Tensor(a!) counts, Tensor tokens) -> ()", &bump_counts);
}
