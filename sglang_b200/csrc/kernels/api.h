// Host launch API of the sm_100a kernels (raw pointers + stream; no torch dependency so the .cu files
// compile in seconds).  binding.cpp adapts these to torch.ops.sglb.*
#pragma once
#include <vector>
#include <cuda_runtime.h>
#include <stdint.h>

namespace sglb {

// elementwise.cu
void rmsnorm(void* x, void* residual, const void* w, void* out, int T, int H, int x_stride, int out_stride, float eps,
             bool gemma, cudaStream_t stream);
void act_and_mul(const void* x, void* out, int T, int I, int x_stride, int act, cudaStream_t stream);
void rope_kv_write(void* q, const void* k, const void* v, const int* positions, const float* cos_sin, void* k_cache,
                   void* v_cache, const int* out_loc, int T, int Hq, int Hkv, int D, int rot, int q_stride, int k_stride,
                   int v_stride, bool neox, bool fp8_cache, float k_scale, float v_scale, int nslabs, long long slab_stride,
                   cudaStream_t stream);
void apply_pt_delta(int* page_table, int stride, const int* rows, const int* cols, const int* vals, const int* n_ptr,
                    int n_max, cudaStream_t stream);
void embedding(const void* weight, const int* ids, const int* gather_idx, void* out, int T, int H, int vocab, int vocab_start,
               cudaStream_t stream);
void gather_rows(const void* x, const int* idx, void* out, int n, int H, int x_stride, cudaStream_t stream);
void attn_fwd_simt(const void* q, const void* k_cache, const void* v_cache, void* out, float* lse, const int* page_table,
                   int pt_stride, const int* req_slots, const int* seq_lens, const int* cu_q, const int* tok2req, int T,
                   int q_stride, int Hq, int Hkv, int D, int page_size, float scale, int causal, int window, float logit_cap,
                   bool fp8_cache, float k_scale, float v_scale, cudaStream_t stream);

// attention_tc.cu
// attention_decode.cu: decode attention with the keys as the UMMA M dimension (persistent, 3-deep 32 KB K/V ring, page-granular fetch)
void attn_decode_set_prof(long long* buf);
std::vector<long long> green_ctx_streams(int device, const std::vector<long long>& sm_counts);
bool attn_decode_tc_supported(int Hq, int Hkv, int D, int page_size);
void attn_decode_tc(const void* q, const void* k_cache, const void* v_cache, void* out, float* lse, float* part_o, float* part_lse,
                    const int* page_table, int pt_stride, const int* req_slots, const int* seq_lens, int B, int q_stride, int Hq, int Hkv,
                    int D, int num_slots, int page_size, int num_splits, float scale, int window, float logit_cap, int* sched, int num_sms,
                    cudaStream_t stream);
void attn_fwd_tc(const void* q, const void* k_cache, const void* v_cache, void* out, float* lse, float* part_o,
                 float* part_lse, const int* page_table, int pt_stride, const int* req_slots, const int* seq_lens,
                 const int* cu_q, int B, int T, int q_stride, int max_q_len, int Hq, int Hkv, int D, int num_slots,
                 int page_size, int num_splits, float scale, int causal, int window, float logit_cap, int* sched, int num_sms,
                 cudaStream_t stream, const unsigned* tree_anc = nullptr);

// gemm_tc.cu
void gemm_bf16_tc(const void* A, const void* W, void* C, const void* bias, int M, int N, int K, int lda, int ldw, int ldc,
                  int epilogue, int bn, int num_sms, cudaStream_t stream);
int gemm_pick_bn(int M, int N, int K, int num_sms);
void gemm_bf16_tc_grouped(const void* A, const void* W, void* C, const float* row_scale, int M_cap, int n_out, int K, int lda,
                          int ldc, int num_experts, const int* tile_expert, const int* num_m_tiles, bool split_silu, int num_sms,
                          cudaStream_t stream);

// kvio.cu: device pool <-> pinned host pool page movers (HiCache L1 <-> L2), zero-copy over UVA, layer_first / page_first host layouts
void kv_transfer_pages(void* dev, void* host_devptr, const int* dev_pages, const int* host_pages, int n, int planes, long long dev_plane_stride,
                       long long host_plane_stride, long long host_page_stride, int chunk_bytes, bool to_host, cudaStream_t stream);

// lora.cu: segment-grouped multi-adapter LoRA (shrink + expand) over a slot-paged adapter pool; static metadata -> graph-capturable
void lora_sgmv(const void* x, int ldx, const void* A_pool, const void* B_pool, int rank_cap, int n_sub, int K, int N, int b1, int b2,
               const int* tile_slot, const int* tile_start, const int* tile_cnt, const int* perm, int tile_cap, float* h, void* y, int ldy,
               cudaStream_t stream);

// gemm_streamk.cu (decode shapes: persistent stream-K swap-AB weight streaming, in-kernel fixup, TMA-store epilogue, fused split-SiLU)
void gemm_bf16_streamk(const void* A, const void* W, void* C, const void* bias, int M, int N, int K, int lda, int ldw, int ldc, int epilogue,
                       int num_sms, cudaStream_t stream);

// gemm_skinny.cu (decode shapes: swap-AB, split-K slabs, transposed epilogue, optional NVLink push)
int gemm_skinny_splits(int M, int N, int K, int max_splits, int num_sms);
int gemm_skinny_norm_splits(int K, int s);
void gemm_bf16_skinny(const void* A, const void* W, void* C, const void* bias, int M, int N, int K, int lda, int ldw, int ldc,
                      int epilogue, int splits, long long slab_stride, cudaStream_t stream);
void gemm_bf16_skinny_push(const void* A, const void* W, int M, int N, int K, int lda, int ldw, int splits, const int64_t* peer_slot_ptrs,
                           const int64_t* peer_flag_ptrs, int rank, int world, const uint32_t* epoch_ptr, uint32_t* done_ctr,
                           cudaStream_t stream);

// gemm_fp8.cu
void gemm_fp8_block(const void* A, const float* xs, const void* W, const float* ws, void* C, int M, int N, int K, int lda, int ldw,
                    int ldc, int xs_stride, int ws_stride, int num_sms, cudaStream_t stream);
void per_token_group_quant_fp8(const void* x, void* q, float* scales, int M, int K, int x_stride, bool ue8m0, cudaStream_t stream);

// moe.cu
void moe_topk_softmax(const void* logits, bool bf16, float* out_w, int* out_ids, int T, int E, int K, int stride, bool renorm,
                      cudaStream_t stream);
void moe_align(const int* topk_ids, const float* topk_w, int n_pairs, int E, int expert_offset, int* sorted_row, int* inv_pos,
               float* row_w, int* tile_expert, int* num_tiles, int max_tiles, cudaStream_t stream);
void moe_gather(const void* x, const int* sorted_row, const int* num_tiles, void* xs, int H, int x_stride, int K, int cap_rows,
                int num_sms, cudaStream_t stream);
void moe_combine(const void* ys, const int* inv_pos, const float* pair_w, void* out, int T, int H, int K, int num_sms,
                 cudaStream_t stream);
void moe_combine_push(const void* ys, const int* inv_pos, const float* pair_w, int T, int H, int K, const int64_t* peer_slot_ptrs,
                      const int64_t* peer_flag_ptrs, int rank, int world, const uint32_t* epoch_ptr, uint32_t* done_ctr, int num_sms,
                      cudaStream_t stream);

// sampling.cu
void sample(const void* logits, bool bf16, int B, int V, int stride, const float* temperature, const int* top_k,
            const float* top_p, const float* min_p, const long long* seeds, const int* out_slots, int* next_token_buf,
            int* out_tokens, float* out_logprobs, unsigned long long global_seed, const long long* step_ptr, long long step,
            cudaStream_t stream);
// mamba.cu
void causal_conv1d(const void* x, int ldx, const void* w, const void* bias, void* conv_state, const int* cu_q, const int* req_slots,
                   const int* seq_lens, int B, int max_q_len, int C, int K, bool silu, bool bf16, float* u, int ldu, cudaStream_t stream);
void selective_state_update(float* state, const int* req_slots, const float* x, int ldx, const float* dt, const float* A, const float* Bm,
                            const float* Cm, int ldbc, const float* Dv, int B, int H, int P, int N, int G, float* y, cudaStream_t stream);
bool gated_delta_update(float* state, const int* req_slots, const float* u, int ldu, int k_off, int v_off, const float* g, const float* beta, int B,
                        int Hv, int r, int Dk, int Dv, float* out, cudaStream_t stream);
void greedy_shard(const void* logits, bool bf16, int B, int stride, int v_valid, int v_offset, int rank, int tp, int b_pad, void* buf,
                  cudaStream_t stream);
void greedy_merge(const void* buf, int tp, int b_pad, int B, const int* out_slots, int* next_token_buf, int* out_tokens, cudaStream_t stream);
void apply_token_bitmask(void* logits, bool bf16, const int* bitmask, const int* row_idx, int n_rows, int V, int stride,
                         cudaStream_t stream);
void apply_penalties(float* logits, const int* counts, const unsigned char* prompt_mask, const float* freq, const float* pres,
                     const float* rep, int B, int V, int stride, cudaStream_t stream);
void bump_counts(int* counts, const int* tokens, int B, int V, cudaStream_t stream);

// comm/nvl_allreduce.cu
void nvl_allreduce(const int64_t* buf_ptrs, const int64_t* flag_ptrs, int64_t mc_ptr, int64_t slot_off, int rank, int world,
                   void* out, int n, uint32_t* epoch_ptr, uint32_t* done_ctr, int channel, int algo, int num_sms,
                   cudaStream_t stream);
void nvl_rs_add_rmsnorm_ag(const int64_t* buf_ptrs, const int64_t* flag_ptrs, int64_t mc_ptr, int64_t slot_off, int64_t out_off_bytes, int rank,
                           int world, void* residual, const void* w, int T, int H, float eps, bool gemma, uint32_t* epoch_ptr, uint32_t* done_ctr,
                           int channel, int num_sms, cudaStream_t stream);
void nvl_ar_add_rmsnorm(const int64_t* buf_ptrs, const int64_t* flag_ptrs, int64_t mc_ptr, int64_t slot_off, int rank, int world,
                        void* residual, const void* w, void* out, int T, int H, float eps, bool gemma, uint32_t* epoch_ptr,
                        uint32_t* done_ctr, int channel, int num_sms, cudaStream_t stream);

void gemm_bf16_tc_push(const void* A, const void* W, int M, int N, int K, int lda, int ldw, const int64_t* peer_slot_ptrs,
                       const int64_t* peer_flag_ptrs, int rank, int world, const uint32_t* epoch_ptr, uint32_t* done_ctr, int bn,
                       int num_sms, cudaStream_t stream);
void nvl_local_reduce_add_rmsnorm(const void* slot, uint32_t* flags, void* residual, const void* w, void* out, int T, int H, int world,
                                  int nslabs, float eps, bool gemma, uint32_t* epoch_ptr, uint32_t* done_ctr, int num_sms,
                                  cudaStream_t stream);

}  // namespace sglb
