"""Every decoder family against its Hugging Face implementation: a tiny random HF model is written to disk with ``save_pretrained`` and served
by the engine from that directory (config parsing + checkpoint name mapping + the model code); prompt logprobs and greedy decode must agree.
(The reference's model tests compare against HF runners the same way: test/registered/models/test_generation_models.py.)"""
import numpy as np
import pytest
import torch

import sglang_b200 as sgl

tf = pytest.importorskip("transformers")

BASE = dict(hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2, vocab_size=512,
            max_position_embeddings=512, tie_word_embeddings=False)

CASES = {
    "llama": ("LlamaConfig", "LlamaForCausalLM", dict(BASE, rope_theta=10000.0)),
    "llama3-rope": ("LlamaConfig", "LlamaForCausalLM", dict(BASE, rope_theta=50000.0, rope_scaling={
        "rope_type": "llama3", "factor": 8.0, "low_freq_factor": 1.0, "high_freq_factor": 4.0, "original_max_position_embeddings": 64})),
    "llama-yarn": ("LlamaConfig", "LlamaForCausalLM", dict(BASE, rope_theta=10000.0, rope_scaling={
        "rope_type": "yarn", "factor": 4.0, "original_max_position_embeddings": 128, "beta_fast": 32.0, "beta_slow": 1.0})),
    "llama-linear-rope": ("LlamaConfig", "LlamaForCausalLM", dict(BASE, rope_theta=10000.0, rope_scaling={"rope_type": "linear", "factor": 2.0})),
    "phi3-longrope": ("Phi3Config", "Phi3ForCausalLM", dict(BASE, pad_token_id=0, original_max_position_embeddings=16, rope_scaling={
        "rope_type": "longrope", "short_factor": [1.0 + 0.1 * i for i in range(8)], "long_factor": [2.0 + 0.3 * i for i in range(8)]})),
    "deepseek-v2": ("DeepseekV2Config", "DeepseekV2ForCausalLM", dict(BASE, num_key_value_heads=4, kv_lora_rank=32, q_lora_rank=24, qk_nope_head_dim=16,
                                                                    qk_rope_head_dim=8, v_head_dim=16, n_routed_experts=8, num_experts_per_tok=2,
                                                                    moe_intermediate_size=32, n_shared_experts=1, first_k_dense_replace=1, n_group=2,
                                                                    topk_group=1, topk_method="group_limited_greedy", routed_scaling_factor=1.0,
                                                                    norm_topk_prob=False)),
    "nemotron-h": ("NemotronHConfig", "NemotronHForCausalLM", dict(hidden_size=64, intermediate_size=128, num_hidden_layers=4, num_attention_heads=4,
                                                                 num_key_value_heads=2, head_dim=16, vocab_size=512, max_position_embeddings=512,
                                                                 hybrid_override_pattern="MM*M", mamba_num_heads=8, mamba_head_dim=16, ssm_state_size=16,
                                                                 conv_kernel=4, n_groups=2, chunk_size=8, tie_word_embeddings=False)),
    "seed-oss": ("SeedOssConfig", "SeedOssForCausalLM", dict(BASE, head_dim=16, attention_bias=True, attention_out_bias=True)),
    "helium": ("HeliumConfig", "HeliumForCausalLM", dict(BASE, head_dim=16)),
    "ernie4.5": ("Ernie4_5Config", "Ernie4_5ForCausalLM", dict(BASE, head_dim=16, use_bias=True)),
    "arcee": ("ArceeConfig", "ArceeForCausalLM", dict(BASE, head_dim=16)),
    "smollm3": ("SmolLM3Config", "SmolLM3ForCausalLM", dict(BASE, pad_token_id=0, num_hidden_layers=4, no_rope_layers=[1, 1, 0, 1], use_sliding_window=True,
                                                            sliding_window=8, layer_types=["full_attention", "sliding_attention"] * 2)),
    "ministral3": ("Ministral3Config", "Ministral3ForCausalLM", dict(BASE, head_dim=16, sliding_window=None)),
    "olmo3": ("Olmo3Config", "Olmo3ForCausalLM", dict(BASE, sliding_window=8, layer_types=["sliding_attention", "full_attention"])),
    "glm": ("GlmConfig", "GlmForCausalLM", dict(BASE, head_dim=16, pad_token_id=0)),
    "glm4": ("Glm4Config", "Glm4ForCausalLM", dict(BASE, head_dim=16, pad_token_id=0)),
    "gemma3": ("Gemma3TextConfig", "Gemma3ForCausalLM", dict(BASE, num_hidden_layers=3, head_dim=16, sliding_window=8, query_pre_attn_scalar=16,
                                                             layer_types=["sliding_attention", "sliding_attention", "full_attention"],
                                                             tie_word_embeddings=True, rope_local_base_freq=10000.0, rope_theta=1000000.0)),
    "cohere": ("CohereConfig", "CohereForCausalLM", dict(BASE, pad_token_id=0, use_qk_norm=True, logit_scale=0.25, tie_word_embeddings=True)),
    "olmoe": ("OlmoeConfig", "OlmoeForCausalLM", dict(BASE, pad_token_id=0, num_experts=4, num_experts_per_tok=2)),
    "nemotron": ("NemotronConfig", "NemotronForCausalLM", dict(BASE, head_dim=16, partial_rotary_factor=0.5)),
    "phi": ("PhiConfig", "PhiForCausalLM", dict(hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4, vocab_size=512,
                                                max_position_embeddings=512, partial_rotary_factor=0.5)),
    "gptj": ("GPTJConfig", "GPTJForCausalLM", dict(n_embd=64, n_layer=2, n_head=4, vocab_size=512, n_positions=512, rotary_dim=8)),
    "gpt-bigcode": ("GPTBigCodeConfig", "GPTBigCodeForCausalLM", dict(n_embd=64, n_layer=2, n_head=4, vocab_size=512, n_positions=512, n_inner=128,
                                                                    multi_query=True)),
    "glm4-moe": ("Glm4MoeConfig", "Glm4MoeForCausalLM", dict(BASE, num_hidden_layers=3, head_dim=16, pad_token_id=0, n_routed_experts=8,
                                                             num_experts_per_tok=2, moe_intermediate_size=32, n_shared_experts=1, first_k_dense_replace=1,
                                                             n_group=2, topk_group=1, routed_scaling_factor=1.5, norm_topk_prob=True, use_qk_norm=True)),
    "dots1": ("Dots1Config", "Dots1ForCausalLM", dict(BASE, num_hidden_layers=3, head_dim=16, n_routed_experts=8, num_experts_per_tok=2,
                                                      moe_intermediate_size=32, n_shared_experts=1, first_k_dense_replace=1, n_group=2, topk_group=1,
                                                      routed_scaling_factor=1.5, norm_topk_prob=True)),
    "bamba": ("BambaConfig", "BambaForCausalLM", dict(BASE, num_hidden_layers=3, attn_layer_indices=[1], mamba_n_heads=8, mamba_d_head=16, mamba_n_groups=2,
                                                      mamba_d_state=16, mamba_d_conv=4, mamba_expand=2, mamba_chunk_size=8, pad_token_id=0,
                                                      rope_parameters={"rope_theta": 10000.0, "rope_type": "default", "partial_rotary_factor": 1.0})),  # (this transformers build rotates the full head for Bamba whatever the factor says)
    "mamba1": ("MambaConfig", "MambaForCausalLM", dict(hidden_size=64, num_hidden_layers=2, vocab_size=512, state_size=8, conv_kernel=4, expand=2,
                                                       time_step_rank=4, pad_token_id=0)),
    "falcon-mamba": ("FalconMambaConfig", "FalconMambaForCausalLM", dict(hidden_size=64, num_hidden_layers=2, vocab_size=512, state_size=8, conv_kernel=4,
                                                                       expand=2, time_step_rank=4, pad_token_id=0)),
    "jamba": ("JambaConfig", "JambaForCausalLM", dict(BASE, num_hidden_layers=4, attn_layer_period=2, attn_layer_offset=1, expert_layer_period=2,
                                                      expert_layer_offset=1, num_experts=4, num_experts_per_tok=2, mamba_d_state=8, mamba_d_conv=4,
                                                      mamba_expand=2, mamba_dt_rank=4, pad_token_id=0, use_mamba_kernels=False)),
    "qwen3-next": ("Qwen3NextConfig", "Qwen3NextForCausalLM", dict(BASE, num_hidden_layers=4, head_dim=16, partial_rotary_factor=0.25,
                                                                 linear_num_key_heads=2, linear_num_value_heads=4, linear_key_head_dim=8,
                                                                 linear_value_head_dim=8, linear_conv_kernel_dim=4, num_experts=4, num_experts_per_tok=2,
                                                                 moe_intermediate_size=32, shared_expert_intermediate_size=48, decoder_sparse_step=1,
                                                                 mlp_only_layers=[], norm_topk_prob=True,
                                                                 layer_types=["linear_attention", "linear_attention", "linear_attention", "full_attention"])),
    "qwen3.5": ("Qwen3_5TextConfig", "Qwen3_5ForCausalLM", dict(BASE, num_hidden_layers=4, head_dim=16, linear_num_key_heads=2, linear_num_value_heads=4,
                                                              linear_key_head_dim=8, linear_value_head_dim=8, linear_conv_kernel_dim=4,
                                                              layer_types=["linear_attention", "linear_attention", "linear_attention", "full_attention"])),
    "qwen3.5-moe": ("Qwen3_5MoeTextConfig", "Qwen3_5MoeForCausalLM", dict(BASE, num_hidden_layers=4, head_dim=16, linear_num_key_heads=2,
                                                                         linear_num_value_heads=4, linear_key_head_dim=8, linear_value_head_dim=8,
                                                                         linear_conv_kernel_dim=4, num_experts=4, num_experts_per_tok=2,
                                                                         moe_intermediate_size=32, shared_expert_intermediate_size=48,
                                                                         layer_types=["linear_attention", "linear_attention", "linear_attention", "full_attention"])),
    "granite4-hybrid": ("GraniteMoeHybridConfig", "GraniteMoeHybridForCausalLM", dict(
        BASE, num_hidden_layers=3, layer_types=["mamba", "attention", "mamba"], mamba_n_heads=8, mamba_d_head=16, mamba_n_groups=2, mamba_d_state=16,
        mamba_d_conv=4, mamba_expand=2, mamba_chunk_size=8, num_local_experts=4, num_experts_per_tok=2, shared_intermediate_size=48,
        embedding_multiplier=3.0, residual_multiplier=0.5, attention_multiplier=0.2, logits_scaling=2.0, position_embedding_type="nope", pad_token_id=0)),
    "llama4": ("Llama4TextConfig", "Llama4ForCausalLM", dict(BASE, num_hidden_layers=4, head_dim=16, intermediate_size=32, intermediate_size_mlp=96,
                                                           num_local_experts=4, num_experts_per_tok=1, interleave_moe_layer_step=2,
                                                           no_rope_layers=[1, 1, 1, 0], attention_chunk_size=8, use_qk_norm=True,
                                                           attn_temperature_tuning=True, floor_scale=4, attn_scale=0.3, pad_token_id=0)),
    "falcon-h1": ("FalconH1Config", "FalconH1ForCausalLM", dict(
        BASE, num_hidden_layers=2, mamba_n_heads=8, mamba_d_head=16, mamba_d_ssm=128, mamba_n_groups=2, mamba_d_state=16, mamba_d_conv=4, mamba_expand=2,
        mamba_chunk_size=8, mamba_rms_norm=True, mamba_norm_before_gate=False, ssm_multipliers=[0.5, 1.5, 0.7, 1.3, 0.9], ssm_in_multiplier=1.2,
        ssm_out_multiplier=0.8, attention_in_multiplier=1.1, attention_out_multiplier=0.9, key_multiplier=0.6, mlp_multipliers=[0.7, 1.4],
        embedding_multiplier=2.0, lm_head_multiplier=0.5, pad_token_id=0)),
    "falcon-h1-norm-before-gate": ("FalconH1Config", "FalconH1ForCausalLM", dict(
        BASE, num_hidden_layers=2, mamba_n_heads=8, mamba_d_head=16, mamba_d_ssm=128, mamba_n_groups=2, mamba_d_state=16, mamba_d_conv=4, mamba_expand=2,
        mamba_chunk_size=8, mamba_rms_norm=True, mamba_norm_before_gate=True, pad_token_id=0)),
    "falcon-h1-no-norm": ("FalconH1Config", "FalconH1ForCausalLM", dict(
        BASE, num_hidden_layers=2, mamba_n_heads=8, mamba_d_head=16, mamba_d_ssm=128, mamba_n_groups=1, mamba_d_state=16, mamba_d_conv=4, mamba_expand=2,
        mamba_chunk_size=8, mamba_rms_norm=False, pad_token_id=0)),
    "granite-moe": ("GraniteMoeConfig", "GraniteMoeForCausalLM", dict(BASE, num_local_experts=4, num_experts_per_tok=2, embedding_multiplier=3.0,
                                                                    residual_multiplier=0.5, attention_multiplier=0.2, logits_scaling=2.0, pad_token_id=0)),
    "granite-moe-shared": ("GraniteMoeSharedConfig", "GraniteMoeSharedForCausalLM", dict(BASE, num_local_experts=4, num_experts_per_tok=2,
                                                                                       shared_intermediate_size=48, embedding_multiplier=3.0,
                                                                                       residual_multiplier=0.5, attention_multiplier=0.2,
                                                                                       logits_scaling=2.0, pad_token_id=0)),
    "cwm": ("CwmConfig", "CwmForCausalLM", dict(BASE, pad_token_id=0, num_hidden_layers=4, sliding_window=8,
                                                layer_types=["full_attention", "sliding_attention", "sliding_attention", "full_attention"])),
    "cohere2": ("Cohere2Config", "Cohere2ForCausalLM", dict(BASE, pad_token_id=0, num_hidden_layers=4, sliding_window=8, logit_scale=0.25, tie_word_embeddings=True,
                                                            layer_types=["sliding_attention", "sliding_attention", "sliding_attention", "full_attention"])),
    "mistral": ("MistralConfig", "MistralForCausalLM", dict(BASE, sliding_window=None)),
    "qwen2": ("Qwen2Config", "Qwen2ForCausalLM", dict(BASE)),
    "qwen3": ("Qwen3Config", "Qwen3ForCausalLM", dict(BASE, head_dim=16)),
    "gemma": ("GemmaConfig", "GemmaForCausalLM", dict(BASE, head_dim=16, tie_word_embeddings=True)),
    "gemma2": ("Gemma2Config", "Gemma2ForCausalLM", dict(BASE, head_dim=16, sliding_window=8, query_pre_attn_scalar=16, tie_word_embeddings=True)),
    "vaultgemma": ("VaultGemmaConfig", "VaultGemmaForCausalLM", dict(BASE, head_dim=16, sliding_window=8, query_pre_attn_scalar=16, tie_word_embeddings=True,
                                                                    layer_types=["sliding_attention", "full_attention"])),
    "apertus": ("ApertusConfig", "ApertusForCausalLM", dict(BASE, hidden_act="xielu", pad_token_id=0)),
    "phi3": ("Phi3Config", "Phi3ForCausalLM", dict(BASE, pad_token_id=0)),
    "granite": ("GraniteConfig", "GraniteForCausalLM", dict(BASE, embedding_multiplier=3.0, residual_multiplier=0.5, attention_multiplier=0.2,
                                                            logits_scaling=2.0)),
    "olmo2": ("Olmo2Config", "Olmo2ForCausalLM", dict(BASE)),
    "olmo": ("OlmoConfig", "OlmoForCausalLM", dict(BASE, clip_qkv=0.9)),
    "exaone4": ("Exaone4Config", "Exaone4ForCausalLM", dict(BASE, head_dim=16, sliding_window=None, sliding_window_pattern=4,
                                                            layer_types=["full_attention", "full_attention"])),
    "exaone4-hybrid": ("Exaone4Config", "Exaone4ForCausalLM", dict(BASE, head_dim=16, num_hidden_layers=4, sliding_window=8, sliding_window_pattern=4,
                                                                   layer_types=["sliding_attention", "sliding_attention", "sliding_attention", "full_attention"])),
    "stablelm": ("StableLmConfig", "StableLmForCausalLM", dict(BASE, partial_rotary_factor=0.25, use_qkv_bias=True, layer_norm_eps=1e-5)),
    "mixtral": ("MixtralConfig", "MixtralForCausalLM", dict(BASE, num_local_experts=4, num_experts_per_tok=2)),
    "qwen2-moe": ("Qwen2MoeConfig", "Qwen2MoeForCausalLM", dict(BASE, num_experts=4, num_experts_per_tok=2, moe_intermediate_size=32,
                                                                shared_expert_intermediate_size=64, decoder_sparse_step=1, mlp_only_layers=[])),
    "qwen3-moe": ("Qwen3MoeConfig", "Qwen3MoeForCausalLM", dict(BASE, head_dim=16, num_experts=4, num_experts_per_tok=2, moe_intermediate_size=32,
                                                                decoder_sparse_step=1, mlp_only_layers=[], norm_topk_prob=True)),
    "gpt2": ("GPT2Config", "GPT2LMHeadModel", dict(n_embd=64, n_layer=2, n_head=4, vocab_size=512, n_positions=512)),
    "gpt-neo": ("GPTNeoConfig", "GPTNeoForCausalLM", dict(hidden_size=64, num_layers=4, num_heads=4, vocab_size=512, max_position_embeddings=512,
                                                        window_size=8, attention_types=[[["global", "local"], 2]])),
    "persimmon": ("PersimmonConfig", "PersimmonForCausalLM", dict(hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4,
                                                                  vocab_size=512, max_position_embeddings=512, qk_layernorm=True, partial_rotary_factor=0.5)),
    "codegen": ("CodeGenConfig", "CodeGenForCausalLM", dict(n_embd=64, n_layer=2, n_head=4, vocab_size=512, n_positions=512, rotary_dim=8)),
    "opt": ("OPTConfig", "OPTForCausalLM", dict(hidden_size=64, ffn_dim=128, num_hidden_layers=2, num_attention_heads=4, vocab_size=512,
                                                max_position_embeddings=512, word_embed_proj_dim=64)),
    "gpt-neox": ("GPTNeoXConfig", "GPTNeoXForCausalLM", dict(hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4,
                                                             vocab_size=512, max_position_embeddings=512, rotary_pct=0.5)),
    "starcoder2": ("Starcoder2Config", "Starcoder2ForCausalLM", dict(BASE, sliding_window=None, use_bias=True)),
    "gpt-oss": ("GptOssConfig", "GptOssForCausalLM", dict(BASE, head_dim=16, intermediate_size=32, num_local_experts=4, num_experts_per_tok=2,
                                                          sliding_window=8, layer_types=["sliding_attention", "full_attention"])),
    "deepseek-v3": ("DeepseekV3Config", "DeepseekV3ForCausalLM", dict(BASE, num_key_value_heads=4, kv_lora_rank=32, q_lora_rank=None, qk_nope_head_dim=16,
                                                                    qk_rope_head_dim=8, v_head_dim=16, n_routed_experts=8, num_experts_per_tok=2,
                                                                    moe_intermediate_size=32, n_shared_experts=1, first_k_dense_replace=1, n_group=2,
                                                                    topk_group=1, routed_scaling_factor=1.5, norm_topk_prob=True)),
    "minimax-m2": ("MiniMaxM2Config", "MiniMaxM2ForCausalLM", dict(BASE, num_local_experts=4, num_experts_per_tok=2, head_dim=16)),
    "ernie4.5-moe": ("Ernie4_5_MoeConfig", "Ernie4_5_MoeForCausalLM", dict(BASE, moe_num_experts=4, moe_k=2, moe_intermediate_size=32,
                                                                        moe_num_shared_experts=1, moe_layer_start_index=1, moe_layer_interval=1)),
    "glm4-moe-lite": ("Glm4MoeLiteConfig", "Glm4MoeLiteForCausalLM", dict(BASE, num_hidden_layers=3, num_key_value_heads=4, kv_lora_rank=32, q_lora_rank=24,
                                                                            qk_nope_head_dim=16, qk_rope_head_dim=8, v_head_dim=16, n_routed_experts=4,
                                                                            num_experts_per_tok=2, moe_intermediate_size=32, n_shared_experts=1,
                                                                            first_k_dense_replace=1, n_group=1, topk_group=1)),
    "hunyuan-dense": ("HunYuanDenseV1Config", "HunYuanDenseV1ForCausalLM", dict(BASE, head_dim=16)),
    "hunyuan-dense-ntk-alpha": ("HunYuanDenseV1Config", "HunYuanDenseV1ForCausalLM", dict(BASE, head_dim=16, rope_parameters={
        "rope_type": "dynamic", "rope_theta": 10000.0, "alpha": 50.0, "factor": 1.0})),
    "hunyuan-moe": ("HunYuanMoEV1Config", "HunYuanMoEV1ForCausalLM", dict(BASE, head_dim=16, num_experts=4, moe_topk=2, num_shared_expert=1)),
    "phimoe": ("PhimoeConfig", "PhimoeForCausalLM", dict(BASE, num_local_experts=4, num_experts_per_tok=2, router_jitter_noise=0.2)),
    "flex-olmo": ("FlexOlmoConfig", "FlexOlmoForCausalLM", dict(BASE, num_experts=4, num_experts_per_tok=2, pad_token_id=0)),
    "mamba2": ("Mamba2Config", "Mamba2ForCausalLM", dict(hidden_size=64, num_hidden_layers=2, vocab_size=512, num_heads=8, head_dim=16, state_size=16,
                                                         n_groups=2, expand=2, conv_kernel=4, chunk_size=8, tie_word_embeddings=True)),
}


@pytest.mark.parametrize("name", list(CASES))
def test_family_matches_hf(name, tmp_path):
    cfg_cls, model_cls, kw = CASES[name]
    if not hasattr(tf, cfg_cls) or not hasattr(tf, model_cls):
        pytest.skip(f"transformers has no {model_cls}")
    torch.manual_seed(0)
    try:
        hf = getattr(tf, model_cls)(getattr(tf, cfg_cls)(**kw)).float().eval()
        with torch.no_grad():
            hf(input_ids=torch.tensor([[1, 2, 3]]))
    except Exception as e:      # the installed transformers build cannot run this family itself
        pytest.skip(f"HF {model_cls} unusable here: {type(e).__name__}: {e}")
    with torch.no_grad():
        for n, p in hf.named_parameters():      # default inits are degenerate (tiny weights, unit norms): make every tensor matter
            if p.dim() > 1:
                p.normal_(0.0, 0.08)
            elif "norm" in n or "ln_" in n or "layernorm" in n:
                p.add_(torch.randn_like(p) * 0.1)
            elif n.endswith("bias") or "sinks" in n or n.endswith(".D") or "dt_bias" in n:
                p.add_(torch.randn_like(p) * 0.05)
    d = tmp_path / name
    hf.save_pretrained(str(d), safe_serialization=True)
    ids = np.random.default_rng(1).integers(3, 500, size=29).tolist()
    eng = sgl.Engine(model_path=str(d), device="cpu", page_size=4, skip_tokenizer_init=True, chunked_prefill_size=16, log_level="error")
    o = eng.generate(input_ids=ids, sampling_params={"max_new_tokens": 6, "temperature": 0, "ignore_eos": True}, return_logprob=True,
                     logprob_start_len=0)
    eng.shutdown()
    full = ids + o["output_ids"]
    with torch.no_grad():
        lp = torch.log_softmax(hf(input_ids=torch.tensor([full])).logits[0].float(), -1)
    ref = [float(lp[t, full[t + 1]]) for t in range(len(ids) - 1)]
    got = [v for v, _, _ in o["meta_info"]["input_token_logprobs"][1:]]
    assert max(abs(a - b) for a, b in zip(ref, got)) < 3e-3, name
    dec = [v for v, _, _ in o["meta_info"]["output_token_logprobs"]]
    n = len(ids)
    for j, tok in enumerate(o["output_ids"]):
        assert abs(float(lp[n - 1 + j, tok]) - dec[j]) < 3e-3, (name, j)


def test_llava_matches_hf(tmp_path):
    """LLaVA-1.5 layout: CLIP tower (feature layer -2, CLS dropped) + 2-layer projector + Llama; preprocessing, image-row splicing and the
    checkpoint dialect of the installed transformers"""
    from sglang_b200.multimodal import processor as P

    torch.manual_seed(0)
    cfg = tf.LlavaConfig(text_config=dict(model_type="llama", hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4,
                                          num_key_value_heads=2, vocab_size=512, max_position_embeddings=1024, rope_theta=10000.0),
                         vision_config=dict(model_type="clip_vision_model", hidden_size=32, intermediate_size=64, num_hidden_layers=3,
                                            num_attention_heads=2, image_size=28, patch_size=14, projection_dim=32),
                         image_token_index=500, vision_feature_layer=-2, vision_feature_select_strategy="default", projector_hidden_act="gelu")
    hf = tf.LlavaForConditionalGeneration(cfg).float().eval()
    with torch.no_grad():
        for n, p in hf.named_parameters():
            if p.dim() > 1:
                p.normal_(0, 0.08)
    d = str(tmp_path / "llava")
    hf.save_pretrained(d, safe_serialization=True)
    img = np.random.default_rng(0).integers(0, 255, size=(40, 50, 3), dtype=np.uint8)
    ids = [5, 6, 500, 7, 8, 9]
    e = sgl.Engine(model_path=d, device="cpu", page_size=4, skip_tokenizer_init=True, log_level="error", disable_radix_cache=True)
    o = e.generate(input_ids=ids, image_data=[img], sampling_params={"max_new_tokens": 4, "temperature": 0, "ignore_eos": True},
                   return_logprob=True, logprob_start_len=0)
    ex, mm = P.process_images(ids, [img], e.model_config, None)
    e.shutdown()
    hf_ids, n = list(ex), mm["num_image_tokens"]
    for off in mm["image_offsets"]:
        hf_ids[off:off + n] = [500] * n
    full = hf_ids + o["output_ids"]
    with torch.no_grad():
        lp = torch.log_softmax(hf(input_ids=torch.tensor([full]), pixel_values=torch.tensor(mm["pixel_values"])).logits[0].float(), -1)
    got = [v for v, _, _ in o["meta_info"]["input_token_logprobs"][1:]]
    ref = [float(lp[t, full[t + 1]]) for t in range(len(hf_ids) - 1)]
    assert max(abs(ref[t] - got[t]) for t in range(len(hf_ids) - 1) if hf_ids[t + 1] != 500) < 2e-3
    dec = [v for v, _, _ in o["meta_info"]["output_token_logprobs"]]
    assert max(abs(float(lp[len(hf_ids) - 1 + j, tok]) - dec[j]) for j, tok in enumerate(o["output_ids"])) < 2e-3


def test_internvl_matches_hf(tmp_path):
    """InternViT (CLS, layer scale, qk-norm off) -> pixel shuffle -> projector -> Qwen2; two images, one tiled 2x1 + thumbnail"""
    from sglang_b200.multimodal import processor as P

    torch.manual_seed(0)
    cfg = tf.InternVLConfig(text_config=dict(model_type="qwen2", hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4,
                                             num_key_value_heads=2, vocab_size=512, max_position_embeddings=2048, rope_theta=10000.0),
                            vision_config=dict(hidden_size=32, intermediate_size=64, num_hidden_layers=2, num_attention_heads=2, image_size=[56, 56],
                                               patch_size=[14, 14], use_qk_norm=False, norm_type="layer_norm", layer_scale_init_value=0.5),
                            image_token_id=500, downsample_ratio=0.5, vision_feature_layer=-1)
    hf = tf.InternVLForConditionalGeneration(cfg).float().eval()
    with torch.no_grad():
        for n, p in hf.named_parameters():
            if p.dim() > 1 or "lambda" in n or "cls_token" in n:
                p.normal_(0, 0.08) if p.dim() > 1 else p.add_(torch.randn_like(p) * 0.1)
    d = str(tmp_path / "internvl")
    hf.save_pretrained(d, safe_serialization=True)
    rng = np.random.default_rng(0)
    imgs = [rng.integers(0, 255, size=(50, 110, 3), dtype=np.uint8), rng.integers(0, 255, size=(60, 60, 3), dtype=np.uint8)]
    ids = [5, 6, 500, 7, 8, 500, 9, 10]
    e = sgl.Engine(model_path=d, device="cpu", page_size=4, skip_tokenizer_init=True, log_level="error", disable_radix_cache=True,
                   chunked_prefill_size=16)
    o = e.generate(input_ids=ids, image_data=imgs, sampling_params={"max_new_tokens": 4, "temperature": 0, "ignore_eos": True},
                   return_logprob=True, logprob_start_len=0)
    ex, mm = P.process_images(ids, imgs, e.model_config, None)
    e.shutdown()
    assert mm["image_token_counts"] == [3 * 4, 1 * 4]          # 2x1 grid + thumbnail, then a single tile; 4 tokens per 56-px tile
    hf_ids = list(ex)
    for off, n in zip(mm["image_offsets"], mm["image_token_counts"]):
        hf_ids[off:off + n] = [500] * n
    full = hf_ids + o["output_ids"]
    with torch.no_grad():
        lp = torch.log_softmax(hf(input_ids=torch.tensor([full]), pixel_values=torch.from_numpy(mm["pixel_values"])).logits[0].float(), -1)
    got = [v for v, _, _ in o["meta_info"]["input_token_logprobs"][1:]]
    ref = [float(lp[t, full[t + 1]]) for t in range(len(hf_ids) - 1)]
    assert max(abs(ref[t] - got[t]) for t in range(len(hf_ids) - 1) if hf_ids[t + 1] != 500) < 2e-3
    dec = [v for v, _, _ in o["meta_info"]["output_token_logprobs"]]
    assert max(abs(float(lp[len(hf_ids) - 1 + j, tok]) - dec[j]) for j, tok in enumerate(o["output_ids"])) < 2e-3
