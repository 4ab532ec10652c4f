"""ModelConfig: architecture description + built-in presets (no network: presets stand in for HF hub configs).

Parity: reference ``python/sglang/srt/configs/model_config.py`` (arch detection, context length, head dims).
``model_path`` may be ``dummy:<preset>`` (random-init weights of that architecture, reference
``DummyModelLoader`` ``model_loader/loader.py:1602``) or a local directory with an HF ``config.json``.
"""
from __future__ import annotations

import logging

import dataclasses
import json
import os
from dataclasses import dataclass
from typing import Any, Dict, List, Optional

logger = logging.getLogger(__name__)


@dataclass
class ModelConfig:
    architecture: str = "LlamaForCausalLM"
    hidden_size: int = 4096
    intermediate_size: int = 14336
    num_hidden_layers: int = 32
    num_attention_heads: int = 32
    num_key_value_heads: int = 8
    head_dim: Optional[int] = None
    vocab_size: int = 128256
    rms_norm_eps: float = 1e-5
    rope_theta: float = 500000.0
    rope_scaling: Optional[Dict[str, Any]] = None
    max_position_embeddings: int = 8192
    tie_word_embeddings: bool = False
    attention_bias: bool = False
    mlp_bias: bool = False
    hidden_act: str = "silu"
    qk_norm: bool = False
    sliding_window: Optional[int] = None
    attn_logit_softcapping: float = 0.0
    final_logit_softcapping: float = 0.0
    rope_is_neox: bool = True
    partial_rotary_factor: float = 1.0
    # MoE
    num_local_experts: int = 0
    num_experts_per_tok: int = 0
    moe_intermediate_size: Optional[int] = None
    norm_topk_prob: bool = True
    shared_expert_intermediate_size: Optional[int] = None      # Qwen2-MoE sigmoid-gated shared expert
    decoder_sparse_step: int = 1
    mlp_only_layers: Optional[List[int]] = None
    # classic pre-LayerNorm decoders (GPT-2 / OPT / GPT-NeoX / StarCoder2)
    pos_emb: str = "rope"            # rope | learned
    pos_offset: int = 0              # OPT stores positions shifted by 2
    parallel_residual: bool = False  # NeoX: x + attn(ln1 x) + mlp(ln2 x)
    # DeepSeek-V2/V3: MLA + DeepSeekMoE
    q_lora_rank: Optional[int] = None
    kv_lora_rank: Optional[int] = None
    qk_nope_head_dim: int = 0
    qk_rope_head_dim: int = 0
    v_head_dim: int = 0
    n_shared_experts: Optional[int] = None
    first_k_dense_replace: int = 0
    moe_layer_freq: int = 1
    n_group: Optional[int] = None
    topk_group: Optional[int] = None
    routed_scaling_factor: float = 1.0
    scoring_func: str = "softmax"
    topk_method: str = "greedy"
    swiglu_limit: Optional[float] = None        # GPT-OSS clamped SwiGLU
    # Llama-variant switches (SeedOss / Ernie-4.5 / Helium / Arcee / SmolLM3 / OLMo-3 / GLM / Gemma-3 / Cohere / Nemotron ...)
    attention_out_bias: Optional[bool] = None   # o_proj bias (defaults to no bias)
    mlp_gated: bool = True                      # False: up_proj -> act -> down_proj (Arcee, Nemotron)
    no_rope_layers: Optional[List[int]] = None  # SmolLM3: 0 = this layer applies no rotary embedding
    rope_local_base_freq: Optional[float] = None   # Gemma-3: rope base of the sliding-window layers
    logit_scale: Optional[float] = None         # Cohere
    clip_qkv: Optional[float] = None            # OLMo: clamp the fused q/k/v projection
    norm_type: str = "rms"                      # rms | layernorm (Cohere: weight only) | layernorm1p (Nemotron: 1 + weight, bias)
    use_qk_norm: bool = False                   # Cohere-2 style flag in HF configs
    # Gemma-2 / Granite extras
    query_pre_attn_scalar: Optional[float] = None
    layer_types: Optional[List[str]] = None
    embedding_multiplier: Optional[float] = None
    residual_multiplier: Optional[float] = None
    attention_multiplier: Optional[float] = None
    logits_scaling: Optional[float] = None
    # Llama-4: no_rope_layers, moe_layers, intermediate_size_mlp, attention_chunk_size, use_qk_norm, attn_temperature_tuning, floor_scale, attn_scale
    llama4: Optional[Dict[str, Any]] = None
    falcon_h1: Optional[Dict[str, Any]] = None       # Falcon-H1 muP multipliers (folded into the weights at load)
    # Qwen3-Next gated DeltaNet layers: num_k_heads, num_v_heads, k_head_dim, v_head_dim, conv_kernel (models/qwen3_next.py)
    linear_attn: Optional[Dict[str, Any]] = None
    # state-space / hybrid decoders (models/mamba.py): num_heads, head_dim, d_state, d_conv, n_groups, norm_groups, chunk_size, pattern
    mamba: Optional[Dict[str, Any]] = None
    # architectures without a native class: {"path", "arch"} -> models/transformers_fallback.py runs the Hugging Face module tree
    router_jitter_noise: Optional[float] = None      # Phi-3.5-MoE sparsemixer band
    hf_fallback: Optional[Dict[str, Any]] = None
    num_labels: int = 1                              # classification / reward heads
    encoder: Optional[Dict[str, Any]] = None         # encoder-only models (BERT / RoBERTa): type_vocab_size, sep_token_id
    # speech encoder of encoder-decoder ASR models (models/whisper.py)
    audio: Optional[Dict[str, Any]] = None
    # multimodal (LLaVA-style)
    vision: Optional[Dict[str, Any]] = None
    image_token_id: Optional[int] = None
    # special tokens
    bos_token_id: int = 1
    eos_token_id: Any = 2
    pad_token_id: int = 0
    dtype: str = "bfloat16"
    name: str = ""

    def __post_init__(self) -> None:
        self.attn_logit_softcapping = float(self.attn_logit_softcapping or 0.0)        # HF writes an explicit null for "off"
        self.final_logit_softcapping = float(self.final_logit_softcapping or 0.0)
        if self.head_dim is None:
            self.head_dim = self.hidden_size // self.num_attention_heads
        if self.moe_intermediate_size is None:
            self.moe_intermediate_size = self.intermediate_size

    @property
    def eos_token_ids(self) -> List[int]:
        e = self.eos_token_id
        if e is None:          # config.json without an EOS id (generation_config.json / the tokenizer supply it): nothing stops on its own
            return []
        return [int(x) for x in e] if isinstance(e, (list, tuple)) else [int(e)]

    @property
    def is_moe(self) -> bool:
        return self.num_local_experts > 0

    @property
    def is_multimodal(self) -> bool:
        return self.vision is not None

    def num_params(self) -> int:
        h, i, L, V = self.hidden_size, self.intermediate_size, self.num_hidden_layers, self.vocab_size
        qkv = h * (self.num_attention_heads + 2 * self.num_key_value_heads) * self.head_dim
        o = self.num_attention_heads * self.head_dim * h
        mlp = 3 * h * (self.moe_intermediate_size if self.is_moe else i) * (self.num_local_experts if self.is_moe else 1)
        emb = V * h * (1 if self.tie_word_embeddings else 2)
        return L * (qkv + o + mlp + 2 * h) + emb + h


_LLAMA3_ROPE = {"rope_type": "llama3", "factor": 8.0, "low_freq_factor": 1.0, "high_freq_factor": 4.0,
                "original_max_position_embeddings": 8192}
_CLIP_L_336 = {"hidden_size": 1024, "intermediate_size": 4096, "num_hidden_layers": 24, "num_attention_heads": 16,
               "image_size": 336, "patch_size": 14, "select_layer": -2, "projector_hidden_act": "gelu"}

PRESETS: Dict[str, Dict[str, Any]] = {
    # ---- tiny CPU-test models
    "tiny-llama": dict(hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2,
                       vocab_size=512, max_position_embeddings=512, rope_theta=10000.0, dtype="float32"),
    "tiny-qwen2-vl": dict(architecture="Qwen2VLForConditionalGeneration", hidden_size=64, intermediate_size=128, num_hidden_layers=2,
                          num_attention_heads=4, num_key_value_heads=2, head_dim=16, vocab_size=512, max_position_embeddings=2048, attention_bias=True,
                          rope_theta=10000.0, dtype="float32", image_token_id=500, rope_scaling={"type": "mrope", "mrope_section": [2, 3, 3]},
                          vision={"embed_dim": 32, "depth": 2, "num_heads": 2, "hidden_size": 64, "mlp_ratio": 2, "patch_size": 14,
                                  "spatial_merge_size": 2, "temporal_patch_size": 2, "min_pixels": 28 * 28, "max_pixels": 28 * 28 * 12}),
    "tiny-whisper": dict(architecture="WhisperForConditionalGeneration", hidden_size=64, intermediate_size=128, num_hidden_layers=2,
                         num_attention_heads=4, num_key_value_heads=4, head_dim=16, vocab_size=512, max_position_embeddings=128, pos_emb="learned",
                         hidden_act="gelu", tie_word_embeddings=True, dtype="float32",
                         audio={"d_model": 64, "encoder_layers": 2, "encoder_attention_heads": 4, "encoder_ffn_dim": 128, "num_mel_bins": 80,
                                "max_source_positions": 100, "chunk_length": 2}),
    "tiny-internvl": dict(architecture="InternVLForConditionalGeneration", hidden_size=64, intermediate_size=128, num_hidden_layers=2,
                          num_attention_heads=4, num_key_value_heads=2, head_dim=16, vocab_size=512, max_position_embeddings=2048, attention_bias=True,
                          rope_theta=10000.0, dtype="float32", image_token_id=500,
                          vision={"kind": "internvl", "hidden_size": 32, "intermediate_size": 64, "num_hidden_layers": 2, "num_attention_heads": 2,
                                  "image_size": 56, "patch_size": 14, "downsample_ratio": 0.5, "select_layer": -1, "max_tiles": 4}),
    "tiny-gpt-oss": dict(architecture="GptOssForCausalLM", hidden_size=64, intermediate_size=32, moe_intermediate_size=32, num_hidden_layers=4,
                         num_attention_heads=4, num_key_value_heads=2, head_dim=16, vocab_size=512, max_position_embeddings=512, dtype="float32",
                         num_local_experts=4, num_experts_per_tok=2, sliding_window=8, attention_bias=True, swiglu_limit=7.0, rope_theta=150000.0,
                         layer_types=["sliding_attention", "full_attention", "sliding_attention", "full_attention"],
                         rope_scaling={"rope_type": "yarn", "factor": 4.0, "beta_fast": 32.0, "beta_slow": 1.0, "truncate": False,
                                       "original_max_position_embeddings": 128}),
    "tiny-mamba2": dict(architecture="Mamba2ForCausalLM", hidden_size=64, intermediate_size=128, num_hidden_layers=3, num_attention_heads=4,
                        num_key_value_heads=2, vocab_size=512, max_position_embeddings=512, dtype="float32", tie_word_embeddings=True,
                        mamba=dict(num_heads=8, head_dim=16, d_state=16, d_conv=4, n_groups=1, norm_groups=1, chunk_size=8)),
    "tiny-nemotron-h": dict(architecture="NemotronHForCausalLM", hidden_size=64, intermediate_size=128, num_hidden_layers=6,
                            num_attention_heads=4, num_key_value_heads=2, head_dim=16, vocab_size=512, max_position_embeddings=512,
                            dtype="float32", mamba=dict(num_heads=8, head_dim=16, d_state=16, d_conv=4, n_groups=2, norm_groups=2,
                                                        chunk_size=8, pattern="M-M*-M")),
    "tiny-llama-gpu": dict(hidden_size=512, intermediate_size=1024, num_hidden_layers=2, num_attention_heads=4,
                           num_key_value_heads=2, head_dim=128, vocab_size=2048, max_position_embeddings=2048, rope_theta=10000.0),
    "tiny-swa-gpu": dict(hidden_size=512, intermediate_size=1024, num_hidden_layers=4, num_attention_heads=4, num_key_value_heads=2, head_dim=128,
                         vocab_size=2048, max_position_embeddings=2048, rope_theta=10000.0, sliding_window=128,
                         layer_types=["sliding_attention", "full_attention", "sliding_attention", "full_attention"]),
    "tiny-llama-gpu8": dict(hidden_size=1024, intermediate_size=2048, num_hidden_layers=2, num_attention_heads=8,
                            num_key_value_heads=2, head_dim=128, vocab_size=2048, max_position_embeddings=2048, rope_theta=10000.0),
    "tiny-mixtral": dict(architecture="MixtralForCausalLM", hidden_size=64, intermediate_size=128, num_hidden_layers=2,
                         num_attention_heads=4, num_key_value_heads=2, vocab_size=512, max_position_embeddings=512,
                         num_local_experts=4, num_experts_per_tok=2, rope_theta=10000.0, dtype="float32"),
    "tiny-mixtral-gpu": dict(architecture="MixtralForCausalLM", hidden_size=512, intermediate_size=1024, num_hidden_layers=2,
                             num_attention_heads=4, num_key_value_heads=2, head_dim=128, vocab_size=2048,
                             max_position_embeddings=2048, num_local_experts=8, num_experts_per_tok=2, rope_theta=10000.0),
    "tiny-qwen2": dict(architecture="Qwen2ForCausalLM", hidden_size=64, intermediate_size=128, num_hidden_layers=2,
                       num_attention_heads=4, num_key_value_heads=2, vocab_size=512, max_position_embeddings=512,
                       attention_bias=True, rope_theta=10000.0, dtype="float32"),
    "tiny-qwen3": dict(architecture="Qwen3ForCausalLM", hidden_size=64, intermediate_size=128, num_hidden_layers=2,
                       num_attention_heads=4, num_key_value_heads=2, head_dim=32, vocab_size=512, max_position_embeddings=512,
                       qk_norm=True, rope_theta=10000.0, dtype="float32"),
    "tiny-gemma": dict(architecture="GemmaForCausalLM", hidden_size=64, intermediate_size=128, num_hidden_layers=2,
                       num_attention_heads=4, num_key_value_heads=1, head_dim=32, vocab_size=512, max_position_embeddings=512,
                       hidden_act="gelu_pytorch_tanh", tie_word_embeddings=True, rope_theta=10000.0, dtype="float32"),
    "tiny-llava": dict(architecture="LlavaForConditionalGeneration", hidden_size=64, intermediate_size=128, num_hidden_layers=2,
                       num_attention_heads=4, num_key_value_heads=2, vocab_size=512, max_position_embeddings=1024,
                       rope_theta=10000.0, dtype="float32", image_token_id=500,
                       vision={"hidden_size": 32, "intermediate_size": 64, "num_hidden_layers": 2, "num_attention_heads": 2,
                               "image_size": 28, "patch_size": 14, "select_layer": -2, "projector_hidden_act": "gelu"}),
    "tiny-deepseek-v2": dict(architecture="DeepseekV2ForCausalLM", hidden_size=64, intermediate_size=128, num_hidden_layers=3,
                             num_attention_heads=4, num_key_value_heads=4, vocab_size=512, max_position_embeddings=512, rope_theta=10000.0,
                             kv_lora_rank=32, q_lora_rank=None, qk_nope_head_dim=16, qk_rope_head_dim=8, v_head_dim=16,
                             num_local_experts=8, num_experts_per_tok=2, moe_intermediate_size=32, n_shared_experts=1,
                             first_k_dense_replace=1, n_group=4, topk_group=2, routed_scaling_factor=1.0, norm_topk_prob=False,
                             rope_is_neox=False, dtype="float32"),
    "tiny-deepseek-v3": dict(architecture="DeepseekV3ForCausalLM", hidden_size=64, intermediate_size=128, num_hidden_layers=3,
                             num_attention_heads=4, num_key_value_heads=4, vocab_size=512, max_position_embeddings=512, rope_theta=10000.0,
                             kv_lora_rank=32, q_lora_rank=24, qk_nope_head_dim=16, qk_rope_head_dim=8, v_head_dim=16,
                             num_local_experts=8, num_experts_per_tok=2, moe_intermediate_size=32, n_shared_experts=1,
                             first_k_dense_replace=1, n_group=4, topk_group=2, routed_scaling_factor=2.5, norm_topk_prob=True,
                             scoring_func="sigmoid", topk_method="noaux_tc", rope_is_neox=False, dtype="float32"),
    "tiny-gpt2": dict(architecture="GPT2LMHeadModel", hidden_size=64, intermediate_size=256, num_hidden_layers=2, num_attention_heads=4,
                      num_key_value_heads=4, vocab_size=512, max_position_embeddings=512, pos_emb="learned", tie_word_embeddings=True,
                      hidden_act="gelu_new", dtype="float32"),
    "tiny-opt": dict(architecture="OPTForCausalLM", hidden_size=64, intermediate_size=256, num_hidden_layers=2, num_attention_heads=4,
                     num_key_value_heads=4, vocab_size=512, max_position_embeddings=512, pos_emb="learned", pos_offset=2,
                     tie_word_embeddings=True, hidden_act="relu", dtype="float32"),
    "tiny-neox": dict(architecture="GPTNeoXForCausalLM", hidden_size=64, intermediate_size=256, num_hidden_layers=2, num_attention_heads=4,
                      num_key_value_heads=4, vocab_size=512, max_position_embeddings=512, partial_rotary_factor=0.25, parallel_residual=True,
                      hidden_act="gelu", rope_theta=10000.0, dtype="float32"),
    "tiny-starcoder2": dict(architecture="Starcoder2ForCausalLM", hidden_size=64, intermediate_size=256, num_hidden_layers=2,
                            num_attention_heads=4, num_key_value_heads=2, vocab_size=512, max_position_embeddings=512,
                            hidden_act="gelu_pytorch_tanh", rope_theta=10000.0, sliding_window=64, tie_word_embeddings=True, dtype="float32"),
    "tiny-phi3": dict(architecture="Phi3ForCausalLM", hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4,
                      num_key_value_heads=4, vocab_size=512, max_position_embeddings=512, rope_theta=10000.0, dtype="float32"),
    "tiny-internlm2": dict(architecture="InternLM2ForCausalLM", hidden_size=64, intermediate_size=128, num_hidden_layers=2,
                           num_attention_heads=4, num_key_value_heads=2, vocab_size=512, max_position_embeddings=512,
                           rope_theta=10000.0, dtype="float32"),
    "tiny-qwen2-moe": dict(architecture="Qwen2MoeForCausalLM", hidden_size=64, intermediate_size=128, num_hidden_layers=3,
                           num_attention_heads=4, num_key_value_heads=2, vocab_size=512, max_position_embeddings=512,
                           attention_bias=True, num_local_experts=4, num_experts_per_tok=2, moe_intermediate_size=32,
                           shared_expert_intermediate_size=64, norm_topk_prob=False, mlp_only_layers=[0], rope_theta=10000.0,
                           dtype="float32"),
    "tiny-qwen3-moe": dict(architecture="Qwen3MoeForCausalLM", hidden_size=64, intermediate_size=128, num_hidden_layers=2,
                           num_attention_heads=4, num_key_value_heads=2, head_dim=32, vocab_size=512, max_position_embeddings=512,
                           qk_norm=True, num_local_experts=8, num_experts_per_tok=2, moe_intermediate_size=32, rope_theta=10000.0,
                           dtype="float32"),
    "tiny-gemma2": dict(architecture="Gemma2ForCausalLM", hidden_size=64, intermediate_size=128, num_hidden_layers=4,
                        num_attention_heads=4, num_key_value_heads=2, head_dim=32, vocab_size=512, max_position_embeddings=512,
                        hidden_act="gelu_pytorch_tanh", tie_word_embeddings=True, sliding_window=16, attn_logit_softcapping=50.0,
                        final_logit_softcapping=30.0, query_pre_attn_scalar=32, rope_theta=10000.0, dtype="float32"),
    "tiny-granite": dict(architecture="GraniteForCausalLM", hidden_size=64, intermediate_size=128, num_hidden_layers=2,
                         num_attention_heads=4, num_key_value_heads=2, vocab_size=512, max_position_embeddings=512,
                         embedding_multiplier=12.0, residual_multiplier=0.22, attention_multiplier=0.0078125 * 8, logits_scaling=8.0,
                         tie_word_embeddings=True, rope_theta=10000.0, dtype="float32"),
    "tiny-olmo2": dict(architecture="Olmo2ForCausalLM", hidden_size=64, intermediate_size=128, num_hidden_layers=2,
                       num_attention_heads=4, num_key_value_heads=2, vocab_size=512, max_position_embeddings=512, rope_theta=10000.0,
                       dtype="float32"),
    # ---- real architectures (random-init offline)
    "deepseek-v2-lite": dict(architecture="DeepseekV2ForCausalLM", hidden_size=2048, intermediate_size=10944, num_hidden_layers=27,
                             num_attention_heads=16, num_key_value_heads=16, vocab_size=102400, max_position_embeddings=163840,
                             rope_theta=10000.0, kv_lora_rank=512, q_lora_rank=None, qk_nope_head_dim=128, qk_rope_head_dim=64,
                             v_head_dim=128, num_local_experts=64, num_experts_per_tok=6, moe_intermediate_size=1408, n_shared_experts=2,
                             first_k_dense_replace=1, n_group=1, topk_group=1, routed_scaling_factor=1.0, norm_topk_prob=False,
                             rope_is_neox=False, rms_norm_eps=1e-6, bos_token_id=100000, eos_token_id=100001),
    "qwen3-30b-a3b": dict(architecture="Qwen3MoeForCausalLM", hidden_size=2048, intermediate_size=6144, num_hidden_layers=48,
                          num_attention_heads=32, num_key_value_heads=4, head_dim=128, vocab_size=151936, max_position_embeddings=40960,
                          rope_theta=1000000.0, qk_norm=True, rms_norm_eps=1e-6, num_local_experts=128, num_experts_per_tok=8,
                          moe_intermediate_size=768, eos_token_id=[151643, 151645]),
    "gemma-2-9b": dict(architecture="Gemma2ForCausalLM", hidden_size=3584, intermediate_size=14336, num_hidden_layers=42,
                       num_attention_heads=16, num_key_value_heads=8, head_dim=256, vocab_size=256000, max_position_embeddings=8192,
                       hidden_act="gelu_pytorch_tanh", tie_word_embeddings=True, sliding_window=4096, attn_logit_softcapping=50.0,
                       final_logit_softcapping=30.0, query_pre_attn_scalar=256, rope_theta=10000.0, rms_norm_eps=1e-6),
    "phi-3-mini": dict(architecture="Phi3ForCausalLM", hidden_size=3072, intermediate_size=8192, num_hidden_layers=32,
                       num_attention_heads=32, num_key_value_heads=32, vocab_size=32064, max_position_embeddings=4096, rope_theta=10000.0),
    "llama-3-8b": dict(hidden_size=4096, intermediate_size=14336, num_hidden_layers=32, num_attention_heads=32,
                       num_key_value_heads=8, vocab_size=128256, max_position_embeddings=8192, rope_theta=500000.0,
                       bos_token_id=128000, eos_token_id=[128001, 128009]),
    # per-rank shapes of llama-3-8b at TP=4 / TP=8 on ONE GPU (no collectives): kernel repro / stress presets (tools/attn_stress.py)
    "llama-3-8b-tp4shard": dict(hidden_size=4096, intermediate_size=3584, num_hidden_layers=32, num_attention_heads=8, head_dim=128,
                                num_key_value_heads=2, vocab_size=32064, max_position_embeddings=8192, rope_theta=500000.0,
                                bos_token_id=1, eos_token_id=[2]),
    "llama-3-8b-tp8shard": dict(hidden_size=4096, intermediate_size=1792, num_hidden_layers=32, num_attention_heads=4, head_dim=128,
                                num_key_value_heads=1, vocab_size=16032, max_position_embeddings=8192, rope_theta=500000.0,
                                bos_token_id=1, eos_token_id=[2]),
    "llama-3.1-8b": dict(hidden_size=4096, intermediate_size=14336, num_hidden_layers=32, num_attention_heads=32,
                         num_key_value_heads=8, vocab_size=128256, max_position_embeddings=131072, rope_theta=500000.0,
                         rope_scaling=_LLAMA3_ROPE, bos_token_id=128000, eos_token_id=[128001, 128008, 128009]),
    "llama-3-70b": dict(hidden_size=8192, intermediate_size=28672, num_hidden_layers=80, num_attention_heads=64,
                        num_key_value_heads=8, vocab_size=128256, max_position_embeddings=8192, rope_theta=500000.0,
                        bos_token_id=128000, eos_token_id=[128001, 128009]),
    "llama-2-7b": dict(hidden_size=4096, intermediate_size=11008, num_hidden_layers=32, num_attention_heads=32,
                       num_key_value_heads=32, vocab_size=32000, max_position_embeddings=4096, rope_theta=10000.0),
    "mixtral-8x7b": dict(architecture="MixtralForCausalLM", hidden_size=4096, intermediate_size=14336, num_hidden_layers=32,
                         num_attention_heads=32, num_key_value_heads=8, vocab_size=32000, max_position_embeddings=32768,
                         rope_theta=1000000.0, num_local_experts=8, num_experts_per_tok=2),
    "qwen2-7b": dict(architecture="Qwen2ForCausalLM", hidden_size=3584, intermediate_size=18944, num_hidden_layers=28,
                     num_attention_heads=28, num_key_value_heads=4, vocab_size=152064, max_position_embeddings=32768,
                     rope_theta=1000000.0, attention_bias=True, rms_norm_eps=1e-6, eos_token_id=[151643, 151645]),
    "qwen3-8b": dict(architecture="Qwen3ForCausalLM", hidden_size=4096, intermediate_size=12288, num_hidden_layers=36,
                     num_attention_heads=32, num_key_value_heads=8, head_dim=128, vocab_size=151936,
                     max_position_embeddings=40960, rope_theta=1000000.0, qk_norm=True, rms_norm_eps=1e-6,
                     eos_token_id=[151643, 151645]),
    "gemma-2b": dict(architecture="GemmaForCausalLM", hidden_size=2048, intermediate_size=16384, num_hidden_layers=18,
                     num_attention_heads=8, num_key_value_heads=1, head_dim=256, vocab_size=256000,
                     max_position_embeddings=8192, hidden_act="gelu_pytorch_tanh", tie_word_embeddings=True,
                     rope_theta=10000.0, rms_norm_eps=1e-6),
    "llava-1.5-7b": dict(architecture="LlavaForConditionalGeneration", hidden_size=4096, intermediate_size=11008,
                         num_hidden_layers=32, num_attention_heads=32, num_key_value_heads=32, vocab_size=32064,
                         max_position_embeddings=4096, rope_theta=10000.0, image_token_id=32000, vision=_CLIP_L_336),
}

_HF_ARCH_MAP = {"MistralForCausalLM": "LlamaForCausalLM", "SeedOssForCausalLM": "LlamaForCausalLM", "HeliumForCausalLM": "LlamaForCausalLM",
                "Ernie4_5ForCausalLM": "LlamaForCausalLM", "ArceeForCausalLM": "LlamaForCausalLM", "SmolLM3ForCausalLM": "LlamaForCausalLM",
                "Ministral3ForCausalLM": "LlamaForCausalLM", "Olmo3ForCausalLM": "Olmo2ForCausalLM", "KimiK2ForCausalLM": "DeepseekV3ForCausalLM",
                "CwmForCausalLM": "LlamaForCausalLM", "Cohere2ForCausalLM": "CohereForCausalLM",
                "Glm4MoeLiteForCausalLM": "DeepseekV3ForCausalLM",
                # Llama / Qwen2 bodies under other names (reference models/llama.py EntryClass list, models/mimo.py): extra heads in the
                # checkpoint (MiMo's ``model.mtp_layers.*``) are not part of the served model and are skipped by the loader
                "InternLM3ForCausalLM": "LlamaForCausalLM", "IQuestCoderForCausalLM": "LlamaForCausalLM", "MiMoForCausalLM": "Qwen2ForCausalLM"}          # GLM-4.7-Flash: MLA + DeepSeek-V3 style MoE under GLM names
# HF config keys that differ from our field names
_HF_KEY_MAP = {"num_experts": "num_local_experts", "n_routed_experts": "num_local_experts", "n_embd": "hidden_size", "n_layer": "num_hidden_layers",
               "n_head": "num_attention_heads", "n_positions": "max_position_embeddings", "n_inner": "intermediate_size",
               "activation_function": "hidden_act", "layer_norm_epsilon": "rms_norm_eps", "layer_norm_eps": "rms_norm_eps",
               "norm_epsilon": "rms_norm_eps", "ffn_dim": "intermediate_size", "rotary_pct": "partial_rotary_factor",
               "use_qkv_bias": "attention_bias", "qk_layernorm": "qk_norm", "use_parallel_residual": "parallel_residual", "rotary_emb_base": "rope_theta", "experts_per_token": "num_experts_per_tok"}
# per-architecture defaults for fields HF configs leave implicit
_ARCH_DEFAULTS = {"GPT2LMHeadModel": dict(pos_emb="learned", tie_word_embeddings=True, hidden_act="gelu_new"),
                  "OPTForCausalLM": dict(pos_emb="learned", pos_offset=2, tie_word_embeddings=True, hidden_act="relu"),
                  "GPTNeoXForCausalLM": dict(parallel_residual=True, hidden_act="gelu"),
                  "Starcoder2ForCausalLM": dict(hidden_act="gelu_pytorch_tanh"),
                  "PhiForCausalLM": dict(parallel_residual=True, norm_type="shared_layernorm", hidden_act="gelu_new", partial_rotary_factor=0.5),
                  "GPTJForCausalLM": dict(parallel_residual=True, norm_type="shared_layernorm", hidden_act="gelu_new", rope_is_neox=False),
                  "GPTBigCodeForCausalLM": dict(pos_emb="learned", tie_word_embeddings=True, hidden_act="gelu_pytorch_tanh"),
                  "CodeGenForCausalLM": dict(parallel_residual=True, norm_type="shared_layernorm", hidden_act="gelu_new", rope_is_neox=False),
                  "GPTNeoForCausalLM": dict(pos_emb="learned", tie_word_embeddings=True, hidden_act="gelu_new"),
                  "PersimmonForCausalLM": dict(hidden_act="relu2", partial_rotary_factor=0.5),
                  "Qwen2ForCausalLM": dict(attention_bias=True), "Qwen2ForSequenceClassification": dict(attention_bias=True),
                  "Qwen2Model": dict(attention_bias=True), "Qwen3ForSequenceClassification": dict(qk_norm=True),
                  "Qwen2ForRewardModel": dict(attention_bias=True), "Qwen3ForRewardModel": dict(qk_norm=True), "Qwen3Model": dict(qk_norm=True), "Qwen2MoeForCausalLM": dict(attention_bias=True),
                  "Qwen2VLForConditionalGeneration": dict(attention_bias=True), "Qwen2_5_VLForConditionalGeneration": dict(attention_bias=True),
                  "Glm4MoeForCausalLM": dict(scoring_func="sigmoid", topk_method="noaux_tc", partial_rotary_factor=0.5, attention_bias=True),
                  "Dots1ForCausalLM": dict(scoring_func="sigmoid", topk_method="noaux_tc", use_qk_norm=True),
                  "Gemma3ForCausalLM": dict(hidden_act="gelu_pytorch_tanh", tie_word_embeddings=True),
                  "OlmoeForCausalLM": dict(norm_topk_prob=False), "FlexOlmoForCausalLM": dict(norm_topk_prob=False), "MiniMaxM2ForCausalLM": dict(norm_topk_prob=True),
                  "PhimoeForCausalLM": dict(attention_bias=True, attention_out_bias=True), "PhiMoEForCausalLM": dict(attention_bias=True, attention_out_bias=True),
                  "HeliumForCausalLM": dict(rope_is_neox=False), "Ernie4_5ForCausalLM": dict(rope_is_neox=False),
                  "Ernie4_5_MoeForCausalLM": dict(rope_is_neox=False, scoring_func="softmax", topk_method="noaux_tc", norm_topk_prob=True),
                  "ArceeForCausalLM": dict(mlp_gated=False, hidden_act="relu2"), "ApertusForCausalLM": dict(mlp_gated=False, qk_norm=True),
                  "GlmForCausalLM": dict(rope_is_neox=False, partial_rotary_factor=0.5, attention_bias=True),
                  "Glm4ForCausalLM": dict(rope_is_neox=False, partial_rotary_factor=0.5, attention_bias=True),
                  "NemotronForCausalLM": dict(mlp_gated=False, hidden_act="relu2", norm_type="layernorm1p", partial_rotary_factor=0.5),
                  "CohereForCausalLM": dict(norm_type="layernorm", rope_is_neox=False, tie_word_embeddings=True, logit_scale=0.0625),
                  "DeepseekV3ForCausalLM": dict(scoring_func="sigmoid", topk_method="noaux_tc", rope_is_neox=False),
                  "DeepseekV2ForCausalLM": dict(rope_is_neox=False),
                  "Qwen3ForCausalLM": dict(qk_norm=True), "Qwen3MoeForCausalLM": dict(qk_norm=True)}


def _has_native_class(arch: str) -> bool:
    from ..models.registry import supported_architectures

    # speculative draft heads are built by their workers (speculative/eagle.py), not through the registry
    if arch.endswith(("Eagle", "Eagle3", "NextN", "MTP")) or "Eagle" in arch:
        return True
    return arch in supported_architectures()


def _transformers_fallback_config(model_path: str, arch: str, kw: Dict[str, Any]) -> Dict[str, Any]:
    """geometry of an architecture served through ``models/transformers_fallback.py``: read through ``AutoConfig`` so that every family's
    own key names (``n_embd``, ``num_heads`` ...) resolve to the standard attributes"""
    import transformers as tf

    t = tf.AutoConfig.from_pretrained(model_path).get_text_config()
    nh = t.num_attention_heads
    out = {k: v for k, v in kw.items() if k in ("rope_theta", "rope_scaling", "tie_word_embeddings", "sliding_window", "final_logit_softcapping",
                                               "bos_token_id", "eos_token_id", "pad_token_id", "dtype", "name", "rms_norm_eps")}
    out.update(architecture="TransformersForCausalLM", hf_fallback={"path": model_path, "arch": arch}, hidden_size=t.hidden_size,
               num_hidden_layers=t.num_hidden_layers, num_attention_heads=nh, num_key_value_heads=getattr(t, "num_key_value_heads", None) or nh,
               head_dim=getattr(t, "head_dim", None) or t.hidden_size // nh, vocab_size=t.vocab_size,
               intermediate_size=getattr(t, "intermediate_size", None) or 4 * t.hidden_size,
               max_position_embeddings=getattr(t, "max_position_embeddings", 8192))
    return out


def load_model_config(model_path: str, overrides: Optional[Dict[str, Any]] = None, context_length: Optional[int] = None) -> ModelConfig:
    overrides = dict(overrides or {})
    impl = overrides.pop("model_impl", "auto")
    if model_path.startswith("dummy:"):
        name = model_path.split(":", 1)[1]
        if name not in PRESETS:
            raise ValueError(f"unknown preset {name!r}; available: {sorted(PRESETS)}")
        kw = {**PRESETS[name], **overrides, "name": name}
    else:
        if os.path.isfile(model_path) and model_path.endswith(".gguf"):
            from ..model_loader.gguf import GGUFFile, config_from_gguf

            hf = config_from_gguf(GGUFFile(model_path))
        else:
            cfg_file = os.path.join(model_path, "config.json")
            if not os.path.exists(cfg_file):
                raise FileNotFoundError(f"{cfg_file} not found (offline box: use dummy:<preset>, a local HF directory or a .gguf file)")
            with open(cfg_file) as f:
                hf = json.load(f)
        text = hf.get("text_config", hf)
        fields = {f.name for f in dataclasses.fields(ModelConfig)}
        text = {_HF_KEY_MAP.get(k, k): v for k, v in text.items()}
        rp = text.get("rope_parameters")          # transformers >= 5: theta and scaling live in one dict
        if isinstance(rp, dict) and "full_attention" in rp:      # Gemma-3: one entry per layer type
            loc = rp.get("sliding_attention") or {}
            if loc.get("rope_theta") is not None:
                text.setdefault("rope_local_base_freq", loc["rope_theta"])
            rp = rp["full_attention"]
        if isinstance(rp, dict):
            if rp.get("rope_theta") is not None:
                text.setdefault("rope_theta", rp["rope_theta"])
            if rp.get("partial_rotary_factor") is not None:
                text["partial_rotary_factor"] = rp["partial_rotary_factor"]          # authoritative in transformers >= 5
            sc = {k: v for k, v in rp.items() if k not in ("rope_theta", "partial_rotary_factor")}
            if text.get("rope_scaling") is None and (sc.get("rope_type", "default") != "default" or "mrope_section" in sc):
                text["rope_scaling"] = sc
        kw = {k: v for k, v in text.items() if k in fields}
        arch = (hf.get("architectures") or ["LlamaForCausalLM"])[0]
        if "text_config" in hf and hf.get("tie_word_embeddings") is not None:
            kw["tie_word_embeddings"] = bool(hf["tie_word_embeddings"])          # composite configs decide at the top level
        kw["architecture"] = _HF_ARCH_MAP.get(arch, arch)
        kw.setdefault("rope_theta", 10000.0)          # the Hugging Face default when a config does not say (this dataclass defaults to Llama-3's)
        for k, v in (_ARCH_DEFAULTS.get(arch) or _ARCH_DEFAULTS.get(kw["architecture"], {})).items():      # aliases inherit their base family's implicit fields
            kw.setdefault(k, v)
        if arch in _ARCH_DEFAULTS:
            kw.setdefault("num_key_value_heads", kw.get("num_attention_heads", 1))
            if kw.get("intermediate_size") is None:
                kw["intermediate_size"] = 4 * kw["hidden_size"]
        if arch in ("Llama4ForCausalLM", "Llama4ForConditionalGeneration"):
            L = text["num_hidden_layers"]
            nr = text.get("no_rope_layers") or [int((i + 1) % 4 != 0) for i in range(L)]
            step = text.get("interleave_moe_layer_step", 1)
            kw["llama4"] = {"no_rope_layers": nr, "moe_layers": text.get("moe_layers") or list(range(step - 1, L, step)),
                            "intermediate_size_mlp": text.get("intermediate_size_mlp"), "attention_chunk_size": text.get("attention_chunk_size", 8192),
                            "use_qk_norm": text.get("use_qk_norm", True), "attn_temperature_tuning": text.get("attn_temperature_tuning", True),
                            "floor_scale": text.get("floor_scale", 8192), "attn_scale": text.get("attn_scale", 0.1)}
            kw["layer_types"] = text.get("layer_types") or ["chunked_attention" if r else "full_attention" for r in nr]
            kw["rope_is_neox"] = False
            kw["architecture"] = "Llama4ForCausalLM"
        if arch in ("Qwen3NextForCausalLM", "Qwen3_5ForCausalLM", "Qwen3_5MoeForCausalLM", "Qwen3_5ForConditionalGeneration",
                    "Qwen3_5MoeForConditionalGeneration"):
            kw["linear_attn"] = {"num_k_heads": text.get("linear_num_key_heads", 16), "num_v_heads": text.get("linear_num_value_heads", 32),
                                 "k_head_dim": text.get("linear_key_head_dim", 128), "v_head_dim": text.get("linear_value_head_dim", 128),
                                 "conv_kernel": text.get("linear_conv_kernel_dim", 4)}
            L = text["num_hidden_layers"]
            kw["layer_types"] = text.get("layer_types") or ["full_attention" if (i + 1) % text.get("full_attention_interval", 4) == 0
                                                            else "linear_attention" for i in range(L)]
            kw.setdefault("partial_rotary_factor", 0.25)
        if arch in ("MambaForCausalLM", "FalconMambaForCausalLM"):
            hid = text["hidden_size"]
            kw["mamba"] = {"d_inner": text.get("intermediate_size") or text.get("expand", 2) * hid, "d_state": text.get("state_size", 16),
                           "d_conv": text.get("conv_kernel", 4), "dt_rank": text.get("time_step_rank") if isinstance(text.get("time_step_rank"), int)
                           else -(-hid // 16), "use_bias": text.get("use_bias", False), "use_conv_bias": text.get("use_conv_bias", True),
                           "bcdt_norm": "plain" if arch == "FalconMambaForCausalLM" else None, "mixer_rms_eps": text.get("mixer_rms_eps", 1e-6),
                           "pattern": "M" * text["num_hidden_layers"]}
            kw["rms_norm_eps"] = text.get("layer_norm_epsilon", 1e-5)
            kw["intermediate_size"] = kw["mamba"]["d_inner"]
            kw.setdefault("num_attention_heads", 1)
            kw.setdefault("tie_word_embeddings", True)
        if arch == "JambaForCausalLM":
            hid, L = text["hidden_size"], text["num_hidden_layers"]
            kinds = text.get("layers_block_type") or ["attention" if i % text.get("attn_layer_period", 8) == text.get("attn_layer_offset", 4) else "mamba"
                                                      for i in range(L)]
            ne = text.get("layers_num_experts") or [text.get("num_local_experts", text.get("num_experts", 16)) if i % text.get("expert_layer_period", 2) == text.get("expert_layer_offset", 1)
                                                    else 1 for i in range(L)]
            dtr = text.get("mamba_dt_rank", "auto")
            kw["mamba"] = {"d_inner": text.get("mamba_expand", 2) * hid, "d_state": text.get("mamba_d_state", 16), "d_conv": text.get("mamba_d_conv", 4),
                           "dt_rank": -(-hid // 16) if dtr in ("auto", None) else dtr, "use_bias": text.get("mamba_proj_bias", False),
                           "use_conv_bias": text.get("mamba_conv_bias", True), "bcdt_norm": "weight",
                           "pattern": "".join("M" if k == "mamba" else "*" for k in kinds), "moe_layers": [int(n > 1) for n in ne]}
            kw["num_local_experts"] = max(ne) if max(ne) > 1 else 0
            kw["num_experts_per_tok"] = text.get("num_experts_per_tok", 2)
        if arch == "FalconH1ForCausalLM":
            kw["falcon_h1"] = {k: text.get(k) for k in ("ssm_multipliers", "ssm_in_multiplier", "ssm_out_multiplier", "attention_in_multiplier",
                                                        "attention_out_multiplier", "key_multiplier", "mlp_multipliers", "lm_head_multiplier")
                               if text.get(k) is not None}
        if arch in ("GraniteMoeForCausalLM", "GraniteMoeSharedForCausalLM"):      # Granite 3.x MoE = the Granite-4 block without Mamba layers
            kw["mamba"] = {"pattern": "*" * text["num_hidden_layers"], "num_heads": 1, "head_dim": 1, "d_state": 1, "d_conv": 2, "n_groups": 1}
            kw["shared_expert_intermediate_size"] = text.get("shared_intermediate_size") if arch == "GraniteMoeSharedForCausalLM" else None
            kw["pos_emb"] = "rope"
            kw["architecture"] = "GraniteMoeHybridForCausalLM"
        if arch == "Cohere2ForCausalLM":             # Command R7B / A: rotary only on the sliding-window layers, global layers are NoPE
            lt = text.get("layer_types") or [("full_attention" if (i + 1) % text.get("sliding_window_pattern", 4) == 0 else "sliding_attention")
                                             for i in range(text["num_hidden_layers"])]
            kw["layer_types"] = lt
            kw["no_rope_layers"] = [int(t == "sliding_attention") for t in lt]
        if arch == "Exaone4ForCausalLM":             # per-head q/k norm; hybrid checkpoints (32B): rotary on the sliding layers only, global layers NoPE
            kw["qk_norm"] = True
            if text.get("sliding_window"):
                lt = text.get("layer_types") or ["sliding_attention"] * text["num_hidden_layers"]
                kw["layer_types"] = lt
                kw["no_rope_layers"] = [int(t == "sliding_attention") for t in lt]
        if arch == "GraniteMoeHybridForCausalLM":
            kw["shared_expert_intermediate_size"] = text.get("shared_intermediate_size")
            kw["pos_emb"] = "rope" if text.get("position_embedding_type", "rope") == "rope" else "none"
        if arch in ("BambaForCausalLM", "GraniteMoeHybridForCausalLM", "FalconH1ForCausalLM"):
            hid = text["hidden_size"]
            heads = text.get("mamba_n_heads", 128)
            dh = text.get("mamba_d_head", "auto")
            dh = (text.get("mamba_expand", 2) * hid // heads) if dh in ("auto", None) else dh
            kinds = text.get("layers_block_type") or text.get("layer_types") or ["attention" if i in (text.get("attn_layer_indices") or []) else "mamba"
                                                      for i in range(text["num_hidden_layers"])]
            kw["mamba"] = {"num_heads": heads, "head_dim": dh, "d_state": text.get("mamba_d_state", 256), "d_conv": text.get("mamba_d_conv", 4),
                           "n_groups": text.get("mamba_n_groups", 1), "chunk_size": text.get("mamba_chunk_size", 256),
                           "use_bias": text.get("mamba_proj_bias", False), "use_conv_bias": text.get("mamba_conv_bias", True), "norm_groups": 1,
                           "pattern": "".join("M" if k == "mamba" else "*" for k in kinds)}
            kw["partial_rotary_factor"] = (text.get("rope_parameters") or {}).get("partial_rotary_factor", text.get("partial_rotary_factor", 0.5))
            if arch == "FalconH1ForCausalLM":
                kw["mamba"].update(norm_groups=text.get("mamba_n_groups", 1), rms_norm=bool(text.get("mamba_rms_norm", False)),
                                   norm_before_gate=bool(text.get("mamba_norm_before_gate", True)), pattern="M" * text["num_hidden_layers"])
                kw["partial_rotary_factor"] = (text.get("rope_parameters") or {}).get("partial_rotary_factor", 1.0)
                if text.get("mamba_d_ssm"):
                    kw["mamba"]["num_heads"] = text["mamba_d_ssm"] // kw["mamba"]["head_dim"]
        if arch in ("NemotronHForCausalLM", "Mamba2ForCausalLM"):
            hid = text["hidden_size"]
            heads = text.get("mamba_num_heads") or text.get("num_heads") or (text.get("expand", 2) * hid // text.get("head_dim", 64))
            hd = text.get("mamba_head_dim") or text.get("head_dim") or (text.get("expand", 2) * hid // heads)
            ngroups = text.get("n_groups", 1)
            kw["mamba"] = {"num_heads": heads, "head_dim": hd, "d_state": text.get("ssm_state_size") or text.get("state_size", 128),
                           "d_conv": text.get("conv_kernel", 4), "n_groups": ngroups, "chunk_size": text.get("chunk_size", 64),
                           "use_bias": text.get("use_bias", False), "use_conv_bias": text.get("use_conv_bias", True),
                           "norm_groups": ngroups if arch == "NemotronHForCausalLM" else 1,
                           "pattern": None, "dt_limit": list(text.get("time_step_limit") or (0.0, float("inf")))}
            if arch == "NemotronHForCausalLM":
                pat = text.get("hybrid_override_pattern")
                if not pat and text.get("layers_block_type"):
                    pat = "".join({"mamba": "M", "attention": "*", "mlp": "-", "moe": "E"}[t] for t in text["layers_block_type"])
                if pat and "E" in pat:
                    raise ValueError("Nemotron-H with MoE blocks ('E' in the hybrid pattern) is not supported yet")
                kw["mamba"]["pattern"] = pat
                if pat:
                    kw["num_hidden_layers"] = len(pat)
                kw["num_local_experts"] = kw["num_experts_per_tok"] = 0
                kw["mamba"]["use_bias"] = text.get("mamba_proj_bias", text.get("use_bias", False))
            kw["rms_norm_eps"] = text.get("layer_norm_epsilon", text.get("rms_norm_eps", 1e-5))
            kw.setdefault("num_key_value_heads", text.get("num_key_value_heads", kw.get("num_attention_heads", 1)))
            if arch == "NemotronHForCausalLM" and text.get("attention_head_dim"):
                kw["head_dim"] = text["attention_head_dim"]
        if hf.get("num_labels") is not None or hf.get("id2label"):
            kw["num_labels"] = int(hf.get("num_labels") or len(hf["id2label"]))
        if arch in ("BertModel", "XLMRobertaModel", "RobertaModel", "BertForSequenceClassification", "XLMRobertaForSequenceClassification",
                    "RobertaForSequenceClassification", "Contriever"):
            kw["encoder"] = {"type_vocab_size": hf.get("type_vocab_size", 2),
                             "sep_token_id": hf.get("sep_token_id", 102 if arch.startswith(("Bert", "Contriever")) else 2)}
            kw["rms_norm_eps"] = hf.get("layer_norm_eps", 1e-12)
            kw["num_key_value_heads"] = hf["num_attention_heads"]
            kw["pos_emb"] = "learned"
            kw.setdefault("hidden_act", hf.get("hidden_act", "gelu"))
        if arch == "WhisperForConditionalGeneration":
            kw.update(hidden_size=hf["d_model"], num_hidden_layers=hf["decoder_layers"], num_attention_heads=hf["decoder_attention_heads"],
                      num_key_value_heads=hf["decoder_attention_heads"], intermediate_size=hf["decoder_ffn_dim"],
                      max_position_embeddings=hf.get("max_target_positions", 448), hidden_act=hf.get("activation_function", "gelu"),
                      rms_norm_eps=1e-5, pos_emb="learned", tie_word_embeddings=True, head_dim=hf["d_model"] // hf["decoder_attention_heads"])
            kw["audio"] = {k: hf[k] for k in ("d_model", "encoder_layers", "encoder_attention_heads", "encoder_ffn_dim", "num_mel_bins",
                                              "max_source_positions")}
        if arch == "Qwen2AudioForConditionalGeneration":
            ac = hf.get("audio_config") or {}
            kw["audio"] = {"llm": True, "d_model": ac.get("d_model", 1280), "encoder_layers": ac.get("encoder_layers", 32),
                           "encoder_attention_heads": ac.get("encoder_attention_heads", 20), "encoder_ffn_dim": ac.get("encoder_ffn_dim", 5120),
                           "num_mel_bins": ac.get("num_mel_bins", 128), "max_source_positions": ac.get("max_source_positions", 1500),
                           "audio_token_id": hf.get("audio_token_id", hf.get("audio_token_index", 151646))}
            kw.setdefault("attention_bias", True)
        if arch == "VoxtralForConditionalGeneration":
            ac = hf.get("audio_config") or {}
            d_, inter = ac.get("hidden_size", 1280), ac.get("intermediate_size", 5120)
            kw["audio"] = {"llm": True, "d_model": d_, "encoder_layers": ac.get("num_hidden_layers", 32), "encoder_attention_heads": ac.get("num_attention_heads", 20),
                           "encoder_ffn_dim": inter, "num_mel_bins": ac.get("num_mel_bins", 128), "max_source_positions": ac.get("max_source_positions", 1500),
                           "audio_token_id": hf.get("audio_token_id", 24), "group": inter // d_,
                           "tokens_per_window": ac.get("max_source_positions", 1500) // (inter // d_)}
        if arch == "InternVLForConditionalGeneration":
            vc = hf["vision_config"]
            size = vc.get("image_size", 448)
            kw["vision"] = {"kind": "internvl", "hidden_size": vc["hidden_size"], "intermediate_size": vc["intermediate_size"],
                            "num_hidden_layers": vc["num_hidden_layers"], "num_attention_heads": vc["num_attention_heads"],
                            "image_size": size[0] if isinstance(size, (list, tuple)) else size,
                            "patch_size": (vc.get("patch_size", 14)[0] if isinstance(vc.get("patch_size", 14), (list, tuple)) else vc.get("patch_size", 14)),
                            "attention_bias": vc.get("attention_bias", False), "use_qk_norm": vc.get("use_qk_norm", False),
                            "norm_type": vc.get("norm_type", "layer_norm"), "layer_norm_eps": vc.get("layer_norm_eps", 1e-6),
                            "hidden_act": vc.get("hidden_act", "gelu"), "select_layer": hf.get("vision_feature_layer", -1),
                            "downsample_ratio": hf.get("downsample_ratio", 0.5)}
            kw["image_token_id"] = hf.get("image_token_id", 151667)
            if (hf.get("text_config") or {}).get("model_type", "qwen2") == "qwen2":
                kw.setdefault("attention_bias", True)
        if arch == "Qwen2_5_VLForConditionalGeneration":
            vc = hf["vision_config"]
            kw["vision"] = {"variant": "qwen2.5", "embed_dim": vc.get("hidden_size", 1280), "depth": vc.get("depth", 32),
                            "num_heads": vc.get("num_heads", 16), "hidden_size": vc.get("out_hidden_size", kw["hidden_size"]),
                            "intermediate_size": vc.get("intermediate_size", 3420), "patch_size": vc.get("patch_size", 14),
                            "spatial_merge_size": vc.get("spatial_merge_size", 2), "temporal_patch_size": vc.get("temporal_patch_size", 2),
                            "window_size": vc.get("window_size", 112), "fullatt_block_indexes": vc.get("fullatt_block_indexes", [7, 15, 23, 31])}
            kw["image_token_id"] = hf.get("image_token_id", 151655)
            kw.setdefault("attention_bias", True)
        if arch == "Gemma3ForConditionalGeneration" and hf.get("vision_config") is not None:
            vc = hf["vision_config"]
            kw["vision"] = {"variant": "siglip", "hidden_size": vc.get("hidden_size", 1152), "intermediate_size": vc.get("intermediate_size", 4304),
                            "num_hidden_layers": vc.get("num_hidden_layers", 27), "num_attention_heads": vc.get("num_attention_heads", 16),
                            "patch_size": vc.get("patch_size", 14), "image_size": vc.get("image_size", 896), "layer_norm_eps": vc.get("layer_norm_eps", 1e-6),
                            "tokens_per_image": hf.get("mm_tokens_per_image", 256), "resize_mode": "squash", "resize_filter": 2,
                            "image_mean": [0.5, 0.5, 0.5], "image_std": [0.5, 0.5, 0.5]}
            kw["image_token_id"] = hf.get("image_token_id", hf.get("image_token_index", 262144))
        if arch in ("Glm4vForConditionalGeneration", "Glm4vMoeForConditionalGeneration"):
            vc = hf["vision_config"]
            kw["vision"] = {"variant": "glm4v", "embed_dim": vc.get("hidden_size", 1536), "depth": vc.get("depth", 24), "num_heads": vc.get("num_heads", 12),
                            "hidden_size": vc.get("out_hidden_size", kw["hidden_size"]), "out_hidden_size": vc.get("out_hidden_size", kw["hidden_size"]),
                            "intermediate_size": vc.get("intermediate_size", 13696), "patch_size": vc.get("patch_size", 14),
                            "image_size": vc.get("image_size", 336), "spatial_merge_size": vc.get("spatial_merge_size", 2),
                            "temporal_patch_size": vc.get("temporal_patch_size", 2), "rms_norm_eps": vc.get("rms_norm_eps", 1e-5),
                            "attention_bias": vc.get("attention_bias", False), "min_pixels": 112 * 112, "max_pixels": 14 * 14 * 2 * 2 * 2 * 6144}
            kw["image_token_id"] = hf.get("image_token_id", 151343)
            for k_, v_ in _ARCH_DEFAULTS["Glm4MoeForCausalLM" if "Moe" in arch else "Glm4ForCausalLM"].items():
                kw.setdefault(k_, v_)
        if arch == "LlavaForConditionalGeneration" and (hf.get("vision_config") or {}).get("model_type") == "pixtral":
            arch = kw["architecture"] = "PixtralForConditionalGeneration"          # pixtral-12b ships in the LLaVA container
        if arch in ("LightOnOcrForConditionalGeneration", "LightOnOCRForConditionalGeneration"):
            kw["qk_norm"] = True          # Qwen3 decoder
        if arch in ("Mistral3ForConditionalGeneration", "PixtralForConditionalGeneration", "LightOnOcrForConditionalGeneration",
                    "LightOnOCRForConditionalGeneration"):
            vc = hf["vision_config"]
            vrp = vc.get("rope_parameters") or {}
            kw["vision"] = {"variant": "pixtral", "hidden_size": vc.get("hidden_size", 1024), "intermediate_size": vc.get("intermediate_size", 4096),
                            "num_hidden_layers": vc.get("num_hidden_layers", 24), "num_attention_heads": vc.get("num_attention_heads", 16),
                            "head_dim": vc.get("head_dim"), "patch_size": vc.get("patch_size", 14), "image_size": vc.get("image_size", 1540),
                            "hidden_act": vc.get("hidden_act", "gelu"),
                            "rope_theta": vrp.get("rope_theta", vc.get("rope_theta", 10000.0)),
                            "spatial_merge_size": 1 if arch.startswith("Pixtral") else hf.get("spatial_merge_size", 2),
                            "projector_hidden_act": hf.get("projector_hidden_act", "gelu"),
                            "projector_bias": hf.get("multimodal_projector_bias", arch.startswith("Pixtral")),
                            "image_break_token_id": hf.get("image_break_token_id", 12), "image_end_token_id": hf.get("image_end_token_id", 13)}
            kw["image_token_id"] = hf.get("image_token_id", hf.get("image_token_index", 10))
        if arch in ("Qwen3VLForConditionalGeneration", "Qwen3VLMoeForConditionalGeneration"):
            vc = hf["vision_config"]
            kw["vision"] = {"variant": "qwen3", "embed_dim": vc.get("hidden_size", 1152), "depth": vc.get("depth", 27), "num_heads": vc.get("num_heads", 16),
                            "hidden_size": vc.get("out_hidden_size", kw["hidden_size"]), "intermediate_size": vc.get("intermediate_size", 4304),
                            "patch_size": vc.get("patch_size", 16), "spatial_merge_size": vc.get("spatial_merge_size", 2),
                            "temporal_patch_size": vc.get("temporal_patch_size", 2), "num_position_embeddings": vc.get("num_position_embeddings", 2304),
                            "deepstack_visual_indexes": vc.get("deepstack_visual_indexes", [8, 16, 24]), "image_mean": [0.5, 0.5, 0.5],
                            "image_std": [0.5, 0.5, 0.5], "min_pixels": 64 * 64, "max_pixels": 16 * 16 * 4 * 4096}
            kw["image_token_id"] = hf.get("image_token_id", 151655)
            kw["qk_norm"] = True
        if arch == "Qwen2VLForConditionalGeneration":
            vc = hf["vision_config"]
            kw["vision"] = {"embed_dim": vc.get("embed_dim", 1280), "depth": vc.get("depth", 32), "num_heads": vc.get("num_heads", 16),
                            "hidden_size": vc.get("hidden_size", kw["hidden_size"]), "mlp_ratio": vc.get("mlp_ratio", 4),
                            "patch_size": vc.get("patch_size", 14), "spatial_merge_size": vc.get("spatial_merge_size", 2),
                            "temporal_patch_size": vc.get("temporal_patch_size", 2)}
            kw["image_token_id"] = hf.get("image_token_id", 151655)
            kw.setdefault("attention_bias", True)
        elif "vision_config" in hf and arch not in ("Qwen2_5_VLForConditionalGeneration", "InternVLForConditionalGeneration", "Qwen3VLForConditionalGeneration", "Mistral3ForConditionalGeneration", "PixtralForConditionalGeneration", "LightOnOcrForConditionalGeneration", "LightOnOCRForConditionalGeneration",
                                                 "Gemma3ForConditionalGeneration", "Glm4vForConditionalGeneration", "Glm4vMoeForConditionalGeneration",
                                                 "Qwen3VLMoeForConditionalGeneration"):
            kw["vision"] = {**_CLIP_L_336, **hf["vision_config"], "select_layer": hf.get("vision_feature_layer", -2),
                            "projector_hidden_act": hf.get("projector_hidden_act", "gelu")}
            kw["image_token_id"] = hf.get("image_token_index", hf.get("image_token_id"))
        kw["dtype"] = {"float16": "float16", "float32": "float32"}.get(hf.get("torch_dtype") or hf.get("dtype") or text.get("dtype") or "bfloat16", "bfloat16")
        if arch == "InternLM3ForCausalLM":           # ``qkv_bias`` for q/k/v, ``bias`` for o_proj and the MLP
            kw["attention_bias"] = bool(text.get("qkv_bias", False))
            kw["attention_out_bias"] = kw["mlp_bias"] = bool(text.get("bias", False))
        if arch in ("Ernie4_5ForCausalLM", "Ernie4_5_MoeForCausalLM") and text.get("use_bias") is not None:
            kw["attention_bias"] = kw["attention_out_bias"] = kw["mlp_bias"] = bool(text["use_bias"])
        if arch == "Ernie4_5_MoeForCausalLM":
            E = text.get("moe_num_experts") or 0
            kw["num_local_experts"] = max(E) if isinstance(E, (list, tuple)) else E
            kw["num_experts_per_tok"] = text.get("moe_k", 2)
            kw["n_shared_experts"] = text.get("moe_num_shared_experts", 0)
            mi = text.get("moe_intermediate_size")
            kw["moe_intermediate_size"] = mi[0] if isinstance(mi, (list, tuple)) else mi
            L = text["num_hidden_layers"]
            lo, every = text.get("moe_layer_start_index", 0), max(1, text.get("moe_layer_interval", 1))
            hi = text.get("moe_layer_end_index", -1)
            hi = L - 1 if hi is None or hi < 0 else hi
            lo = min(lo) if isinstance(lo, (list, tuple)) else lo
            kw["mlp_only_layers"] = [i for i in range(L) if not (lo <= i <= hi and (i + 1) % every == 0)]
        if arch in ("HunYuanDenseV1ForCausalLM", "HunYuanMoEV1ForCausalLM"):
            rp_ = text.get("rope_parameters") or text.get("rope_scaling") or {}
            hd_ = kw.get("head_dim") or kw["hidden_size"] // kw["num_attention_heads"]
            if rp_.get("alpha") and (rp_.get("rope_type") or rp_.get("type")) == "dynamic":       # NTK-alpha: a larger base, no position-dependent part
                kw["rope_theta"] = float(rp_.get("rope_theta", kw.get("rope_theta", 10000.0))) * float(rp_["alpha"]) ** (hd_ / (hd_ - 2))
                kw["rope_scaling"] = None
            first = lambda v_: v_[0] if isinstance(v_, (list, tuple)) else v_
            if arch.startswith("HunYuanMoE"):
                kw["num_local_experts"] = first(text.get("num_local_experts", text.get("num_experts", 0)))          # (key already renamed by _HF_KEY_MAP)
                kw["num_experts_per_tok"] = first(text.get("moe_topk", 1))
                if text.get("moe_intermediate_size"):
                    kw["moe_intermediate_size"] = first(text["moe_intermediate_size"])
        if arch in ("NemotronForCausalLM",) and text.get("norm_eps") is not None:
            kw["rms_norm_eps"] = text["norm_eps"]
        if arch == "GPTNeoForCausalLM":              # num_layers / num_heads / window_size / attention_layers; MLP width defaults to 4 x hidden
            kw["num_hidden_layers"], kw["num_attention_heads"] = text["num_layers"], text["num_heads"]
            kw["num_key_value_heads"] = text["num_heads"]
            kw["intermediate_size"] = text.get("intermediate_size") or 4 * text["hidden_size"]
            kw["sliding_window"] = text.get("window_size", 256)
            al = text.get("attention_layers") or [t for types, n in text.get("attention_types", [[["global", "local"], text["num_layers"] // 2]]) for _ in range(n) for t in types]
            kw["layer_types"] = ["sliding_attention" if t == "local" else "full_attention" for t in al]
        if arch in ("GPTJForCausalLM", "CodeGenForCausalLM", "MiniMaxM2ForCausalLM") and text.get("rotary_dim"):
            kw["partial_rotary_factor"] = text["rotary_dim"] / (kw.get("head_dim") or kw["hidden_size"] // kw["num_attention_heads"])
        if arch == "GPTBigCodeForCausalLM":
            kw["num_key_value_heads"] = 1 if text.get("multi_query", True) else kw["num_attention_heads"]
        if text.get("use_sliding_window") is False:       # Qwen2 / SmolLM3 ship a window size next to a switch that disables it
            kw["sliding_window"] = None
        if text.get("hidden_activation"):          # Gemma-2 spells it differently
            kw["hidden_act"] = text["hidden_activation"]
        kw["name"] = os.path.basename(model_path.rstrip("/"))
        if impl == "transformers" or (impl == "auto" and not model_path.endswith(".gguf") and not _has_native_class(kw["architecture"])):
            try:
                kw = _transformers_fallback_config(model_path, arch, kw)
            except Exception as e:          # not a Hugging Face model directory either: the registry reports the unknown architecture later
                if impl == "transformers":
                    raise
                logger.debug("no Transformers fallback for %s: %s", arch, e)
        kw.update(overrides)
    cfg = ModelConfig(**kw)
    if context_length is not None:
        cfg.max_position_embeddings = min(cfg.max_position_embeddings, context_length) if context_length > 0 else cfg.max_position_embeddings
    return cfg
