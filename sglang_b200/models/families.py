"""More decoder families on top of the shared Llama-style implementation (``models/llama.py``):

* ``Phi3ForCausalLM``       checkpoint stores fused ``qkv_proj`` / ``gate_up_proj`` -> split + TP-shard on load
* ``InternLM2ForCausalLM``  ``wqkv`` interleaved per KV group, ``w1/w3/w2`` MLP names, ``tok_embeddings`` / ``output``
* ``Qwen2MoeForCausalLM`` / ``Qwen3MoeForCausalLM``  sparse MoE blocks (+ sigmoid-gated shared expert for Qwen2-MoE), dense layers
  selected by ``mlp_only_layers`` / ``decoder_sparse_step``
* ``Gemma2ForCausalLM``     sandwich norms (pre/post feed-forward), alternating local/global attention, attention and final-logit
  soft-capping, ``query_pre_attn_scalar`` scaling
* ``GraniteForCausalLM``    Llama with embedding / residual / attention / logits multipliers
* ``Olmo2ForCausalLM``      post-norm placement (norm after attention / MLP) + q/k RMSNorm over the full projection

Parity: reference ``python/sglang/srt/models/{phi3,internlm2,qwen2_moe,qwen3_moe,gemma2,granite,olmo2}.py``.
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Tuple

import torch
from torch import nn

from .. import ops
from ..configs.model_config import ModelConfig
from ..layers.common import RadixAttention, RMSNorm
from ..layers.linear import MergedColumnParallelLinear, ReplicatedLinear, RowParallelLinear
from ..layers.moe import FusedMoE
from ..model_executor.forward_batch import ForwardBatch
from ..parallel.parallel_state import get_tp_group
from .llama import LlamaAttention, LlamaDecoderLayer, LlamaForCausalLM, LlamaMLP, _layer_id
from .mixtral import _EP_SIZE


# ---------------------------------------------------------------- Phi-3
class Phi3ForCausalLM(LlamaForCausalLM):
    def load_weights(self, weights: Iterable[Tuple[str, torch.Tensor]]) -> List[str]:
        cfg = self.cfg
        q = cfg.num_attention_heads * cfg.head_dim
        kv = cfg.num_key_value_heads * cfg.head_dim

        def split() -> Iterable[Tuple[str, torch.Tensor]]:
            for name, w in weights:
                if name.endswith("self_attn.qkv_proj.weight"):
                    base = name[: -len("qkv_proj.weight")]
                    yield base + "q_proj.weight", w[:q]
                    yield base + "k_proj.weight", w[q:q + kv]
                    yield base + "v_proj.weight", w[q + kv:]
                elif name.endswith("mlp.gate_up_proj.weight"):
                    base = name[: -len("gate_up_proj.weight")]
                    half = w.shape[0] // 2
                    yield base + "gate_proj.weight", w[:half]
                    yield base + "up_proj.weight", w[half:]
                else:
                    yield name, w

        return super().load_weights(split())


# ---------------------------------------------------------------- InternLM2
class InternLM2ForCausalLM(LlamaForCausalLM):
    _RENAME = [("model.tok_embeddings.", "model.embed_tokens."), (".attention.wo.", ".self_attn.o_proj."),
               (".feed_forward.w1.", ".mlp.gate_proj."), (".feed_forward.w3.", ".mlp.up_proj."), (".feed_forward.w2.", ".mlp.down_proj."),
               (".attention_norm.", ".input_layernorm."), (".ffn_norm.", ".post_attention_layernorm."), ("output.weight", "lm_head.weight")]

    def load_weights(self, weights: Iterable[Tuple[str, torch.Tensor]]) -> List[str]:
        cfg = self.cfg
        G = cfg.num_attention_heads // cfg.num_key_value_heads
        D = cfg.head_dim

        def remap() -> Iterable[Tuple[str, torch.Tensor]]:
            for name, w in weights:
                if ".attention.wqkv." in name:
                    base = name.replace(".attention.wqkv.", ".self_attn.{}.")
                    w4 = w.view(cfg.num_key_value_heads, G + 2, D, w.shape[-1])       # per KV group: G query heads, then k, then v
                    yield base.format("q_proj"), w4[:, :G].reshape(-1, w.shape[-1])
                    yield base.format("k_proj"), w4[:, G].reshape(-1, w.shape[-1])
                    yield base.format("v_proj"), w4[:, G + 1].reshape(-1, w.shape[-1])
                    continue
                for a, b in self._RENAME:
                    if a in name:
                        name = name.replace(a, b)
                yield name, w

        return super().load_weights(remap())


# ---------------------------------------------------------------- Qwen2-MoE / Qwen3-MoE
class QwenSparseMoeBlock(nn.Module):
    def __init__(self, cfg: ModelConfig, layer_id: int, dtype, quant, prefix: str):
        super().__init__()
        self.gate = ReplicatedLinear(cfg.hidden_size, cfg.num_local_experts, False, dtype, None, prefix + ".gate")
        self.experts = FusedMoE(cfg.num_local_experts, cfg.num_experts_per_tok, cfg.hidden_size, cfg.moe_intermediate_size, dtype,
                                cfg.norm_topk_prob, _EP_SIZE["v"], quant, layer_id)
        self.shared_expert = None
        if cfg.shared_expert_intermediate_size:
            self.shared_expert = _SharedExpert(cfg, dtype, quant, prefix + ".shared_expert")
            self.shared_expert_gate = ReplicatedLinear(cfg.hidden_size, 1, False, dtype, None, prefix + ".shared_expert_gate")

    def forward(self, x: torch.Tensor):
        y = self.experts(x, self.gate(x))
        if self.shared_expert is None:
            return y
        if hasattr(y, "resolve"):          # the shared-expert sum needs the reduced tensor
            y = y.resolve()
        s = self.shared_expert(x)
        return y + torch.sigmoid(self.shared_expert_gate(x).float()).to(s.dtype) * s


class _SharedExpert(nn.Module):
    def __init__(self, cfg: ModelConfig, dtype, quant, prefix: str):
        super().__init__()
        I = cfg.shared_expert_intermediate_size
        self.gate_up_proj = MergedColumnParallelLinear(cfg.hidden_size, [I, I], False, dtype, quant, prefix + ".gate_up_proj")
        self.down_proj = RowParallelLinear(I, cfg.hidden_size, False, True, dtype, quant, prefix + ".down_proj")

    def forward(self, x):
        return self.down_proj(ops.silu_and_mul(self.gate_up_proj(x)))


class Qwen2MoeForCausalLM(LlamaForCausalLM):
    def _mlp_cls(self):
        cfg = self.cfg

        def make(c, layer_id_or_dtype, *rest):
            # LlamaDecoderLayer calls mlp_cls(cfg, layer_id, dtype, quant, prefix) for non-LlamaMLP classes
            layer_id, dtype, quant, prefix = layer_id_or_dtype, rest[0], rest[1], rest[2]
            gid = layer_id + self.model_start_layer()
            sparse = (gid not in (cfg.mlp_only_layers or [])) and cfg.num_local_experts > 0 and (gid + 1) % max(1, cfg.decoder_sparse_step) == 0
            return QwenSparseMoeBlock(c, layer_id, dtype, quant, prefix + ".mlp") if sparse else LlamaMLP(c, dtype, quant, prefix + ".mlp")

        return make

    def model_start_layer(self) -> int:
        from ..parallel.parallel_state import get_pp_group
        from .llama import pp_layer_range

        pp = get_pp_group()
        return pp_layer_range(self.cfg.num_hidden_layers, pp.rank, pp.world_size)[0]

    def load_weights(self, weights: Iterable[Tuple[str, torch.Tensor]]) -> List[str]:
        rest, loaded = [], []
        modules = dict(self.named_modules())
        shard_of = {"gate_proj": "w1", "up_proj": "w3", "down_proj": "w2"}
        for name, w in weights:
            if ".mlp.experts." in name:
                lid = _layer_id(name)
                if not (self.model.start_layer <= lid < self.model.end_layer):
                    continue
                local = lid - self.model.start_layer
                eid = int(name.split(".experts.")[1].split(".")[0])
                shard = shard_of[name.split(".")[-2]]
                mod = modules.get(f"model.layers.{local}.mlp.experts")
                if mod is None:
                    continue
                mod.weight_loader(mod.w2 if shard == "w2" else mod.w13, w, eid, shard)
                loaded.append(name)
            else:
                rest.append((name, w))
        return loaded + super().load_weights(rest)


class Qwen3MoeForCausalLM(Qwen2MoeForCausalLM):
    pass


# ---------------------------------------------------------------- Gemma-2
class Gemma2DecoderLayer(LlamaDecoderLayer):
    def __init__(self, cfg: ModelConfig, layer_id: int, dtype, quant, rope, prefix: str, mlp_cls=LlamaMLP):
        super().__init__(cfg, layer_id, dtype, quant, rope, prefix, LlamaMLP)
        self.pre_feedforward_layernorm = RMSNorm(cfg.hidden_size, cfg.rms_norm_eps, dtype, True)
        self.post_feedforward_layernorm = RMSNorm(cfg.hidden_size, cfg.rms_norm_eps, dtype, True)
        gid = int(prefix.rsplit(".", 1)[1])
        a = self.self_attn.attn
        local = (gid % 2 == 0) if not cfg.layer_types else cfg.layer_types[gid] == "sliding_attention"
        a.window = cfg.sliding_window if (local and cfg.sliding_window) else -1
        if cfg.query_pre_attn_scalar:
            a.scale = float(cfg.query_pre_attn_scalar) ** -0.5

    def forward(self, h, residual, fb: ForwardBatch):
        # gemma-2 residual stream: x + post_attn_norm(attn(input_norm(x))) ; then x + post_ffn_norm(mlp(pre_ffn_norm(x)))
        if residual is not None:
            h = h + residual
        x = h
        a = self.self_attn(self.input_layernorm(x), fb)
        a = a.resolve() if hasattr(a, "resolve") else a
        x = x + self.post_attention_layernorm(a)
        m = self.mlp(self.pre_feedforward_layernorm(x))
        m = m.resolve() if hasattr(m, "resolve") else m
        x = x + self.post_feedforward_layernorm(m)
        return x, None


class Gemma2ForCausalLM(LlamaForCausalLM):
    decoder_layer_cls = Gemma2DecoderLayer


class VaultGemmaDecoderLayer(Gemma2DecoderLayer):
    """VaultGemma (differentially private Gemma): the Gemma-2 block without the two post-norms"""

    def __init__(self, cfg: ModelConfig, layer_id: int, dtype, quant, rope, prefix: str, mlp_cls=LlamaMLP):
        super().__init__(cfg, layer_id, dtype, quant, rope, prefix, mlp_cls)
        del self.post_attention_layernorm
        del self.post_feedforward_layernorm

    def forward(self, h, residual, fb: ForwardBatch):
        if residual is not None:
            h = h + residual
        a = self.self_attn(self.input_layernorm(h), fb)
        h = h + (a.resolve() if hasattr(a, "resolve") else a)
        m = self.mlp(self.pre_feedforward_layernorm(h))
        return h + (m.resolve() if hasattr(m, "resolve") else m), None


class VaultGemmaForCausalLM(LlamaForCausalLM):
    decoder_layer_cls = VaultGemmaDecoderLayer


# ---------------------------------------------------------------- Granite (scalar multipliers)
class GraniteForCausalLM(LlamaForCausalLM):
    def __init__(self, cfg: ModelConfig, quant_name: Optional[str] = None, device="cpu"):
        super().__init__(cfg, quant_name, device)
        for layer in self.model.layers:
            if cfg.attention_multiplier:
                layer.self_attn.attn.scale = float(cfg.attention_multiplier)
            layer.residual_multiplier = float(cfg.residual_multiplier or 1.0)
        if (cfg.residual_multiplier or 1.0) != 1.0:
            for layer in self.model.layers:
                layer.forward = _granite_layer_forward.__get__(layer)

    def forward(self, fb: ForwardBatch, inputs_embeds=None):
        if inputs_embeds is None and self.model.pp.is_first and (self.cfg.embedding_multiplier or 1.0) != 1.0:
            inputs_embeds = self.model.embed(fb) * float(self.cfg.embedding_multiplier)
        return self.model(fb, inputs_embeds)

    def compute_logits(self, hidden: torch.Tensor, key: str = "logits") -> torch.Tensor:
        y = super().compute_logits(hidden, key)
        return y / float(self.cfg.logits_scaling) if (self.cfg.logits_scaling or 1.0) != 1.0 else y


def _granite_layer_forward(self, h, residual, fb: ForwardBatch):
    if residual is not None:
        h = h + residual
    x = h
    a = self.self_attn(self.input_layernorm(x), fb)
    a = a.resolve() if hasattr(a, "resolve") else a
    x = x + a * self.residual_multiplier
    m = self.mlp(self.post_attention_layernorm(x))
    m = m.resolve() if hasattr(m, "resolve") else m
    return x + m * self.residual_multiplier, None


# ---------------------------------------------------------------- OLMo-2 (post-norm + full-width q/k norm)
class Olmo2Attention(LlamaAttention):
    def __init__(self, cfg: ModelConfig, layer_id: int, dtype, quant, rope, prefix: str):
        super().__init__(cfg, layer_id, dtype, quant, rope, prefix)
        tp = get_tp_group().world_size
        self.q_norm = RMSNorm(cfg.num_attention_heads * cfg.head_dim // tp, cfg.rms_norm_eps, dtype)
        self.k_norm = RMSNorm(max(1, cfg.num_key_value_heads // tp) * cfg.head_dim, cfg.rms_norm_eps, dtype)
        self.qk_norm = False        # handled here (over the whole projection, not per head)

    def forward(self, x: torch.Tensor, fb: ForwardBatch):
        qkv = self.qkv_proj(x)
        q, k, v = qkv.split([self.q_size, self.kv_size, self.kv_size], dim=-1)
        q = self.q_norm(q.contiguous())
        k = self.k_norm(k.contiguous())
        o = self.attn(q, k, v.contiguous(), fb)
        return self.o_proj(o)


class Olmo2DecoderLayer(LlamaDecoderLayer):
    def __init__(self, cfg: ModelConfig, layer_id: int, dtype, quant, rope, prefix: str, mlp_cls=LlamaMLP):
        super().__init__(cfg, layer_id, dtype, quant, rope, prefix, LlamaMLP)
        self.self_attn = Olmo2Attention(cfg, layer_id, dtype, quant, rope, prefix + ".self_attn")
        del self.input_layernorm
        self.post_feedforward_layernorm = RMSNorm(cfg.hidden_size, cfg.rms_norm_eps, dtype)

    def forward(self, h, residual, fb: ForwardBatch):
        if residual is not None:
            h = h + residual
        a = self.self_attn(h, fb)
        h = h + self.post_attention_layernorm(a.resolve() if hasattr(a, "resolve") else a)
        m = self.mlp(h)
        h = h + self.post_feedforward_layernorm(m.resolve() if hasattr(m, "resolve") else m)
        return h, None


class Olmo2ForCausalLM(LlamaForCausalLM):
    decoder_layer_cls = Olmo2DecoderLayer


# ---------------------------------------------------------------- EXAONE-4: post-norm blocks (like OLMo-2) around the ordinary attention with per-head q/k norm
class Exaone4DecoderLayer(LlamaDecoderLayer):
    def __init__(self, cfg: ModelConfig, layer_id: int, dtype, quant, rope, prefix: str, mlp_cls=LlamaMLP):
        super().__init__(cfg, layer_id, dtype, quant, rope, prefix, LlamaMLP)      # LlamaAttention: qk_norm per head, layer_types, NoPE layers
        del self.input_layernorm
        self.post_feedforward_layernorm = RMSNorm(cfg.hidden_size, cfg.rms_norm_eps, dtype)

    def forward(self, h, residual, fb: ForwardBatch):
        if residual is not None:
            h = h + residual
        a = self.self_attn(h, fb)
        h = h + self.post_attention_layernorm(a.resolve() if hasattr(a, "resolve") else a)
        m = self.mlp(h)
        h = h + self.post_feedforward_layernorm(m.resolve() if hasattr(m, "resolve") else m)
        return h, None


class Exaone4ForCausalLM(LlamaForCausalLM):
    """LG EXAONE-4.0 (reference models/exaone4.py): 1.2B is all-global with rotary everywhere; 32B interleaves sliding-window layers (rotary) with
    global NoPE layers"""
    decoder_layer_cls = Exaone4DecoderLayer
