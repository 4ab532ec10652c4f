"""More decoder families on the shared Llama skeleton: GLM / GLM-4, Gemma-3 (text), Cohere (Command-R), OLMoE, Nemotron (dense).

Parity: reference ``python/sglang/srt/models/{glm4,gemma3_causal,commandr,olmoe,nemotron}.py``.  Every class here is checked against the Hugging Face
implementation in ``tests/test_hf_parity.py`` (tiny random model -> ``save_pretrained`` -> served from disk).
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Tuple

import torch
import torch.nn.functional as F
from torch import nn

from ..configs.model_config import ModelConfig
from ..layers.common import RMSNorm, RotaryEmbedding
from ..layers.linear import ReplicatedLinear
from ..layers.moe import FusedMoE
from ..model_executor.forward_batch import ForwardBatch
from ..parallel.parallel_state import get_tp_group
from .llama import LlamaAttention, LlamaDecoderLayer, LlamaForCausalLM, LlamaMLP, _layer_id


def _resolve(y):
    return y.resolve() if hasattr(y, "resolve") else y


def _gid(prefix: str) -> int:
    return int(prefix.rsplit(".", 1)[1])


# ---------------------------------------------------------------- GLM / GLM-4 (fused gate_up in the checkpoint, sandwich norms in GLM-4)
class GlmForCausalLM(LlamaForCausalLM):
    def load_weights(self, weights: Iterable[Tuple[str, torch.Tensor]]) -> List[str]:
        def split():
            for name, w in weights:
                if ".mlp.gate_up_proj." in name:        # [gate; up] stacked on the output dim
                    g, u = w.chunk(2, dim=0)
                    yield name.replace("gate_up_proj", "gate_proj"), g
                    yield name.replace("gate_up_proj", "up_proj"), u
                else:
                    yield name, w
        return super().load_weights(split())


class Glm4DecoderLayer(LlamaDecoderLayer):
    def __init__(self, cfg: ModelConfig, layer_id: int, dtype, quant, rope, prefix: str, mlp_cls=LlamaMLP):
        super().__init__(cfg, layer_id, dtype, quant, rope, prefix, LlamaMLP)
        self.post_self_attn_layernorm = RMSNorm(cfg.hidden_size, cfg.rms_norm_eps, dtype)
        self.post_mlp_layernorm = RMSNorm(cfg.hidden_size, cfg.rms_norm_eps, dtype)

    def forward(self, h, residual, fb: ForwardBatch):
        if residual is not None:
            h = h + residual
        a = _resolve(self.self_attn(self.input_layernorm(h), fb))
        h = h + self.post_self_attn_layernorm(a)
        m = _resolve(self.mlp(self.post_attention_layernorm(h)))
        return h + self.post_mlp_layernorm(m), None


class Glm4ForCausalLM(GlmForCausalLM):
    decoder_layer_cls = Glm4DecoderLayer


# ---------------------------------------------------------------- Gemma-3 text (per-head q/k norm, local/global rope bases, sandwich norms)
class Gemma3DecoderLayer(LlamaDecoderLayer):
    def __init__(self, cfg: ModelConfig, layer_id: int, dtype, quant, rope, prefix: str, mlp_cls=LlamaMLP):
        super().__init__(cfg, layer_id, dtype, quant, rope, prefix, LlamaMLP)
        self.pre_feedforward_layernorm = RMSNorm(cfg.hidden_size, cfg.rms_norm_eps, dtype, True)
        self.post_feedforward_layernorm = RMSNorm(cfg.hidden_size, cfg.rms_norm_eps, dtype, True)
        sa = self.self_attn
        sa.qk_norm = True
        sa.q_norm = RMSNorm(cfg.head_dim, cfg.rms_norm_eps, dtype, True)
        sa.k_norm = RMSNorm(cfg.head_dim, cfg.rms_norm_eps, dtype, True)
        gid = _gid(prefix)
        kinds = cfg.layer_types or ["full_attention" if (i + 1) % 6 == 0 else "sliding_attention" for i in range(cfg.num_hidden_layers)]
        local = kinds[gid] == "sliding_attention"
        sa.attn.window = cfg.sliding_window if (local and cfg.sliding_window) else -1
        if local and cfg.rope_local_base_freq:
            sa.attn.rope = RotaryEmbedding(cfg.head_dim, cfg.max_position_embeddings, cfg.rope_local_base_freq, None, cfg.rope_is_neox, 1.0,
                                           rope.cos_sin.device)
        if cfg.query_pre_attn_scalar:
            sa.attn.scale = float(cfg.query_pre_attn_scalar) ** -0.5

    def forward(self, h, residual, fb: ForwardBatch):
        if residual is not None:
            h = h + residual
        a = _resolve(self.self_attn(self.input_layernorm(h), fb))
        h = h + self.post_attention_layernorm(a)
        m = _resolve(self.mlp(self.pre_feedforward_layernorm(h)))
        return h + self.post_feedforward_layernorm(m), None


class Gemma3ForCausalLM(LlamaForCausalLM):
    decoder_layer_cls = Gemma3DecoderLayer

    def load_weights(self, weights):
        return super().load_weights((n[len("language_model."):] if n.startswith("language_model.") else n, w) for n, w in weights)


# ---------------------------------------------------------------- Cohere / Command-R (mean-centred LayerNorm, parallel block, logit scale)
class CohereNorm(nn.Module):
    def __init__(self, shape, eps: float, dtype):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(shape, dtype=dtype), requires_grad=False)
        self.eps = eps

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        f = x.float()
        f = f - f.mean(-1, keepdim=True)
        return (f * torch.rsqrt(f.pow(2).mean(-1, keepdim=True) + self.eps) * self.weight.float()).to(x.dtype)


class CohereAttention(LlamaAttention):
    def __init__(self, cfg: ModelConfig, layer_id: int, dtype, quant, rope, prefix: str):
        super().__init__(cfg, layer_id, dtype, quant, rope, prefix)
        self.use_qk = bool(cfg.use_qk_norm)
        if self.use_qk:
            self.q_norm = CohereNorm((self.attn.num_heads, cfg.head_dim), cfg.rms_norm_eps, dtype)
            self.k_norm = CohereNorm((self.attn.num_kv_heads, cfg.head_dim), cfg.rms_norm_eps, dtype)
        self.qk_norm = False

    def forward(self, x: torch.Tensor, fb: ForwardBatch):
        if not self.use_qk:
            return super().forward(x, fb)
        T = x.shape[0]
        q, k, v = self.qkv_proj(x).split([self.q_size, self.kv_size, self.kv_size], dim=-1)
        q = self.q_norm(q.reshape(T, self.attn.num_heads, self.head_dim)).reshape(T, self.q_size)
        k = self.k_norm(k.reshape(T, self.attn.num_kv_heads, self.head_dim)).reshape(T, self.kv_size)
        return self.o_proj(self.attn(q, k, v.contiguous(), fb))


class CohereDecoderLayer(LlamaDecoderLayer):
    def __init__(self, cfg: ModelConfig, layer_id: int, dtype, quant, rope, prefix: str, mlp_cls=LlamaMLP):
        super().__init__(cfg, layer_id, dtype, quant, rope, prefix, LlamaMLP)
        self.self_attn = CohereAttention(cfg, layer_id, dtype, quant, rope, prefix + ".self_attn")
        self.input_layernorm = CohereNorm(cfg.hidden_size, cfg.rms_norm_eps, dtype)
        del self.post_attention_layernorm

    def forward(self, h, residual, fb: ForwardBatch):
        if residual is not None:
            h = h + residual
        x = self.input_layernorm(h)
        return h + _resolve(self.self_attn(x, fb)) + _resolve(self.mlp(x)), None


class CohereForCausalLM(LlamaForCausalLM):
    decoder_layer_cls = CohereDecoderLayer

    def __init__(self, cfg: ModelConfig, quant_name: Optional[str] = None, device="cpu"):
        super().__init__(cfg, quant_name, device)
        self.model.norm = CohereNorm(cfg.hidden_size, cfg.rms_norm_eps, self.model.embed_tokens.weight.dtype)

    def compute_logits(self, hidden: torch.Tensor, key: str = "logits") -> torch.Tensor:
        y = super().compute_logits(hidden, key)
        return y * float(self.cfg.logit_scale) if self.cfg.logit_scale else y

    def load_weights(self, weights):
        tp = get_tp_group()

        def fix():
            for n, w in weights:
                if n.endswith(("q_norm.weight", "k_norm.weight")) and w.dim() == 2 and tp.world_size > 1:       # [heads, head_dim]: shard heads
                    per = max(1, w.shape[0] // tp.world_size)
                    w = w[(tp.rank * per) % w.shape[0]:(tp.rank * per) % w.shape[0] + per]
                yield n, w
        return super().load_weights(fix())


# ---------------------------------------------------------------- OLMoE (full-width q/k RMSNorm + softmax top-k experts)
class OlmoeMoE(nn.Module):
    def __init__(self, cfg: ModelConfig, layer_id: int, dtype, quant, prefix: str):
        super().__init__()
        from .mixtral import _EP_SIZE

        self.gate = ReplicatedLinear(cfg.hidden_size, cfg.num_local_experts, False, dtype, None, prefix + ".gate")
        self.experts = FusedMoE(cfg.num_local_experts, cfg.num_experts_per_tok, cfg.hidden_size, cfg.moe_intermediate_size, dtype,
                                cfg.norm_topk_prob, _EP_SIZE["v"], quant, layer_id)

    def forward(self, x):
        return self.experts(x, self.gate(x))


class OlmoeDecoderLayer(LlamaDecoderLayer):
    def __init__(self, cfg: ModelConfig, layer_id: int, dtype, quant, rope, prefix: str, mlp_cls=LlamaMLP):
        from .families import Olmo2Attention

        super().__init__(cfg, layer_id, dtype, quant, rope, prefix, OlmoeMoE)
        self.self_attn = Olmo2Attention(cfg, layer_id, dtype, quant, rope, prefix + ".self_attn")


class OlmoeForCausalLM(LlamaForCausalLM):
    decoder_layer_cls = OlmoeDecoderLayer

    def _mlp_cls(self):
        return OlmoeMoE

    def load_weights(self, weights):
        mods = dict(self.named_modules())
        rest, loaded = [], []
        for name, w in weights:
            if ".mlp.experts." in name:
                lid = _layer_id(name)
                if not (self.model.start_layer <= lid < self.model.end_layer):
                    continue
                mod = mods[f"model.layers.{lid - self.model.start_layer}.mlp.experts"]
                tail = name.split(".experts.")[1]
                if tail.split(".")[0].isdigit():                      # per-expert tensors
                    eid, proj = int(tail.split(".")[0]), tail.split(".")[1]
                    shard = {"gate_proj": "w1", "up_proj": "w3", "down_proj": "w2"}[proj]
                    mod.weight_loader(mod.w2 if shard == "w2" else mod.w13, w, eid, shard)
                elif tail.startswith("gate_up_proj"):                 # stacked [E, 2I, H]
                    for e in range(w.shape[0]):
                        g, u = w[e].chunk(2, dim=0)
                        mod.weight_loader(mod.w13, g, e, "w1")
                        mod.weight_loader(mod.w13, u, e, "w3")
                elif tail.startswith("down_proj"):                    # stacked [E, H, I]
                    for e in range(w.shape[0]):
                        mod.weight_loader(mod.w2, w[e], e, "w2")
                loaded.append(name)
            else:
                rest.append((name, w))
        return loaded + super().load_weights(rest)


# ---------------------------------------------------------------- Nemotron (dense): LayerNorm1P, partial rotary, squared-ReLU MLP
class LayerNorm1P(nn.Module):
    def __init__(self, n: int, eps: float, dtype):
        super().__init__()
        self.weight = nn.Parameter(torch.zeros(n, dtype=dtype), requires_grad=False)
        self.bias = nn.Parameter(torch.zeros(n, dtype=dtype), requires_grad=False)
        self.eps = eps

    def forward(self, x, residual=None):
        if residual is not None:
            x = x + residual
            residual = x
        y = F.layer_norm(x.float(), (x.shape[-1],), self.weight.float() + 1.0, self.bias.float(), self.eps).to(x.dtype)
        return y if residual is None else (y, residual)


class NemotronDecoderLayer(LlamaDecoderLayer):
    def __init__(self, cfg: ModelConfig, layer_id: int, dtype, quant, rope, prefix: str, mlp_cls=LlamaMLP):
        super().__init__(cfg, layer_id, dtype, quant, rope, prefix, LlamaMLP)
        self.input_layernorm = LayerNorm1P(cfg.hidden_size, cfg.rms_norm_eps, dtype)
        self.post_attention_layernorm = LayerNorm1P(cfg.hidden_size, cfg.rms_norm_eps, dtype)

    def forward(self, h, residual, fb: ForwardBatch):
        if residual is not None:
            h = h + residual
        h = h + _resolve(self.self_attn(self.input_layernorm(h), fb))
        return h + _resolve(self.mlp(self.post_attention_layernorm(h))), None


class NemotronForCausalLM(LlamaForCausalLM):
    decoder_layer_cls = NemotronDecoderLayer

    def __init__(self, cfg: ModelConfig, quant_name: Optional[str] = None, device="cpu"):
        super().__init__(cfg, quant_name, device)
        self.model.norm = LayerNorm1P(cfg.hidden_size, cfg.rms_norm_eps, self.model.embed_tokens.weight.dtype)


# ---------------------------------------------------------------- OLMo (v1) and StableLM: the Llama stack with LayerNorm
class PlainLayerNorm(nn.Module):
    """LayerNorm with optional affine parameters; OLMo-1 uses none at all (``elementwise_affine=False``), StableLM weight + bias"""

    def __init__(self, n: int, eps: float, dtype, affine: bool):
        super().__init__()
        self.eps, self.n = eps, n
        if affine:
            self.weight = nn.Parameter(torch.ones(n, dtype=dtype), requires_grad=False)
            self.bias = nn.Parameter(torch.zeros(n, dtype=dtype), requires_grad=False)
        else:
            self.weight = self.bias = None

    def forward(self, x, residual=None):
        if residual is not None:
            x = x + residual
            residual = x
        y = F.layer_norm(x.float(), (self.n,), None if self.weight is None else self.weight.float(), None if self.bias is None else self.bias.float(),
                         self.eps).to(x.dtype)
        return y if residual is None else (y, residual)


class _ClipAttention(LlamaAttention):
    """OLMo ``clip_qkv``: the fused projection output is clamped before RoPE"""

    def forward(self, x: torch.Tensor, fb: ForwardBatch):
        qkv = self.qkv_proj(x)
        if self.clip is not None:
            qkv = qkv.clamp(-self.clip, self.clip)
        q, k, v = qkv.split([self.q_size, self.kv_size, self.kv_size], dim=-1)
        return self.o_proj(self.attn(q.contiguous(), k.contiguous(), v.contiguous(), fb))


class _LNDecoderLayer(LlamaDecoderLayer):
    affine = True

    def __init__(self, cfg: ModelConfig, layer_id: int, dtype, quant, rope, prefix: str, mlp_cls=LlamaMLP):
        super().__init__(cfg, layer_id, dtype, quant, rope, prefix, LlamaMLP)
        self.input_layernorm = PlainLayerNorm(cfg.hidden_size, cfg.rms_norm_eps, dtype, self.affine)
        self.post_attention_layernorm = PlainLayerNorm(cfg.hidden_size, cfg.rms_norm_eps, dtype, self.affine)
        if cfg.clip_qkv:
            self.self_attn = _ClipAttention(cfg, layer_id, dtype, quant, rope, prefix + ".self_attn")
            self.self_attn.clip = float(cfg.clip_qkv)

    def forward(self, h, residual, fb: ForwardBatch):
        if residual is not None:
            h = h + residual
        h = h + _resolve(self.self_attn(self.input_layernorm(h), fb))
        return h + _resolve(self.mlp(self.post_attention_layernorm(h))), None


class _OlmoLayer(_LNDecoderLayer):
    affine = False


class OlmoForCausalLM(LlamaForCausalLM):
    """OLMo-1 (allenai/OLMo-*-hf): non-parametric LayerNorm, optional qkv clipping, otherwise Llama (reference models/olmo.py)"""
    decoder_layer_cls = _OlmoLayer

    def __init__(self, cfg: ModelConfig, quant_name: Optional[str] = None, device="cpu"):
        super().__init__(cfg, quant_name, device)
        self.model.norm = PlainLayerNorm(cfg.hidden_size, cfg.rms_norm_eps, self.model.embed_tokens.weight.dtype, False)


class StableLmForCausalLM(LlamaForCausalLM):
    """StableLM-2 / StableLM-3B-4E1T: LayerNorm with bias, partial rotary (25 %), optional qkv bias, gated SiLU MLP (reference models/stablelm.py)"""
    decoder_layer_cls = _LNDecoderLayer

    def __init__(self, cfg: ModelConfig, quant_name: Optional[str] = None, device="cpu"):
        super().__init__(cfg, quant_name, device)
        self.model.norm = PlainLayerNorm(cfg.hidden_size, cfg.rms_norm_eps, self.model.embed_tokens.weight.dtype, True)


# ---------------------------------------------------------------- Apertus: non-gated MLP with the learnable xIELU activation, per-head q/k norm
class XIELU(nn.Module):
    """xIELU (arXiv 2411.13010): x > 0: a_p x^2 + b x ; x <= 0: a_n (expm1(min(x, eps)) - x) + b x, with a_p = softplus(alpha_p),
    a_n = b + softplus(alpha_n) learned per layer"""

    def __init__(self, dtype, beta: float = 0.5, eps: float = -1e-6):
        super().__init__()
        self.alpha_p = nn.Parameter(torch.zeros(1, dtype=dtype), requires_grad=False)
        self.alpha_n = nn.Parameter(torch.zeros(1, dtype=dtype), requires_grad=False)
        self.beta, self.eps = beta, eps

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        xf = x.float()
        a_p = F.softplus(self.alpha_p.float())
        a_n = self.beta + F.softplus(self.alpha_n.float())
        y = torch.where(xf > 0, a_p * xf * xf + self.beta * xf, (torch.expm1(torch.clamp(xf, max=self.eps)) - xf) * a_n + self.beta * xf)
        return y.to(x.dtype)


class ApertusDecoderLayer(LlamaDecoderLayer):
    def __init__(self, cfg: ModelConfig, layer_id: int, dtype, quant, rope, prefix: str, mlp_cls=LlamaMLP):
        super().__init__(cfg, layer_id, dtype, quant, rope, prefix, LlamaMLP)      # cfg.mlp_gated = False -> up_proj / act / down_proj
        if cfg.hidden_act == "xielu":
            self.mlp.act = None
            self.mlp.act_fn = XIELU(dtype)
            self.mlp.forward = self._mlp_forward

    def _mlp_forward(self, x):
        m = self.mlp
        return m.down_proj(m.act_fn(m.up_proj(x)), defer_reduce=True)


class ApertusForCausalLM(LlamaForCausalLM):
    """Swiss-AI Apertus (reference models/apertus.py): Llama skeleton, q/k RMSNorm per head, xIELU MLP; HF names its norms
    attention_layernorm / feedforward_layernorm"""
    decoder_layer_cls = ApertusDecoderLayer

    def load_weights(self, weights):
        ren = ((".attention_layernorm.", ".input_layernorm."), (".feedforward_layernorm.", ".post_attention_layernorm."))

        def it():
            for n, w in weights:
                for a, b in ren:
                    n = n.replace(a, b)
                yield n, w

        return super().load_weights(it())


# ---------------------------------------------------------------- GLM-4.5 (Glm4Moe) and dots.llm1: DeepSeek-V3 style MoE + ordinary GQA attention
def _ds():
    from . import deepseek

    return deepseek


class _DsMoeLayer(LlamaDecoderLayer):
    """standard attention; dense MLP for the first ``first_k_dense_replace`` layers, DeepSeek MoE (sigmoid router + correction bias, grouped
    top-k, shared experts, routed scaling) afterwards"""

    def __init__(self, cfg: ModelConfig, layer_id: int, dtype, quant, rope, prefix: str, mlp_cls=LlamaMLP):
        super().__init__(cfg, layer_id, dtype, quant, rope, prefix, LlamaMLP)
        gid = _gid(prefix)
        if cfg.num_local_experts > 0 and gid >= (cfg.first_k_dense_replace or 0) and gid not in (cfg.mlp_only_layers or []):
            self.mlp = _ds().DeepseekV2MoE(cfg, layer_id, dtype, quant, prefix + ".mlp")
        if cfg.use_qk_norm:
            sa = self.self_attn
            sa.qk_norm = True
            sa.q_norm = RMSNorm(cfg.head_dim, cfg.rms_norm_eps, dtype)
            sa.k_norm = RMSNorm(cfg.head_dim, cfg.rms_norm_eps, dtype)


class Glm4MoeForCausalLM(LlamaForCausalLM):
    decoder_layer_cls = _DsMoeLayer

    def load_weights(self, weights):
        # expert / router tensors use the DeepSeek names: reuse that loader, everything else is Llama-shaped
        loaded, rest = _ds().load_moe_tensors(self, weights)
        return loaded + super().load_weights(rest)


class Dots1ForCausalLM(Glm4MoeForCausalLM):
    pass


class Ernie4_5_MoeForCausalLM(Glm4MoeForCausalLM):
    """ERNIE-4.5 MoE (reference ``python/sglang/srt/models/ernie4.py``): GPT-J style RoPE, soft-max router whose top-k choice adds the
    ``moe_statics`` correction bias while the combine weights stay the unbiased probabilities (renormalised), shared experts of width
    ``moe_intermediate_size * moe_num_shared_experts``; the sparse layers are ``start <= i <= end`` with ``(i + 1) % interval == 0`` (the config
    loader turns that into ``mlp_only_layers``)."""

    def load_weights(self, weights):
        def rename():
            for name, w in weights:
                if name.endswith("mlp.moe_statics.e_score_correction_bias"):      # [1, E] under the block, [E] under the gate here
                    yield name.replace("mlp.moe_statics.", "mlp.gate."), w.reshape(-1)
                elif ".mtp_" in name or name.startswith("model.mtp_"):             # multi-token-prediction heads of the release checkpoints
                    continue
                else:
                    yield name, w
        return super().load_weights(rename())


# ---------------------------------------------------------------- HunYuan (dense V1 / MoE V1): per-head q/k RMSNorm applied AFTER the rotary embedding
class HunYuanDecoderLayer(LlamaDecoderLayer):
    """reference ``python/sglang/srt/models/hunyuan.py``.  The norm sits between RoPE and the cache write, so the fused RoPE + KV-write kernel
    cannot be used: rotate, normalise, then hand rotated keys to the attention layer with its own rotary step switched off."""

    def __init__(self, cfg: ModelConfig, layer_id: int, dtype, quant, rope, prefix: str, mlp_cls=LlamaMLP):
        super().__init__(cfg, layer_id, dtype, quant, rope, prefix, mlp_cls)
        sa = self.self_attn
        sa.query_layernorm = RMSNorm(cfg.head_dim, cfg.rms_norm_eps, dtype)
        sa.key_layernorm = RMSNorm(cfg.head_dim, cfg.rms_norm_eps, dtype)
        sa._rope = sa.attn.rope
        sa.attn.rope = None
        sa.forward = _hunyuan_attn_forward.__get__(sa)


def _hunyuan_attn_forward(self, x: torch.Tensor, fb: ForwardBatch):
    from ..ops import torch_ref

    T = x.shape[0]
    q, k, v = self.qkv_proj(x).split([self.q_size, self.kv_size, self.kv_size], dim=-1)
    nh, nkv, hd = self.attn.num_heads, self.attn.num_kv_heads, self.head_dim
    cs = self._rope.cos_sin
    q = torch_ref.apply_rope(q.reshape(T, nh, hd), fb.positions, cs, True)
    k = torch_ref.apply_rope(k.reshape(T, nkv, hd), fb.positions, cs, True)
    q = self.query_layernorm(q.reshape(T * nh, hd).to(x.dtype)).view(T, nh * hd)
    k = self.key_layernorm(k.reshape(T * nkv, hd).to(x.dtype)).view(T, nkv * hd)
    return self.o_proj(self.attn(q.contiguous(), k.contiguous(), v.contiguous(), fb), defer_reduce=True)


class HunYuanDenseV1ForCausalLM(LlamaForCausalLM):
    decoder_layer_cls = HunYuanDecoderLayer


class _HunYuanGate(nn.Module):
    def __init__(self, cfg: ModelConfig, dtype, prefix: str):
        super().__init__()
        from ..layers.linear import ReplicatedLinear

        self.wg = ReplicatedLinear(cfg.hidden_size, cfg.num_local_experts, False, dtype, None, prefix + ".wg")


class HunYuanMoE(nn.Module):
    """soft-max top-k experts (renormalised) plus an always-on shared MLP"""

    def __init__(self, cfg: ModelConfig, layer_id: int, dtype, quant, prefix: str):
        super().__init__()
        from ..layers.moe import FusedMoE
        from .mixtral import _EP_SIZE

        self.gate = _HunYuanGate(cfg, dtype, prefix + ".gate")
        self.experts = FusedMoE(cfg.num_local_experts, cfg.num_experts_per_tok, cfg.hidden_size, cfg.moe_intermediate_size, dtype, True, _EP_SIZE["v"], quant,
                                layer_id)
        self.shared_mlp = LlamaMLP(cfg, dtype, quant, prefix + ".shared_mlp")

    def forward(self, x: torch.Tensor):
        y = _resolve(self.experts(x, self.gate.wg(x)))
        return y + _resolve(self.shared_mlp(x))


def _make_hunyuan_moe():
    from .families import Qwen2MoeForCausalLM

    class HunYuanMoEV1ForCausalLM(Qwen2MoeForCausalLM):          # (its loader already routes ``mlp.experts.<e>.{gate,up,down}_proj``)
        decoder_layer_cls = HunYuanDecoderLayer

        def _mlp_cls(self):
            return lambda c, layer_id, dtype, quant, prefix: HunYuanMoE(c, layer_id, dtype, quant, prefix + ".mlp")

    return HunYuanMoEV1ForCausalLM


# ---------------------------------------------------------------- FlexOlmo: OLMo-2 post-norm blocks (q/k norm over the whole projection) around OLMoE experts
def _make_flex_olmo():
    from .families import Olmo2Attention, Olmo2DecoderLayer

    class FlexOlmoDecoderLayer(Olmo2DecoderLayer):
        def __init__(self, cfg: ModelConfig, layer_id: int, dtype, quant, rope, prefix: str, mlp_cls=LlamaMLP):
            LlamaDecoderLayer.__init__(self, cfg, layer_id, dtype, quant, rope, prefix, OlmoeMoE)
            self.self_attn = Olmo2Attention(cfg, layer_id, dtype, quant, rope, prefix + ".self_attn")
            del self.input_layernorm
            self.post_feedforward_layernorm = RMSNorm(cfg.hidden_size, cfg.rms_norm_eps, dtype)

    class FlexOlmoForCausalLM(OlmoeForCausalLM):
        """reference ``python/sglang/srt/models/flex_olmo.py``"""
        decoder_layer_cls = FlexOlmoDecoderLayer

    return FlexOlmoForCausalLM
