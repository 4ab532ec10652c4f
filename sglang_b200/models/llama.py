"""Decoder-only transformer family: Llama / Mistral / Qwen2 / Qwen3 / Gemma (dense) — one implementation driven by
``ModelConfig`` flags (attention bias, qk-norm, gemma norm/activation/embedding scale, tied embeddings).

Parity: reference ``python/sglang/srt/models/llama.py`` (``LlamaMLP:111``, ``LlamaAttention:240``,
``LlamaDecoderLayer:341``, ``LlamaModel:419``, ``LlamaForCausalLM:496`` ``forward:563`` ``load_weights:663``),
``qwen2.py``, ``qwen3.py``, ``gemma.py``.  Pipeline parallelism: a stage owns layers [start, end) (``llama.py:640-644``).
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Optional, Tuple

import torch
from torch import nn

from .. import ops
from ..configs.model_config import ModelConfig
from ..layers.common import RadixAttention, RMSNorm, RotaryEmbedding, get_act_and_mul
from ..layers.linear import (MergedColumnParallelLinear, ParallelLMHead, QKVParallelLinear, RowParallelLinear,
                             VocabParallelEmbedding, _skinny_allowed)
from ..layers.quantization import UnquantizedMethod, get_quant_method
from ..model_executor.forward_batch import ForwardBatch
from ..environ import envs
from ..layers.dp_attention import get_dp_ctx
from ..layers.sp import get_sp_ctx
from ..parallel.parallel_state import get_attn_tp_group, get_pp_group, get_tp_group

_DT = {"bfloat16": torch.bfloat16, "float16": torch.float16, "float32": torch.float32}


def pp_layer_range(num_layers: int, pp_rank: int, pp_size: int) -> Tuple[int, int]:
    per, rem = divmod(num_layers, pp_size)
    start = pp_rank * per + min(pp_rank, rem)
    return start, start + per + (1 if pp_rank < rem else 0)


class LlamaMLP(nn.Module):
    def __init__(self, cfg: ModelConfig, dtype, quant, prefix: str):
        super().__init__()
        self.gate_up_proj = MergedColumnParallelLinear(cfg.hidden_size, [cfg.intermediate_size] * 2, cfg.mlp_bias, dtype, quant,
                                                       prefix + ".gate_up_proj")
        self.down_proj = RowParallelLinear(cfg.intermediate_size, cfg.hidden_size, cfg.mlp_bias, True, dtype, quant,
                                           prefix + ".down_proj")
        self.act = get_act_and_mul(cfg.hidden_act)
        self.fuse_silu = cfg.hidden_act == "silu" and not cfg.mlp_bias

    def __new__(cls, cfg: ModelConfig, *a, **kw):
        if cls is LlamaMLP and not cfg.mlp_gated:
            return super().__new__(PlainMLP)
        return super().__new__(cls)

    def forward(self, x):
        gu = self.gate_up_proj
        if (self.fuse_silu and x.dim() == 2 and getattr(gu, "fast_path_ok", True) and isinstance(gu.quant, UnquantizedMethod)
                and ops.gemm_impl() != "lib" and ops.streamk_ok(x, gu.weight) and gu.weight.shape[0] % 16 == 0):
            # decode shapes: ONE persistent stream-K tcgen05 kernel for gate_up_proj + SiLU*mul (tiles pair 64 gate rows with their up rows)
            h = ops.gemm_streamk(x, gu.weight, silu_split=True)
        elif self.fuse_silu and x.dim() == 2 and _skinny_allowed(self.gate_up_proj, x):
            # decode shapes: ONE tcgen05 kernel for gate_up_proj + SiLU*mul ([gate; up] rows paired inside each tile)
            h = ops.gemm_skinny(x, self.gate_up_proj.weight, silu_split=True)
        else:
            h = self.act(self.gate_up_proj(x))
        return self.down_proj(h, defer_reduce=True)


class PlainMLP(LlamaMLP):
    """non-gated MLP (Arcee, Nemotron): up_proj -> activation -> down_proj"""

    def __init__(self, cfg: ModelConfig, dtype, quant, prefix: str):
        nn.Module.__init__(self)
        from ..layers.common import get_act
        from ..layers.linear import ColumnParallelLinear

        self.up_proj = ColumnParallelLinear(cfg.hidden_size, cfg.intermediate_size, cfg.mlp_bias, False, dtype, quant, prefix + ".up_proj")
        self.down_proj = RowParallelLinear(cfg.intermediate_size, cfg.hidden_size, cfg.mlp_bias, True, dtype, quant, prefix + ".down_proj")
        self.act = get_act(cfg.hidden_act)

    def forward(self, x):
        return self.down_proj(self.act(self.up_proj(x)), defer_reduce=True)


class LlamaAttention(nn.Module):
    def __init__(self, cfg: ModelConfig, layer_id: int, dtype, quant, rope: RotaryEmbedding, prefix: str):
        super().__init__()
        atp = get_attn_tp_group()          # == the TP group unless DP-attention is on
        self.qkv_proj = QKVParallelLinear(cfg.hidden_size, cfg.head_dim, cfg.num_attention_heads, cfg.num_key_value_heads,
                                          cfg.attention_bias, dtype, quant, prefix + ".qkv_proj", tp_group=atp)
        self.o_proj = RowParallelLinear(cfg.num_attention_heads * cfg.head_dim, cfg.hidden_size, bool(cfg.attention_out_bias), True, dtype, quant,
                                        prefix + ".o_proj", tp_group=atp)
        self.q_size, self.kv_size = self.qkv_proj.q_size, self.qkv_proj.kv_size
        self.head_dim = cfg.head_dim
        self.qk_norm = cfg.qk_norm
        gid = int(prefix.split(".layers.")[1].split(".")[0]) if ".layers." in prefix else layer_id      # global layer index
        if cfg.no_rope_layers is not None and gid < len(cfg.no_rope_layers) and not cfg.no_rope_layers[gid]:
            rope = None                                                                                  # SmolLM3 NoPE layers
        if cfg.qk_norm:
            self.q_norm = RMSNorm(cfg.head_dim, cfg.rms_norm_eps, dtype)
            self.k_norm = RMSNorm(cfg.head_dim, cfg.rms_norm_eps, dtype)
        window = cfg.sliding_window if (cfg.sliding_window and cfg.sliding_window > 0) else -1
        if cfg.layer_types is not None and gid < len(cfg.layer_types) and cfg.layer_types[gid] != "sliding_attention":
            window = -1
        self.attn = RadixAttention(self.qkv_proj.num_heads, cfg.head_dim, self.qkv_proj.num_kv_heads, layer_id,
                                   rope=rope, sliding_window=window, logit_cap=cfg.attn_logit_softcapping)

    def forward(self, x: torch.Tensor, fb: ForwardBatch) -> torch.Tensor:
        if not self.qk_norm and self.attn.rope is not None and x.dim() == 2 and _skinny_allowed(self.qkv_proj, x) and self.o_proj.bias is None:
            # decode shapes: split-K QKV GEMM; the RoPE + KV-write kernel folds the partial slabs while it reads q/k/v
            slabs = ops.gemm_skinny_slabs(x, self.qkv_proj.weight, self.qkv_proj.bias)
            q, k, v = slabs[0].split([self.q_size, self.kv_size, self.kv_size], dim=-1)
            o = self.attn(q, k, v, fb, nslabs=slabs.shape[0], slab_stride=slabs.stride(0))
            return self.o_proj(o, defer_reduce=True)
        qkv = self.qkv_proj(x)
        q, k, v = qkv.split([self.q_size, self.kv_size, self.kv_size], dim=-1)
        if self.qk_norm:
            T = q.shape[0]
            q = self.q_norm(q.reshape(T * self.attn.num_heads, self.head_dim)).view(T, self.q_size)
            k = self.k_norm(k.reshape(T * self.attn.num_kv_heads, self.head_dim)).view(T, self.kv_size)
        o = self.attn(q, k, v, fb)
        return self.o_proj(o, defer_reduce=True)


class LlamaDecoderLayer(nn.Module):
    def __init__(self, cfg: ModelConfig, layer_id: int, dtype, quant, rope, prefix: str, mlp_cls=LlamaMLP):
        super().__init__()
        gemma = ("Gemma" in cfg.architecture)
        self.self_attn = LlamaAttention(cfg, layer_id, dtype, quant, rope, prefix + ".self_attn")
        self.mlp = mlp_cls(cfg, dtype, quant, prefix + ".mlp") if mlp_cls is LlamaMLP else mlp_cls(cfg, layer_id, dtype, quant, prefix)
        self.input_layernorm = RMSNorm(cfg.hidden_size, cfg.rms_norm_eps, dtype, gemma)
        self.post_attention_layernorm = RMSNorm(cfg.hidden_size, cfg.rms_norm_eps, dtype, gemma)

    def forward(self, h: torch.Tensor, residual: Optional[torch.Tensor], fb: ForwardBatch):
        if residual is None:
            residual = h
            h = self.input_layernorm(h)
        else:
            h, residual = self.input_layernorm(h, residual)
        sp = get_sp_ctx()
        if sp is not None:
            # sequence parallel: h / residual are this rank's rows; gather before each tensor-parallel block, whose row-parallel GEMM
            # (o_proj, down_proj) hands back reduce-scattered rows
            h = self.self_attn(sp.gather(h)[: sp.T], fb)
            h, residual = self.post_attention_layernorm(h, residual)
            return self.mlp(sp.gather(h)), residual
        h = self.self_attn(h, fb)
        h, residual = self.post_attention_layernorm(h, residual)
        ctx = get_dp_ctx()
        if ctx is not None:
            experts = getattr(self.mlp, "experts", None)
            if (experts is not None and envs.SGLB_MOE_A2A.get() and experts.ep_size == experts.tp.world_size and experts.moe_tp == 1
                    and getattr(self.mlp, "shared_experts", None) is None and getattr(self.mlp, "shared_expert", None) is None):
                # expert-parallel MoE without shared experts: tokens stay home, only routed (token, k) pairs cross ranks (all-to-all)
                ctx.a2a_active = True
                try:
                    return self.mlp(h), residual
                finally:
                    ctx.a2a_active = False
            # DP-attention: the MLP / MoE is parallel over the whole TP group -> run it on the union of all groups' tokens
            y = self.mlp(ctx.gather(h, "tokens"))
            y = y.resolve() if hasattr(y, "resolve") else y
            return ctx.slice(y, "tokens"), residual
        h = self.mlp(h)
        return h, residual


class LlamaModel(nn.Module):
    def __init__(self, cfg: ModelConfig, quant_name: Optional[str] = None, mlp_cls=LlamaMLP, device="cpu", layer_cls=None):
        layer_cls = layer_cls or LlamaDecoderLayer
        super().__init__()
        self.cfg = cfg
        dtype = _DT[cfg.dtype]
        quant_factory = lambda: get_quant_method(quant_name)
        pp = get_pp_group()
        self.pp = pp
        self.start_layer, self.end_layer = pp_layer_range(cfg.num_hidden_layers, pp.rank, pp.world_size)
        self.gemma = ("Gemma" in cfg.architecture)
        self.embed_tokens = VocabParallelEmbedding(cfg.vocab_size, cfg.hidden_size, dtype) if pp.is_first else None
        rope = RotaryEmbedding(cfg.head_dim, cfg.max_position_embeddings, cfg.rope_theta, cfg.rope_scaling, cfg.rope_is_neox,
                               cfg.partial_rotary_factor, device)
        self.layers = nn.ModuleList([
            layer_cls(cfg, i - self.start_layer, dtype, quant_factory(), rope, f"model.layers.{i}", mlp_cls)
            for i in range(self.start_layer, self.end_layer)])
        self.norm = RMSNorm(cfg.hidden_size, cfg.rms_norm_eps, dtype, self.gemma) if pp.is_last else None

    def embed(self, fb: ForwardBatch) -> torch.Tensor:
        ctx = get_dp_ctx()
        if ctx is not None:        # vocab-parallel over the full TP group: embed the union of the groups' token ids, keep ours
            ids = fb.next_token_buf[fb.req_slots.long()] if fb.input_ids is None else fb.input_ids
            h = ctx.slice(self.embed_tokens(ctx.gather(ids.to(torch.int32), "tokens")), "tokens")
        elif fb.input_ids is None:   # decode: token ids live on the device, indexed by request slot
            h = self.embed_tokens(fb.next_token_buf, fb.req_slots)
        else:
            h = self.embed_tokens(fb.input_ids)
        if self.gemma:
            h = h * (self.cfg.hidden_size ** 0.5)
        if fb.mm_embeds is not None:
            rows, emb = fb.mm_embeds
            h.index_copy_(0, rows, emb.to(h.dtype))
        return h

    def forward(self, fb: ForwardBatch, inputs_embeds: Optional[torch.Tensor] = None) -> torch.Tensor:
        if self.pp.is_first:
            h = inputs_embeds if inputs_embeds is not None else self.embed(fb)
            residual = None
        else:
            h, residual = fb.pp_hidden
        sp = get_sp_ctx()
        if sp is not None:
            h = sp.scatter(h)
        cap = fb.capture_aux
        aux = [] if cap is not None else None
        hook = getattr(fb, "post_layer_hook", None)
        for i, layer in enumerate(self.layers):
            if cap is not None and (i + self.start_layer) in cap:       # EAGLE-3 auxiliary features: the residual stream entering this layer
                h = h.resolve() if hasattr(h, "resolve") else h
                aux.append(h if residual is None else h + residual)
            h, residual = layer(h, residual, fb)
            if hook is not None:         # DeepStack (Qwen3-VL): visual features join the residual stream after the first few layers
                h = hook(i + self.start_layer, h)
        if aux:
            fb.aux_hidden = torch.cat(aux, dim=-1)
        if not self.pp.is_last:
            if hasattr(h, "resolve"):
                h = h.resolve()
            return h, residual
        if residual is None:      # families whose layers keep the residual stream in h (Gemma-2, Granite, OLMo-2)
            return self.norm(h.resolve() if hasattr(h, "resolve") else h)
        h, _ = self.norm(h, residual)
        if sp is not None:
            h = sp.gather(h)[: sp.T]
        return h


class LlamaForCausalLM(nn.Module):
    decoder_layer_cls = None      # subclasses may swap the decoder layer (models/families.py)
    # checkpoint name -> (merged param name, shard id)   (reference llama.py:663 stacked_params_mapping)
    stacked_params = [(".qkv_proj", ".q_proj", "q"), (".qkv_proj", ".k_proj", "k"), (".qkv_proj", ".v_proj", "v"),
                      (".gate_up_proj", ".gate_proj", 0), (".gate_up_proj", ".up_proj", 1)]

    def __init__(self, cfg: ModelConfig, quant_name: Optional[str] = None, device="cpu"):
        super().__init__()
        self.cfg = cfg
        if not cfg.mlp_gated:       # checkpoints of non-gated MLPs have a real ".up_proj": do not fold it into a gate_up pair
            self.stacked_params = [t for t in self.stacked_params if t[0] != ".gate_up_proj"]
        self.model = LlamaModel(cfg, quant_name, self._mlp_cls(), device, getattr(self, "decoder_layer_cls", None))
        pp = get_pp_group()
        if pp.is_last:
            if cfg.tie_word_embeddings and pp.is_first:
                self.lm_head = None
            else:
                self.lm_head = ParallelLMHead(cfg.vocab_size, cfg.hidden_size, _DT[cfg.dtype])
        else:
            self.lm_head = None

    def _mlp_cls(self):
        return LlamaMLP

    def forward(self, fb: ForwardBatch, inputs_embeds=None):
        return self.model(fb, inputs_embeds)

    def compute_logits(self, hidden: torch.Tensor, key: str = "logits") -> torch.Tensor:
        """``key`` names the row set under DP-attention (sampled rows vs all rows of a prompt-logprob pass): every group must make
        the same sequence of calls per forward, with zero rows when it has nothing to contribute"""
        head = self.lm_head if self.lm_head is not None else self.model.embed_tokens
        ctx = get_dp_ctx()
        if ctx is not None:
            hidden = ctx.gather(hidden, key)
        y = ops.linear(hidden, head.weight)
        tp = get_tp_group()
        if tp.world_size > 1:
            y = tp.all_gather(y, dim=-1)
        if ctx is not None:
            y = ctx.slice(y, key)
        y = y[:, : self.cfg.vocab_size]
        if self.cfg.final_logit_softcapping > 0:
            c = self.cfg.final_logit_softcapping
            y = torch.tanh(y.float() / c) * c
        return y

    # ---- weights
    def named_weight_targets(self) -> Dict[str, torch.Tensor]:
        return dict(self.named_parameters())

    def load_weights(self, weights: Iterable[Tuple[str, torch.Tensor]]) -> List[str]:
        params = dict(self.named_parameters())
        modules = dict(self.named_modules())
        loaded: List[str] = []
        for name, w in weights:
            if "rotary_emb.inv_freq" in name:
                continue
            lid = _layer_id(name)
            if lid is not None:
                if not (self.model.start_layer <= lid < self.model.end_layer):
                    continue
                name = name.replace(f"layers.{lid}.", f"layers.{lid - self.model.start_layer}.")
            if name.startswith("lm_head") and self.lm_head is None:
                continue
            target, shard = name, None
            for merged, orig, sid in self.stacked_params:
                if orig in name:
                    target, shard = name.replace(orig, merged), sid
                    break
            if target not in params:
                continue
            p = params[target]
            mod = modules[target.rsplit(".", 1)[0]]
            if hasattr(mod, "weight_loader"):
                if target.endswith(".bias") and shard is None and p.dim() == 1 and not isinstance(mod, RowParallelLinear):
                    n = p.shape[0]
                    p.data.copy_(w.narrow(0, get_tp_group().rank * n, n) if w.shape[0] != n else w)
                elif target.endswith(".bias") and shard is not None:
                    _load_stacked_bias(mod, p, w, shard)
                else:
                    mod.weight_loader(p, w, shard) if shard is not None else mod.weight_loader(p, w)
            else:
                p.data.copy_(w)
            loaded.append(name)
        return loaded


def _load_stacked_bias(mod, p, w, shard):
    fake = w.unsqueeze(1)
    mod.weight_loader(p.data.unsqueeze(1), fake, shard)


def _layer_id(name: str) -> Optional[int]:
    parts = name.split(".")
    for i, s in enumerate(parts):
        if s == "layers" and i + 1 < len(parts) and parts[i + 1].isdigit():
            return int(parts[i + 1])
    return None


class Qwen2ForCausalLM(LlamaForCausalLM):
    pass


class Qwen3ForCausalLM(LlamaForCausalLM):
    pass


class MistralForCausalLM(LlamaForCausalLM):
    pass


class GemmaForCausalLM(LlamaForCausalLM):
    pass
