"""Llama-4 (text): interleaved RoPE / NoPE attention layers — RoPE layers attend inside fixed *chunks* and L2-normalise q/k after the rotation,
NoPE layers attend globally with a length-dependent query temperature — and MoE layers whose routed expert sees ``sigmoid(score) * x`` next to
an always-on shared expert.

Parity: reference ``python/sglang/srt/models/llama4.py``; checked against HF ``Llama4ForCausalLM`` (``tests/test_hf_parity.py``).

Engine notes: q/k post-processing happens after the rotation, so this family rotates in PyTorch and hands ``RadixAttention`` un-roped tensors
(``rope=None`` path: plain KV write); chunked layers express "same chunk and causal" as an attention mask built from the scheduler's host
mirrors, which routes them to the reference attention path (a chunk-window variant of the tcgen05 kernel is the follow-up).
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Tuple

import torch
import torch.nn.functional as F
from torch import nn

from ..configs.model_config import ModelConfig
from ..layers.common import RadixAttention
from ..layers.linear import MergedColumnParallelLinear, QKVParallelLinear, ReplicatedLinear, RowParallelLinear
from ..model_executor.forward_batch import ForwardBatch
from ..ops import torch_ref
from ..parallel.parallel_state import get_attn_tp_group, get_tp_group
from .llama import LlamaDecoderLayer, LlamaForCausalLM, LlamaMLP, _layer_id


def _l2(x: torch.Tensor, eps: float) -> torch.Tensor:
    f = x.float()
    return (f * torch.rsqrt(f.pow(2).mean(-1, keepdim=True) + eps)).to(x.dtype)


class Llama4Attention(nn.Module):
    def __init__(self, cfg: ModelConfig, layer_id: int, gid: int, dtype, quant, rope, prefix: str):
        super().__init__()
        atp = get_attn_tp_group()
        self.qkv_proj = QKVParallelLinear(cfg.hidden_size, cfg.head_dim, cfg.num_attention_heads, cfg.num_key_value_heads, cfg.attention_bias, dtype,
                                          quant, prefix + ".qkv_proj", tp_group=atp)
        self.o_proj = RowParallelLinear(cfg.num_attention_heads * cfg.head_dim, cfg.hidden_size, cfg.attention_bias, True, dtype, quant,
                                        prefix + ".o_proj", tp_group=atp)
        self.q_size, self.kv_size, self.hd = self.qkv_proj.q_size, self.qkv_proj.kv_size, cfg.head_dim
        x = cfg.llama4
        self.use_rope = bool(x["no_rope_layers"][gid])
        self.qk_norm = bool(x.get("use_qk_norm", True)) and self.use_rope
        self.temp = bool(x.get("attn_temperature_tuning", True)) and not self.use_rope
        self.floor_scale, self.attn_scale, self.eps = float(x.get("floor_scale", 8192)), float(x.get("attn_scale", 0.1)), cfg.rms_norm_eps
        self.cos_sin = rope.cos_sin
        self.attn = RadixAttention(self.qkv_proj.num_heads, cfg.head_dim, self.qkv_proj.num_kv_heads, layer_id, rope=None)
        kinds = cfg.layer_types or ["chunked_attention" if self.use_rope else "full_attention"] * (gid + 1)
        self.attn.chunk = int(x.get("attention_chunk_size") or 0) if kinds[gid] == "chunked_attention" else 0

    def forward(self, h: torch.Tensor, fb: ForwardBatch):
        T = h.shape[0]
        q, k, v = self.qkv_proj(h).split([self.q_size, self.kv_size, self.kv_size], dim=-1)
        q, k = q.reshape(T, -1, self.hd), k.reshape(T, -1, self.hd)
        if self.use_rope:
            q = torch_ref.apply_rope(q, fb.positions, self.cos_sin, False)         # interleaved pairs (complex multiplication in the reference)
            k = torch_ref.apply_rope(k, fb.positions, self.cos_sin, False)
        if self.qk_norm:
            q, k = _l2(q, self.eps), _l2(k, self.eps)
        if self.temp:           # NoPE layers: queries sharpen logarithmically with the position
            s = torch.log1p(torch.floor((fb.positions.float() + 1.0) / self.floor_scale)) * self.attn_scale + 1.0
            q = (q.float() * s[:, None, None]).to(q.dtype)
        o = self.attn(q.reshape(T, self.q_size), k.reshape(T, self.kv_size), v.contiguous(), fb)
        return self.o_proj(o)


class Llama4MoE(nn.Module):
    """top-k (1) router, ``sigmoid`` score applied to the expert INPUT; shared expert always on.  Experts are TP-sharded along the intermediate dim."""

    def __init__(self, cfg: ModelConfig, layer_id: int, dtype, quant, prefix: str):
        super().__init__()
        tp = get_tp_group()
        self.tp = tp
        E, H, I = cfg.num_local_experts, cfg.hidden_size, cfg.intermediate_size
        assert I % tp.world_size == 0
        self.E, self.k, self.Il = E, cfg.num_experts_per_tok, I // tp.world_size
        self.router = ReplicatedLinear(H, E, False, dtype, None, prefix + ".router")
        mk = lambda *s: nn.Parameter(torch.zeros(*s, dtype=dtype), requires_grad=False)
        self.w_gate, self.w_up, self.w_down = mk(E, self.Il, H), mk(E, self.Il, H), mk(E, H, self.Il)
        self.shared_gate_up = MergedColumnParallelLinear(H, [I, I], False, dtype, quant, prefix + ".shared_expert.gate_up_proj")
        self.shared_down = RowParallelLinear(I, H, False, False, dtype, quant, prefix + ".shared_expert.down_proj")

    def load(self, leaf: str, w: torch.Tensor) -> bool:
        r, Il = self.tp.rank, self.Il
        if leaf == "gate_up_proj":            # [E, H, 2I]: gate | up
            g, u = w.transpose(1, 2).chunk(2, dim=1)
            self.w_gate.data.copy_(g[:, r * Il:(r + 1) * Il])
            self.w_up.data.copy_(u[:, r * Il:(r + 1) * Il])
        elif leaf == "down_proj":             # [E, I, H]
            self.w_down.data.copy_(w.transpose(1, 2)[:, :, r * Il:(r + 1) * Il])
        else:
            return False
        return True

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        T, H = x.shape
        logits = self.router(x).float()
        v, ids = torch.topk(logits, self.k, dim=-1)
        score = torch.sigmoid(v).to(x.dtype)                               # [T, k]
        g, u = self.shared_gate_up(x).chunk(2, -1)
        out = self.shared_down(F.silu(g) * u).float()
        flat = ids.reshape(-1)
        order = torch.argsort(flat, stable=True)
        counts = torch.bincount(flat, minlength=self.E).tolist()
        tok, sc = order // self.k, score.reshape(-1)[order]
        s = 0
        for e, n in enumerate(counts):
            if n:
                rows = tok[s:s + n]
                xe = x[rows] * sc[s:s + n, None]                            # the score scales the expert's INPUT
                hh = F.silu(xe @ self.w_gate[e].t()) * (xe @ self.w_up[e].t())
                out.index_add_(0, rows, (hh @ self.w_down[e].t()).float())
                s += n
        out = out.to(x.dtype)
        return self.tp.all_reduce(out) if self.tp.world_size > 1 else out


class Llama4DecoderLayer(LlamaDecoderLayer):
    def __init__(self, cfg: ModelConfig, layer_id: int, dtype, quant, rope, prefix: str, mlp_cls=LlamaMLP):
        import dataclasses

        gid = int(prefix.rsplit(".", 1)[1])
        moe = gid in (cfg.llama4.get("moe_layers") or [])
        dense_cfg = dataclasses.replace(cfg, intermediate_size=int(cfg.llama4.get("intermediate_size_mlp") or cfg.intermediate_size))
        super().__init__(dense_cfg, layer_id, dtype, quant, rope, prefix, LlamaMLP)
        self.self_attn = Llama4Attention(cfg, layer_id, gid, dtype, quant, rope, prefix + ".self_attn")
        if moe:
            self.mlp = Llama4MoE(cfg, layer_id, dtype, quant, prefix + ".mlp")

    def forward(self, h, residual, fb: ForwardBatch):
        if residual is not None:
            h = h + residual
        a = self.self_attn(self.input_layernorm(h), fb)
        h = h + (a.resolve() if hasattr(a, "resolve") else a)
        m = self.mlp(self.post_attention_layernorm(h))
        return h + (m.resolve() if hasattr(m, "resolve") else m), None


class Llama4ForCausalLM(LlamaForCausalLM):
    decoder_layer_cls = Llama4DecoderLayer
    supports_cuda_graph = False          # chunk masks are built on the host

    def load_weights(self, weights: Iterable[Tuple[str, torch.Tensor]]) -> List[str]:
        mods = dict(self.named_modules())
        rest, loaded = [], []
        for name, w in weights:
            n = name[len("language_model."):] if name.startswith("language_model.") else name
            n = n.replace(".feed_forward.", ".mlp.")
            lid = _layer_id(n) if ".layers." in n else None
            if lid is not None and not (self.model.start_layer <= lid < self.model.end_layer):
                continue
            local = None if lid is None else lid - self.model.start_layer
            if ".mlp.experts." in n:
                if mods[f"model.layers.{local}.mlp"].load(n.rsplit(".", 1)[1], w):
                    loaded.append(name)
            elif ".mlp.shared_expert." in n:
                m = mods[f"model.layers.{local}.mlp"]
                leaf = n.split(".shared_expert.")[1]
                if leaf.startswith("gate_proj"):
                    m.shared_gate_up.weight_loader(m.shared_gate_up.weight, w, 0)
                elif leaf.startswith("up_proj"):
                    m.shared_gate_up.weight_loader(m.shared_gate_up.weight, w, 1)
                else:
                    m.shared_down.weight_loader(m.shared_down.weight, w)
                loaded.append(name)
            else:
                rest.append((n, w))
        return loaded + super().load_weights(rest)
