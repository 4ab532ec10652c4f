"""Mixtral-style sparse MoE decoder (top-k softmax router + SwiGLU experts).

Parity: reference ``python/sglang/srt/models/mixtral.py`` (MixtralMoE block: replicated gate + FusedMoE with
``w1/w3 -> w13`` stacking, expert weight mapping in ``load_weights``).
"""
from __future__ import annotations

from typing import Iterable, List, Tuple

import torch
from torch import nn

from ..configs.model_config import ModelConfig
from ..layers.linear import ReplicatedLinear
from ..layers.moe import FusedMoE
from .llama import LlamaForCausalLM, _layer_id

_EP_SIZE = {"v": 1}


def set_ep_size(n: int) -> None:
    _EP_SIZE["v"] = n


class MixtralMoE(nn.Module):
    def __init__(self, cfg: ModelConfig, layer_id: int, dtype, quant, prefix: str):
        super().__init__()
        self.gate = ReplicatedLinear(cfg.hidden_size, cfg.num_local_experts, False, dtype, None, prefix + ".gate")
        self.experts = FusedMoE(cfg.num_local_experts, cfg.num_experts_per_tok, cfg.hidden_size, cfg.moe_intermediate_size,
                                dtype, cfg.norm_topk_prob, _EP_SIZE["v"], quant, layer_id)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.experts(x, self.gate(x))


class MixtralForCausalLM(LlamaForCausalLM):
    def _mlp_cls(self):
        return MixtralMoE

    def load_weights(self, weights: Iterable[Tuple[str, torch.Tensor]]) -> List[str]:
        rest, loaded = [], []
        params = dict(self.named_parameters())
        modules = dict(self.named_modules())
        for name, w in weights:
            if ".block_sparse_moe.experts." in name:
                lid = _layer_id(name)
                if not (self.model.start_layer <= lid < self.model.end_layer):
                    continue
                local = lid - self.model.start_layer
                eid = int(name.split(".experts.")[1].split(".")[0])
                shard = name.split(".")[-2]  # w1 / w2 / w3
                mod = modules[f"model.layers.{local}.mlp.experts"]
                p = mod.w2 if shard == "w2" else mod.w13
                mod.weight_loader(p, w, eid, shard)
                loaded.append(name)
            elif ".block_sparse_moe.gate." in name:
                rest.append((name.replace(".block_sparse_moe.gate.", ".mlp.gate."), w))
            else:
                rest.append((name, w))
        return loaded + super().load_weights(rest)


# ---------------------------------------------------------------- MiniMax-M2: Mixtral block names, sigmoid + bias router, full-width q/k norm
class _FullWidthRMSNorm(nn.Module):
    """RMSNorm over ALL heads of a projection (MiniMax-M2 ``qk_norm_type = per_layer``).  Heads are sharded over TP ranks, so the mean
    square is completed with one all-reduce of a [T, 1] column; ``replicas`` discounts KV heads that several ranks hold copies of."""

    def __init__(self, total: int, local: int, eps: float, dtype, tp, replicas: int = 1):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(local, dtype=dtype), requires_grad=False)
        self.total, self.local, self.eps, self.tp, self.replicas = total, local, eps, tp, replicas

    def weight_loader(self, p: nn.Parameter, w: torch.Tensor, shard=None) -> None:
        r = self.tp.rank // self.replicas
        p.data.copy_(w.narrow(0, r * self.local, self.local))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        ss = x.float().pow(2).sum(-1, keepdim=True)
        if self.tp.world_size > 1:
            ss = self.tp.all_reduce(ss.contiguous()) / self.replicas
        return (x.float() * torch.rsqrt(ss / self.total + self.eps)).to(x.dtype) * self.weight


class MiniMaxM2MoE(MixtralMoE):
    def __init__(self, cfg: ModelConfig, layer_id: int, dtype, quant, prefix: str):
        super().__init__(cfg, layer_id, dtype, quant, prefix)
        self.top_k = cfg.num_experts_per_tok
        self.e_score_correction_bias = nn.Parameter(torch.zeros(cfg.num_local_experts, dtype=torch.float32), requires_grad=False)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        s = torch.sigmoid(self.gate(x).float())
        ids = (s + self.e_score_correction_bias).topk(self.top_k, dim=-1).indices          # the bias steers the choice, not the weights
        w = s.gather(1, ids)
        w = w / w.sum(-1, keepdim=True)
        return self.experts.forward_routed(x, w.contiguous(), ids.to(torch.int32).contiguous())


class MiniMaxM2ForCausalLM(MixtralForCausalLM):
    """reference ``python/sglang/srt/models/minimax_m2.py``"""

    def __init__(self, cfg: ModelConfig, quant_name=None, device="cpu"):
        super().__init__(cfg, quant_name, device)
        from ..parallel.parallel_state import get_attn_tp_group

        tp = get_attn_tp_group()
        for layer in self.model.layers:
            sa = layer.self_attn
            dt = sa.o_proj.weight.dtype if sa.o_proj.weight.dtype.is_floating_point and sa.o_proj.weight.dtype.itemsize > 1 else torch.bfloat16
            sa.q_norm = _FullWidthRMSNorm(cfg.num_attention_heads * cfg.head_dim, sa.q_size, cfg.rms_norm_eps, dt, tp)
            sa.k_norm = _FullWidthRMSNorm(cfg.num_key_value_heads * cfg.head_dim, sa.kv_size, cfg.rms_norm_eps, dt, tp, sa.qkv_proj.kv_replicas)
            sa.forward = _minimax_attn_forward.__get__(sa)

    def _mlp_cls(self):
        return MiniMaxM2MoE

    def load_weights(self, weights: Iterable[Tuple[str, torch.Tensor]]) -> List[str]:
        def rename():
            for name, w in weights:
                if name.endswith(".block_sparse_moe.e_score_correction_bias"):
                    yield name.replace(".block_sparse_moe.", ".mlp."), w.float()
                elif ".mtp" in name or name.startswith("mtp."):
                    continue
                else:
                    yield name, w
        return super().load_weights(rename())


def _minimax_attn_forward(self, x: torch.Tensor, fb) -> torch.Tensor:
    q, k, v = self.qkv_proj(x).split([self.q_size, self.kv_size, self.kv_size], dim=-1)
    o = self.attn(self.q_norm(q).contiguous(), self.k_norm(k).contiguous(), v.contiguous(), fb)
    return self.o_proj(o, defer_reduce=True)
