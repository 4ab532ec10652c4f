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
