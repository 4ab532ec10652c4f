"""Phi-3.5-MoE (``PhimoeForCausalLM``): Mixtral-shaped block names, LayerNorm (with bias) instead of RMSNorm, biased attention projections and LM
head, and the *sparsemixer* router.

Parity: reference ``python/sglang/srt/models/phimoe.py``.  Inference-time sparsemixer: expert 1 is the arg-max logit; its weight is the
soft-max over the experts whose logit lies within a relative band ``2 * jitter`` of that maximum; expert 2 is the arg-max of the remaining
logits with the same banded soft-max computed over the remainder.  The two weights are NOT renormalised."""
from __future__ import annotations

from typing import Iterable, List, Optional, Tuple

import torch
from torch import nn

from ..configs.model_config import ModelConfig
from ..model_executor.forward_batch import ForwardBatch
from .families2 import PlainLayerNorm, _resolve
from .llama import LlamaDecoderLayer
from .mixtral import MixtralForCausalLM, MixtralMoE


def sparsemixer_route(logits: torch.Tensor, jitter: float) -> Tuple[torch.Tensor, torch.Tensor]:
    """logits fp32 [T, E] -> (weights [T, 2], expert ids [T, 2])"""
    def pick(scores, banded_from):
        top, idx = scores.max(-1, keepdim=True)
        factor = banded_from.abs().clamp(min=top)
        out_of_band = ((top - banded_from) / factor) > 2 * jitter
        gates = torch.softmax(scores.masked_fill(out_of_band, float("-inf")), -1)
        return gates.gather(-1, idx), idx

    w1, i1 = pick(logits, logits)
    rest = logits.scatter(-1, i1, float("-inf"))
    w2, i2 = pick(rest, logits)          # (the band test of the second pick compares against the ORIGINAL logits, as the published code does)
    return torch.cat([w1, w2], -1), torch.cat([i1, i2], -1)


class PhiMoE(MixtralMoE):
    def __init__(self, cfg: ModelConfig, layer_id: int, dtype, quant, prefix: str):
        super().__init__(cfg, layer_id, dtype, quant, prefix)
        self.jitter = float(getattr(cfg, "router_jitter_noise", None) or 0.01)
        assert cfg.num_experts_per_tok == 2, "sparsemixer is a top-2 router"

    def forward(self, x: torch.Tensor):
        w, ids = sparsemixer_route(self.gate(x).float(), self.jitter)
        return self.experts.forward_routed(x, w.float().contiguous(), ids.to(torch.int32).contiguous())


class _PhiMoeLayer(LlamaDecoderLayer):
    def __init__(self, cfg: ModelConfig, layer_id: int, dtype, quant, rope, prefix: str, mlp_cls=None):
        super().__init__(cfg, layer_id, dtype, quant, rope, prefix, mlp_cls)
        self.input_layernorm = PlainLayerNorm(cfg.hidden_size, cfg.rms_norm_eps, dtype, True)
        self.post_attention_layernorm = PlainLayerNorm(cfg.hidden_size, cfg.rms_norm_eps, dtype, True)

    def forward(self, h, residual, fb: ForwardBatch):
        if residual is not None:
            h = h + residual
        h = h + _resolve(self.self_attn(self.input_layernorm(h), fb))
        return h + _resolve(self.mlp(self.post_attention_layernorm(h))), None


class PhimoeForCausalLM(MixtralForCausalLM):
    decoder_layer_cls = _PhiMoeLayer

    def __init__(self, cfg: ModelConfig, quant_name: Optional[str] = None, device="cpu"):
        super().__init__(cfg, quant_name, device)
        dt = self.model.embed_tokens.weight.dtype
        self.model.norm = PlainLayerNorm(cfg.hidden_size, cfg.rms_norm_eps, dt, True)
        self.lm_head_bias = nn.Parameter(torch.zeros(cfg.vocab_size, dtype=dt), requires_grad=False)

    def _mlp_cls(self):
        return PhiMoE

    def compute_logits(self, hidden: torch.Tensor, key: str = "logits") -> torch.Tensor:
        y = super().compute_logits(hidden, key)
        return y + self.lm_head_bias[: y.shape[-1]].to(y.dtype)

    def load_weights(self, weights: Iterable[Tuple[str, torch.Tensor]]) -> List[str]:
        loaded = []

        def rest():
            for name, w in weights:
                if name == "lm_head.bias":
                    self.lm_head_bias.data.copy_(w)
                    loaded.append(name)
                else:
                    yield name, w
        out = super().load_weights(rest())
        return loaded + out
