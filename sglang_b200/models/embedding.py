"""Embedding / reward / classification heads on the Llama backbone.

Parity: reference ``models/llama_embedding.py``, ``llama_reward.py``, ``llama_classification.py`` +
``layers/pooler.py`` (last-token pooling with optional L2 normalisation)."""
from __future__ import annotations

from typing import Optional

import torch
from torch import nn

from .llama import _DT, LlamaForCausalLM


class LlamaEmbeddingModel(LlamaForCausalLM):
    is_embedding = True

    def pool(self, hidden_last: torch.Tensor, normalize: bool = True) -> torch.Tensor:
        h = hidden_last.float()
        return torch.nn.functional.normalize(h, dim=-1) if normalize else h


class LlamaForSequenceClassification(LlamaForCausalLM):
    is_embedding = True

    def __init__(self, cfg, quant_name: Optional[str] = None, device="cpu", num_labels: Optional[int] = None):
        super().__init__(cfg, quant_name, device)
        self.score = nn.Linear(cfg.hidden_size, num_labels or cfg.num_labels, bias=False, dtype=_DT[cfg.dtype])

    def pool(self, hidden_last: torch.Tensor, normalize: bool = False) -> torch.Tensor:
        return self.score(hidden_last).float()


def make_classifier(base_cls):
    """``<Family>ForSequenceClassification``: the family's decoder + a bias-free ``score`` head on the last token (reward models, classifiers)"""
    return type(base_cls.__name__.replace("ForCausalLM", "") + "ForSequenceClassification", (LlamaForSequenceClassification, base_cls), {})


class _MlpRewardHead(LlamaForCausalLM):
    """``Qwen2ForRewardModel`` / ``Qwen3ForRewardModel`` (reference ``models/qwen2_rm.py``, ``qwen3_rm.py``): Linear -> ReLU -> Linear on the
    last token, raw score(s) out"""
    is_embedding = True

    def __init__(self, cfg, quant_name: Optional[str] = None, device="cpu"):
        super().__init__(cfg, quant_name, device)
        dt = _DT[cfg.dtype]
        self.score = nn.Sequential(nn.Linear(cfg.hidden_size, cfg.hidden_size, dtype=dt), nn.ReLU(), nn.Linear(cfg.hidden_size, cfg.num_labels, dtype=dt))

    def pool(self, hidden_last: torch.Tensor, normalize: bool = False) -> torch.Tensor:
        return self.score(hidden_last).float()


def make_reward_model(base_cls, name: str):
    return type(name, (_MlpRewardHead, base_cls), {})


class LlamaForSequenceClassificationWithNormal_Weights(LlamaForSequenceClassification):
    """reference ``models/llama_reward.py``: per-label scores mixed by a second, input-dependent ``weights`` head"""

    def __init__(self, cfg, quant_name: Optional[str] = None, device="cpu"):
        super().__init__(cfg, quant_name, device)
        self.weights = nn.Linear(cfg.hidden_size, cfg.num_labels, bias=False, dtype=_DT[cfg.dtype])

    def pool(self, hidden_last: torch.Tensor, normalize: bool = False) -> torch.Tensor:
        s, w = self.score(hidden_last).float(), self.weights(hidden_last).float()
        return (s * w).sum(-1, keepdim=True)


def make_v_head_reward(base_cls, name: str):
    """``InternLM2ForRewardModel`` (reference ``models/internlm2_reward.py``): a bias-free ``v_head`` on the last token"""

    class _VHead(base_cls):
        is_embedding = True

        def __init__(self, cfg, quant_name: Optional[str] = None, device="cpu"):
            super().__init__(cfg, quant_name, device)
            self.v_head = nn.Linear(cfg.hidden_size, 1, bias=False, dtype=_DT[cfg.dtype])

        def pool(self, hidden_last: torch.Tensor, normalize: bool = False) -> torch.Tensor:
            return self.v_head(hidden_last).float()

    _VHead.__name__ = _VHead.__qualname__ = name
    return _VHead
