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
