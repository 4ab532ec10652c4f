"""Expert-distribution recorder + placement rebalancer (EPLB-lite).

Parity: reference ``python/sglang/srt/eplb/`` (``expert_distribution.py`` recorder behind
``/start|stop|dump_expert_distribution_record``; ``eplb_algorithms`` greedy balanced placement of logical experts onto
EP ranks).  The recorder is hooked into ``FusedMoE.forward`` (``layer.expert_load``); ``rebalance`` returns a
logical->physical permutation that equalises per-rank load (longest-processing-time greedy), which the weight loader can
apply by permuting the expert dimension of ``w13``/``w2`` and the router's output ids.
"""
from __future__ import annotations

from typing import Any, Dict, Optional

import numpy as np
import torch

from ..managers import io_struct as ios


class ExpertDistributionRecorder:
    def __init__(self, num_layers: int, num_experts: int):
        self.counts = np.zeros((num_layers, num_experts), dtype=np.int64)
        self.enabled = False

    def record(self, layer_id: int, topk_ids: torch.Tensor) -> None:
        if not self.enabled:
            return
        c = torch.bincount(topk_ids.reshape(-1).long(), minlength=self.counts.shape[1])[: self.counts.shape[1]]
        self.counts[layer_id] += c.cpu().numpy()

    def dump(self) -> Dict[str, Any]:
        tot = self.counts.sum(1, keepdims=True).clip(min=1)
        return {"counts": self.counts.tolist(), "imbalance": (self.counts.max(1) / (tot[:, 0] / self.counts.shape[1])).tolist()}


def rebalance(counts: np.ndarray, ep_size: int) -> np.ndarray:
    """per layer: permutation ``perm[l, physical_slot] = logical_expert`` with slots grouped by EP rank, balancing load"""
    L, E = counts.shape
    per = E // ep_size
    perm = np.zeros((L, E), dtype=np.int64)
    for l in range(L):
        order = np.argsort(-counts[l])
        load = np.zeros(ep_size)
        fill = [[] for _ in range(ep_size)]
        for e in order:
            cands = [r for r in range(ep_size) if len(fill[r]) < per]
            r = min(cands, key=lambda r: load[r])
            fill[r].append(int(e))
            load[r] += counts[l, e]
        perm[l] = np.concatenate([np.asarray(f) for f in fill])
    return perm


def attach(runner: Any) -> Optional[ExpertDistributionRecorder]:
    layers = [m for m in runner.model.modules() if m.__class__.__name__ == "FusedMoE"]
    if not layers:
        return None
    rec = ExpertDistributionRecorder(len(layers), layers[0].num_experts)
    for m in layers:
        m.expert_load = rec
    runner.expert_recorder = rec
    return rec


def handle_request(runner: Any, msg: ios.ExpertDistributionReq) -> ios.GenericResp:
    rec = runner.expert_recorder or attach(runner)
    if rec is None:
        return ios.GenericResp(False, "model has no MoE layers")
    if msg.action == "start":
        rec.enabled = True
        rec.counts[:] = 0
        return ios.GenericResp(True, "recording")
    if msg.action == "stop":
        rec.enabled = False
        return ios.GenericResp(True, "stopped")
    return ios.GenericResp(True, "dump", rec.dump())
