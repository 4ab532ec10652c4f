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
    """Per-layer logical-expert token counts.

    * accumulation happens ON THE DEVICE (``scatter_add_`` into one int64 [layers, experts] tensor): no host sync on the forward path and
      safe inside captured graphs; the host only reads it when somebody asks (``dump`` / the rebalance manager);
    * ``mode``: ``stat`` (cumulative counts), ``window`` (counts of the last ``window`` recorded forward passes: a ring of per-pass deltas),
      ``per_pass`` (additionally keeps every pass on the host for offline analysis, bounded by ``max_passes``);
    * ``balancedness`` = mean / max rank load per layer under a given placement (1.0 = perfectly even), the number the reference logs.
    Reference: eplb/expert_distribution.py (recorder modes stat / stat_approx / per_pass / per_token), ``--expert-distribution-recorder-mode``."""

    def __init__(self, num_layers: int, num_experts: int, mode: str = "stat", window: int = 64, max_passes: int = 1024, device: Any = "cpu"):
        assert mode in ("stat", "window", "per_pass"), mode
        self.L, self.E, self.mode, self.window, self.max_passes = num_layers, num_experts, mode, max(1, window), max_passes
        self.device = device
        self.dev_counts = torch.zeros(num_layers, num_experts, dtype=torch.int64, device=device)
        self._last = np.zeros((num_layers, num_experts), dtype=np.int64)        # cumulative counts at the last ``end_pass``
        self._ring: list = []
        self.passes: list = []
        self.num_passes = 0
        self.enabled = False

    # kept for callers that used the round-1 attribute
    @property
    def counts(self) -> np.ndarray:
        return self.dev_counts.cpu().numpy().copy()

    def reset(self) -> None:
        self.dev_counts.zero_()
        self._last[:] = 0
        self._ring, self.passes, self.num_passes = [], [], 0

    def record(self, layer_id: int, topk_ids: torch.Tensor) -> None:
        if not self.enabled:
            return
        ids = topk_ids.reshape(-1).long().clamp_(0, self.E - 1)
        self.dev_counts[layer_id].scatter_add_(0, ids, torch.ones_like(ids))

    def end_pass(self) -> None:
        """called once per forward pass by the scheduler when a windowed / per-pass mode is on (one small D2H copy, off the critical path)"""
        if not self.enabled or self.mode == "stat":
            return
        cur = self.dev_counts.cpu().numpy().copy()        # copy: on a CPU device .cpu() aliases the live tensor
        delta = cur - self._last
        self._last = cur
        self.num_passes += 1
        self._ring.append(delta)
        if len(self._ring) > self.window:
            self._ring.pop(0)
        if self.mode == "per_pass" and len(self.passes) < self.max_passes:
            self.passes.append(delta)

    def load(self) -> np.ndarray:
        """[layers, experts] counts the rebalancer should plan on (windowed modes: the recent window only)"""
        if self.mode == "stat" or not self._ring:
            return self.dev_counts.cpu().numpy().copy()
        return np.sum(self._ring, axis=0)

    @staticmethod
    def balancedness(counts: np.ndarray, ep_size: int, placement: Optional[np.ndarray] = None) -> np.ndarray:
        """per layer mean / max load over EP ranks; ``placement[l, slot] = logical expert`` (default: experts in order, contiguous per rank)"""
        L, E = counts.shape
        out = np.ones(L)
        for l in range(L):
            order = placement[l] if placement is not None else np.arange(E)
            per = len(order) // ep_size
            rank_load = np.asarray([counts[l, order[r * per:(r + 1) * per]].sum() for r in range(ep_size)], dtype=np.float64)
            out[l] = rank_load.mean() / rank_load.max() if rank_load.max() > 0 else 1.0
        return out

    def dump(self, ep_size: int = 1) -> Dict[str, Any]:
        c = self.load()
        tot = c.sum(1, keepdims=True).clip(min=1)
        d = {"mode": self.mode, "num_passes": self.num_passes, "counts": c.tolist(),
             "imbalance": (c.max(1) / (tot[:, 0] / self.E)).tolist()}
        if ep_size > 1:
            d["balancedness"] = self.balancedness(c, ep_size).tolist()
        if self.mode == "per_pass":
            d["passes"] = [p.tolist() for p in self.passes]
        return d


class EPLBManager:
    """Automatic re-planning (reference eplb/eplb_manager.py, ``--enable-eplb --eplb-rebalance-num-iterations N``): every ``interval`` forward
    passes look at the recent load; if the worst layer's balancedness under the CURRENT placement is below ``threshold``, ask the runner for a
    live re-plan (``ModelRunner.eplb_rebalance``: weights move between ranks, dispatch tables switch) and start a fresh window."""

    def __init__(self, runner: Any, recorder: ExpertDistributionRecorder, ep_size: int, interval: int = 1000, threshold: float = 0.8,
                 redundant_slots_per_rank: int = 0):
        self.runner, self.rec, self.ep = runner, recorder, max(1, ep_size)
        self.interval, self.threshold, self.redundant = max(1, interval), threshold, redundant_slots_per_rank
        self.since = 0
        self.rebalances = 0
        self.last_balancedness: Optional[float] = None
        recorder.enabled = True

    def on_forward_pass(self) -> bool:
        self.rec.end_pass()
        self.since += 1
        if self.since < self.interval:
            return False
        self.since = 0
        load = self.rec.load()
        if load.sum() == 0:
            return False
        placement = getattr(self.runner, "expert_placement", None)       # [layers, physical slots] -> logical expert (eplb/placement.py)
        bal = float(self.rec.balancedness(load, self.ep, placement).min())
        self.last_balancedness = bal
        if bal >= self.threshold:
            return False
        self.runner.eplb_rebalance(self.redundant, load)
        self.rebalances += 1
        self.rec.reset()
        return True


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
    a = getattr(runner, "args", None)
    rec = ExpertDistributionRecorder(len(layers), layers[0].num_experts, mode=getattr(a, "expert_distribution_recorder_mode", "stat") or "stat",
                                     window=getattr(a, "eplb_window", 64) or 64, device=next(layers[0].parameters()).device)
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
        rec.reset()
        return ios.GenericResp(True, "recording")
    if msg.action == "stop":
        rec.enabled = False
        return ios.GenericResp(True, "stopped")
    return ios.GenericResp(True, "dump", rec.dump(getattr(getattr(runner, "ep", None), "world_size", 1) or 1))
