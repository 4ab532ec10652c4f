"""SPMD divergence detectors.

* ``rank_consensus(group)`` — decorator: every rank of ``group`` must return the *same* value (the SPMD scheduler's batch decisions, sampled
  token ids, ...).  The value is hashed and min/max-reduced over the Gloo twin of the group; a mismatch raises (or logs) naming the function.
  Parity: reference ``python/sglang/srt/utils/rank_consensus_checker.py:28,397``.  Off unless ``SGLB_RANK_CONSENSUS`` is ``log`` / ``crash``.
* ``SlowRankDetector`` — ranks exchange their recent step times every N steps; a rank that is consistently slower than the group median by
  ``ratio`` is reported (thermal throttling, a noisy neighbour, a failing link).  Parity: ``utils/slow_rank_detector.py``.
"""
from __future__ import annotations

import functools
import hashlib
import logging
import os
import pickle
import time
from typing import Any, Callable, List, Optional

import numpy as np
import torch

logger = logging.getLogger(__name__)


class RankDivergence(RuntimeError):
    pass


def digest(obj: Any) -> int:
    def norm(o):
        if isinstance(o, torch.Tensor):
            return ("T", tuple(o.shape), o.detach().cpu().numpy().tobytes())
        if isinstance(o, np.ndarray):
            return ("A", o.shape, o.tobytes())
        if isinstance(o, (list, tuple)):
            return tuple(norm(x) for x in o)
        if isinstance(o, dict):
            return tuple(sorted((k, norm(v)) for k, v in o.items()))
        return o
    return int.from_bytes(hashlib.blake2b(pickle.dumps(norm(obj)), digest_size=7).digest(), "little")


def check_consensus(group: Any, value: Any, what: str, mode: Optional[str] = None) -> bool:
    """True when all ranks agree.  ``group``: a GroupCoordinator (its ``cpu_group`` carries the reduction)."""
    mode = mode or os.environ.get("SGLB_RANK_CONSENSUS", "off")
    if mode == "off" or group is None or group.world_size <= 1:
        return True
    import torch.distributed as dist

    h = digest(value)
    t = torch.tensor([h, -h], dtype=torch.int64)
    dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group.cpu_group)
    lo, hi = int(t[0]), -int(t[1])
    if lo == hi:
        return True
    msg = f"rank divergence in {what}: rank {group.rank} digest {h:#x}, group min {lo:#x} max {hi:#x}"
    if mode == "crash":
        raise RankDivergence(msg)
    logger.error(msg)
    return False


def rank_consensus(group_getter: Callable[[], Any], project: Optional[Callable[[Any], Any]] = None):
    def deco(fn):
        @functools.wraps(fn)
        def wrapper(*a, **kw):
            out = fn(*a, **kw)
            if os.environ.get("SGLB_RANK_CONSENSUS", "off") != "off":
                check_consensus(group_getter(), project(out) if project else out, fn.__qualname__)
            return out
        return wrapper
    return deco


class SlowRankDetector:
    def __init__(self, group: Any, window: int = 64, every: int = 256, ratio: float = 1.3):
        self.group, self.window, self.every, self.ratio = group, window, every, ratio
        self.times: List[float] = []
        self.steps = 0
        self.reports: List[dict] = []
        self._t0: Optional[float] = None

    def step_begin(self) -> None:
        self._t0 = time.perf_counter()

    def step_end(self) -> Optional[dict]:
        if self._t0 is None:
            return None
        self.times.append(time.perf_counter() - self._t0)
        self.times = self.times[-self.window:]
        self.steps += 1
        if self.steps % self.every:
            return None
        return self.exchange()

    def exchange(self) -> Optional[dict]:
        mine = float(np.median(self.times)) if self.times else 0.0
        allm = self.group.all_gather_pyobj(mine) if self.group is not None and self.group.world_size > 1 else [mine]
        med = float(np.median(allm))
        slow = [i for i, v in enumerate(allm) if med > 0 and v > self.ratio * med]
        if not slow:
            return None
        rep = {"step": self.steps, "median_s": med, "slow_ranks": slow, "times_s": allm}
        self.reports.append(rep)
        logger.warning("slow ranks %s: step time %s vs group median %.4fs", slow, [round(allm[i], 4) for i in slow], med)
        return rep
