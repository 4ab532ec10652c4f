"""Prefill delayers: trade a bounded amount of TTFT for larger, less fragmented prefill batches.

Parity: reference ``python/sglang/srt/managers/prefill_delayer.py`` (``PrefillDelayer:73``, negotiation ``:171-352``,
``RecentPrefillBatchSizeTracker:22``) and ``min_free_slots_delayer.py`` (``resolve_min_free_slots:4``, ``MinFreeSlotsDelayer:26``).

Why it exists: under DP-attention every forward is a lock-step collective over all DP groups, so one group that starts a prefill
stalls the decode of every other group.  The delayer lets the groups *agree* when to prefill:

* all groups have something to prefill -> go now, unless the running batch is so full that the new requests would not fit next to
  it (slot trigger) or the waiting queue is still short compared with the running batch (queue trigger, opt-in); then wait;
* only some groups have something -> wait (bounded by ``max_delay_passes``) for the others to catch up;
* a group whose KV pool is nearly empty forces admission (token-usage low watermark): an idle GPU never waits.

Each scheduler pass calls ``negotiate`` exactly once (also when it has nothing to prefill) so the gather stays matched.  The
gather is a caller-supplied function (Gloo object all-gather over the TP group's CPU twin in a real job, identity in a single
process), which keeps the policy a pure function of the gathered table and trivially unit-testable.
"""
from __future__ import annotations

import dataclasses
import time
from collections import deque
from typing import Callable, Deque, List, NamedTuple, Optional, Sequence


class RecentPrefillBatchSizeTracker:
    """high-watermark of the last ``window`` non-empty prefill attempts"""

    def __init__(self, window: int = 16):
        assert window > 0
        self._sizes: Deque[int] = deque(maxlen=window)

    @property
    def max_prefill_bs(self) -> int:
        return max(self._sizes, default=0)

    def observe(self, n: int) -> int:
        if n > 0:
            self._sizes.append(n)
        return self.max_prefill_bs


@dataclasses.dataclass(frozen=True)
class _Wait:
    passes: int = 0
    t0: float = dataclasses.field(default_factory=time.perf_counter)


class Decision(NamedTuple):
    allow: bool
    reason: str          # "" | no_wait | wait_success | wait_timeout | token_watermark | delay
    status: str          # all | none | mixed
    waited_passes: int = 0
    waited_s: float = 0.0


def resolve_min_free_slots(user_value: Optional[int], max_running: int) -> Optional[int]:
    if user_value is None:
        return None
    v = min(int(user_value), max(0, int(max_running)))
    return v if v > 1 else None


class MinFreeSlotsDelayer:
    """rank-local: hold new admissions until ``min_free_slots`` running slots are free, so they enter as one batch"""

    def __init__(self, min_free_slots: int):
        self.min_free_slots = min_free_slots

    def should_delay(self, running_bs: int, free_slots: int) -> bool:
        return running_bs > 0 and free_slots < self.min_free_slots


Row = Sequence[int]     # (prefillable, force_allow, running_bs, max_prefill_bs, waiting_len)


class PrefillDelayer:
    def __init__(self, gather: Optional[Callable[[List[int]], List[Row]]] = None, dp_size: int = 1, dp_attention: bool = False,
                 max_delay_passes: int = 30, token_usage_low_watermark: Optional[float] = None,
                 queue_min_ratio: Optional[float] = None, max_delay_ms: Optional[float] = None,
                 prefill_max_requests: Optional[int] = None):
        self.gather = gather or (lambda row: [row])
        self.dp_size = dp_size
        self.dp_attention = dp_attention
        self.max_delay_passes = max_delay_passes
        self.low_watermark = token_usage_low_watermark
        self.queue_min_ratio = queue_min_ratio
        self.max_delay_ms = 5000.0 if max_delay_ms is None else max_delay_ms
        self.prefill_max_requests = prefill_max_requests
        self.tracker = RecentPrefillBatchSizeTracker()
        self._wait: Optional[_Wait] = None
        self._skip_first = True
        self.stats = {"delay": 0, "no_wait": 0, "wait_success": 0, "wait_timeout": 0, "token_watermark": 0}

    # ---- the policy: a pure function of (previous wait, gathered table)
    def decide(self, prev: Optional[_Wait], table: List[Row], max_running: int):
        pre = [int(r[0]) for r in table]
        force = any(int(r[1]) for r in table)
        status = "all" if min(pre) > 0 else ("none" if max(pre) == 0 else "mixed")
        waited = dict(waited_passes=prev.passes if prev else 0, waited_s=(time.perf_counter() - prev.t0) if prev else 0.0)

        def delay_or_timeout():
            if (prev.passes if prev else 0) < self.max_delay_passes - 1:
                w = prev or _Wait()
                return dataclasses.replace(w, passes=w.passes + 1), Decision(False, "delay", status)
            return None, Decision(True, "wait_timeout", status, **waited)

        if status == "none":
            return None, Decision(True, "", status, **waited)
        if force:
            return None, Decision(True, "token_watermark", status, **waited)
        if status == "mixed":
            return delay_or_timeout()
        # status == "all"
        if not self.dp_attention:
            max_running = (max_running + self.dp_size - 1) // self.dp_size
        run_max = max(int(r[2]) for r in table)
        pbs_max = max(int(r[3]) for r in table)
        wait_max = max(int(r[4]) for r in table)
        queue_cond = False
        if self.queue_min_ratio is not None and run_max > 0:
            cap = self.prefill_max_requests if self.prefill_max_requests is not None else pbs_max
            need = min(int(run_max * self.queue_min_ratio), cap)
            queue_cond = need > 0 and wait_max < need
            if queue_cond and prev is not None and (time.perf_counter() - prev.t0) * 1e3 >= self.max_delay_ms:
                queue_cond = False
        slot_cond = max_running - run_max < pbs_max
        if slot_cond or queue_cond:
            if self._skip_first:        # the very first merge would otherwise keep decode from ever reaching max batch
                self._skip_first = False
            else:
                return delay_or_timeout()
        return None, Decision(True, "wait_success" if prev is not None else "no_wait", status, **waited)

    # ---- one call per scheduler pass on every rank
    def negotiate(self, local_prefillable: bool, token_usage: float, running_bs: int, waiting_len: int, max_running: int) -> Decision:
        force = bool(local_prefillable and self.low_watermark is not None and token_usage < self.low_watermark)
        est = self.tracker.max_prefill_bs or (min(waiting_len, max(1, max_running - running_bs)) if local_prefillable else 0)
        table = self.gather([int(local_prefillable), int(force), int(running_bs), int(est), int(waiting_len)])
        self._wait, d = self.decide(self._wait, table, max_running)
        if d.reason:
            self.stats[d.reason] = self.stats.get(d.reason, 0) + 1
        return d

    def observe_prefill(self, actual_bs: int) -> None:
        self.tracker.observe(actual_bs)
