"""Waiting-queue ordering + token-budget admission.

Parity: reference ``python/sglang/srt/managers/schedule_policy.py`` (``CacheAwarePolicy:200`` lpm /
dfs-weight, ``CacheAgnosticPolicy:207`` fcfs / lof / random, priority scheduling ``calc_priority:237``,
``PrefillAdder:511``, ``add_chunked_req:1004``, ``add_one_req:1208``).  Budgets are kept in *pages*.
"""
from __future__ import annotations

import random
from collections import defaultdict
from typing import Any, Dict, List, Optional

from .schedule_batch import Req

CACHE_AWARE = ("lpm", "dfs-weight")
# LPM degenerates to O(n * match) per iteration; the reference falls back to fcfs above this queue length
LPM_FALLBACK_QUEUE_LEN = 128


class SchedulePolicy:
    def __init__(self, policy: str, kv: Any, enable_priority: bool = False, seed: int = 0):
        self.policy = policy
        self.kv = kv
        self.enable_priority = enable_priority
        self._rng = random.Random(seed)

    def calc_priority(self, waiting: List[Req]) -> None:
        policy = self.policy
        if policy == "lpm" and len(waiting) > LPM_FALLBACK_QUEUE_LEN:
            policy = "fcfs"
        if policy in CACHE_AWARE:
            for r in waiting:
                r._prefix_hint = self.kv.match_prefix(r)
        if policy == "lpm":
            waiting.sort(key=lambda r: -r._prefix_hint)
        elif policy == "dfs-weight":
            self._dfs_weight(waiting)
        elif policy == "fcfs":
            pass  # arrival order
        elif policy == "lof":
            waiting.sort(key=lambda r: -r.sampling_params.max_new_tokens)
        elif policy == "random":
            self._rng.shuffle(waiting)
        if self.enable_priority:
            waiting.sort(key=lambda r: -r.priority)  # stable: keeps policy order inside a priority class

    def _dfs_weight(self, waiting: List[Req]) -> None:
        """group requests by their matched tree node, visit heavier subtrees first"""
        groups: Dict[int, List[Req]] = defaultdict(list)
        for r in waiting:
            node = r._match[1] if getattr(r, "_match", None) else None
            groups[id(node)].append(r)
        ordered: List[Req] = []
        for _, rs in sorted(groups.items(), key=lambda kv_: -len(kv_[1])):
            ordered.extend(sorted(rs, key=lambda r: -r._prefix_hint))
        waiting[:] = ordered


class AddResult:
    CONTINUE = 0
    NO_TOKEN = 1   # out of KV budget
    OTHER = 2      # out of per-batch budget / no request slot


class PrefillAdder:
    def __init__(self, kv: Any, page_size: int, new_token_ratio: float, max_prefill_tokens: int,
                 chunked_prefill_size: Optional[int], running_reqs: List[Req], max_running: int,
                 clip_max_new_tokens: int = 4096):
        self.kv, self.ps = kv, page_size
        self.new_token_ratio = new_token_ratio
        self.rem_input_tokens = max_prefill_tokens
        self.rem_chunk_tokens = chunked_prefill_size
        self.clip = clip_max_new_tokens
        self.can_run: List[Req] = []
        self.new_chunked_req: Optional[Req] = None
        self.max_running = max_running
        self.num_running = len(running_reqs)
        # pages that must stay free for the decode of already-running requests
        self.reserved_pages = 0
        for r in running_reqs:
            self.reserved_pages += self._decode_reserve(r)
        self.used_pages = 0

    def _decode_reserve(self, r: Req) -> int:
        remain = max(0, min(r.sampling_params.max_new_tokens - len(r.output_ids), self.clip))
        return int(remain * self.new_token_ratio + self.ps - 1) // self.ps

    def rem_pages(self) -> int:
        avail = self.kv.available_pages()
        if hasattr(self.kv, "swa_available_pages"):        # hybrid sliding-window models: every new token needs a page in BOTH pools
            avail = min(avail, self.kv.swa_available_pages())
        return avail - self.reserved_pages - self.used_pages

    def _fits(self, new_pages: int, decode_reserve: int) -> bool:
        return new_pages + decode_reserve <= self.rem_pages()

    def add_chunked_req(self, req: Req) -> bool:
        """continue an in-flight chunked request (reference add_chunked_req:1004)"""
        total = req.seqlen
        remaining = total - req.kv_len
        budget = self.rem_chunk_tokens if self.rem_chunk_tokens is not None else remaining
        take = min(remaining, budget)
        truncated = take < remaining
        if truncated:
            take -= take % self.ps
            if take <= 0:
                return False
        req.prefix_len = req.kv_len
        req.extend_len = take
        req.fill_len = req.kv_len + take
        new_pages = (req.fill_len + self.ps - 1) // self.ps - req.num_pages
        self.used_pages += max(0, new_pages)
        self.rem_input_tokens -= take
        if self.rem_chunk_tokens is not None:
            self.rem_chunk_tokens -= take
        self.can_run.append(req)
        if truncated:
            req.is_chunked += 1
            self.new_chunked_req = req
        return True

    def add_one_req(self, req: Req) -> int:
        if self.num_running + len(self.can_run) >= self.max_running:
            return AddResult.OTHER
        if self.kv.slots.available() == 0:
            return AddResult.OTHER
        total = req.seqlen
        prefix = self.kv.match_prefix(req)
        extend = total - prefix
        if self.rem_chunk_tokens is not None and self.rem_chunk_tokens <= 0:
            return AddResult.OTHER
        if self.rem_input_tokens <= 0 and self.can_run:
            return AddResult.OTHER
        take = extend
        truncated = False
        if self.rem_chunk_tokens is not None and extend > self.rem_chunk_tokens:
            take = self.rem_chunk_tokens - (prefix + self.rem_chunk_tokens) % self.ps if self.ps > 1 else self.rem_chunk_tokens
            take = min(take, self.rem_chunk_tokens)
            if take <= 0:
                return AddResult.OTHER
            truncated = True
        elif self.rem_chunk_tokens is None and extend > self.rem_input_tokens and self.can_run:
            return AddResult.OTHER
        fill = prefix + take
        new_pages = (fill + self.ps - 1) // self.ps - prefix // self.ps
        reserve = 0 if truncated else int(min(req.sampling_params.max_new_tokens - len(req.output_ids), self.clip)
                                           * self.new_token_ratio + self.ps - 1) // self.ps
        # locking the prefix moves its pages from evictable to protected: account for it
        lock_cost = self._lock_cost(req)
        if not self._fits(new_pages + lock_cost, reserve):
            return AddResult.NO_TOKEN
        if not self.kv.admit(req):
            return AddResult.OTHER
        req.extend_len = take
        req.fill_len = fill
        self.used_pages += new_pages
        self.reserved_pages += reserve
        self.rem_input_tokens -= take
        if self.rem_chunk_tokens is not None:
            self.rem_chunk_tokens -= take
        self.can_run.append(req)
        if truncated:
            req.is_chunked += 1
            self.new_chunked_req = req
        return AddResult.CONTINUE

    def _lock_cost(self, req: Req) -> int:
        """evictable pages that become protected when this request's prefix is locked"""
        m = getattr(req, "_match", None)
        return self.kv.tree.lock_cost(m[1]) if m else 0
