"""KV bookkeeping shared by the scheduler: host page table + allocator + radix tree.

Stands in for the reference's ``mem_cache/allocation.py`` (``alloc_for_extend:281``,
``alloc_for_decode:512``), ``mem_cache/common.py`` (``release_kv_cache:198``,
``evict_from_tree_cache:114``) and the ``req_to_token`` table (``memory_pool.py:256``).

The host page table ``page_table[req_slot, i]`` (int32) is the single source of truth for which page
holds tokens ``[i*ps, (i+1)*ps)`` of a request.  Every mutation is recorded as a (row, col, val) delta;
the model runner ships the deltas inside the per-step packed metadata buffer and a tiny CUDA kernel
applies them to the device copy, so the device table is never rebuilt.
"""
from __future__ import annotations

from typing import Any, List, Optional, Tuple

import numpy as np

from .allocator import PagedAllocator, ReqSlotPool


class KVManager:
    def __init__(self, allocator: PagedAllocator, tree: Any, page_size: int, max_reqs: int, max_context_len: int):
        self.allocator = allocator
        self.tree = tree
        self.ps = page_size
        self.max_reqs = max_reqs
        self.max_pages_per_req = (max_context_len + page_size - 1) // page_size + 1
        self.slots = ReqSlotPool(max_reqs)
        self.page_table = np.zeros((max_reqs + 1, self.max_pages_per_req), dtype=np.int32)
        self._delta_rows: List[np.ndarray] = []
        self._delta_cols: List[np.ndarray] = []
        self._delta_vals: List[np.ndarray] = []

    # ------------------------------------------------------------------ accounting
    def available_pages(self) -> int:
        return self.allocator.available_pages() + self.tree.evictable_pages()

    def total_pages(self) -> int:
        return self.allocator.total_pages

    def token_usage(self) -> float:
        return 1.0 - self.available_pages() / max(1, self.total_pages())

    def _alloc(self, n: int) -> Optional[np.ndarray]:
        short = n - self.allocator.available_pages()
        if short > 0:
            self.tree.evict(short)
        return self.allocator.alloc(n)

    def _record(self, rows, cols, vals) -> None:
        self._delta_rows.append(np.asarray(rows, dtype=np.int32).reshape(-1))
        self._delta_cols.append(np.asarray(cols, dtype=np.int32).reshape(-1))
        self._delta_vals.append(np.asarray(vals, dtype=np.int32).reshape(-1))

    def take_deltas(self) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        if not self._delta_rows:
            e = np.empty(0, dtype=np.int32)
            return e, e, e
        out = (np.concatenate(self._delta_rows), np.concatenate(self._delta_cols), np.concatenate(self._delta_vals))
        self._delta_rows, self._delta_cols, self._delta_vals = [], [], []
        return out

    def req_pages(self, req) -> np.ndarray:
        return self.page_table[req.req_slot, : req.num_pages]

    # ------------------------------------------------------------------ admission
    def match_prefix(self, req) -> int:
        """Peek the cached prefix length (tokens) without locking."""
        ids = req.all_ids()
        limit = max(0, len(ids) - 1)
        if getattr(req, "return_logprob", False) and getattr(req, "logprob_start_len", -1) >= 0:
            limit = min(limit, req.logprob_start_len)   # prompt logprobs need those positions recomputed
        pages, node = self.tree.match_prefix(ids[:limit], req.extra_key)
        req._match = (pages, node)
        return int(pages.size) * self.ps

    def admit(self, req) -> bool:
        """Give the request a slot, install + lock its cached prefix.  Returns False if no slot is free."""
        slot = self.slots.alloc()
        if slot is None:
            return False
        pages, node = getattr(req, "_match", None) or (None, None)
        if pages is None:
            self.match_prefix(req)
            pages, node = req._match
        req._match = None
        req.req_slot = slot
        n = int(pages.size)
        req.num_pages = n
        req.num_cached_pages = n
        req.last_node = node
        req.prefix_len = n * self.ps
        req.kv_len = req.prefix_len
        req.cached_tokens += req.prefix_len
        self.tree.inc_lock_ref(node)
        if n:
            self.page_table[slot, :n] = pages
            self._record(np.full(n, slot), np.arange(n), pages)
        return True

    # ------------------------------------------------------------------ extend / decode allocation
    def alloc_for_extend(self, req, fill_len: int) -> bool:
        """make sure pages exist for tokens [0, fill_len)"""
        need = (fill_len + self.ps - 1) // self.ps - req.num_pages
        if need > 0:
            if req.num_pages + need > self.max_pages_per_req:
                return False
            pages = self._alloc(need)
            if pages is None:
                return False
            cols = np.arange(req.num_pages, req.num_pages + need)
            self.page_table[req.req_slot, cols] = pages
            self._record(np.full(need, req.req_slot), cols, pages)
            req.num_pages += need
        return True

    def trim(self, req, n_tokens: int) -> None:
        """give back pages beyond those needed for ``n_tokens`` tokens (speculative decoding over-allocates for drafts)"""
        need = max((n_tokens + self.ps - 1) // self.ps, req.num_cached_pages)
        if req.num_pages > need:
            extra = self.page_table[req.req_slot, need:req.num_pages].copy()
            self.allocator.free(extra)
            req.num_pages = need

    def slots_for(self, req, start: int, end: int) -> np.ndarray:
        pos = np.arange(start, end, dtype=np.int64)
        return (self.page_table[req.req_slot, pos // self.ps].astype(np.int64) * self.ps + pos % self.ps).astype(np.int32)

    def alloc_for_decode(self, reqs: List[Any], req_slots: np.ndarray, positions: np.ndarray) -> Optional[np.ndarray]:
        """positions[i] = index of the token being written for request i.  Returns slot ids or None on OOM
        (nothing is allocated in that case)."""
        ps = self.ps
        cols = positions // ps
        need = np.flatnonzero((positions % ps) == 0)
        if need.size:
            if (cols[need] >= self.max_pages_per_req).any():
                return None
            pages = self._alloc(int(need.size))
            if pages is None:
                return None
            self.page_table[req_slots[need], cols[need]] = pages
            self._record(req_slots[need], cols[need], pages)
            for i in need.tolist():
                reqs[i].num_pages += 1
        return (self.page_table[req_slots, cols].astype(np.int64) * ps + positions % ps).astype(np.int32)

    def pages_needed_for_decode(self, positions: np.ndarray) -> int:
        return int(((positions % self.ps) == 0).sum())

    # ------------------------------------------------------------------ caching / release
    def cache_unfinished(self, req) -> None:
        """After a prefill chunk: publish the page-aligned computed prefix to the tree and swap the
        request onto the tree's pages (reference radix_cache.py:516)."""
        ps = self.ps
        ids = req.all_ids()
        n_al = min(req.kv_len, len(ids)) // ps
        if n_al <= req.num_cached_pages or getattr(self.tree, "disable", False):
            return
        own = self.page_table[req.req_slot, :n_al].copy()
        matched = self.tree.insert(ids[: n_al * ps], own, req.extra_key, req.priority)
        # duplicates: pages in [num_cached_pages, matched) were already in the tree -> free ours
        if matched > req.num_cached_pages:
            self.allocator.free(own[req.num_cached_pages:matched])
        pages, node = self.tree.match_prefix(ids[: n_al * ps], req.extra_key)
        assert pages.size == n_al, (pages.size, n_al)
        self.tree.inc_lock_ref(node)
        self.tree.dec_lock_ref(req.last_node)
        req.last_node = node
        changed = np.flatnonzero(pages != own)
        if changed.size:
            self.page_table[req.req_slot, changed] = pages[changed]
            self._record(np.full(changed.size, req.req_slot), changed, pages[changed])
        req.num_cached_pages = n_al

    def release(self, req, insert: bool = True) -> None:
        """Request leaves the engine (finished / aborted / retracted): hand its computed pages to the tree
        (page-aligned part) and free the rest (reference radix_cache.py:459 cache_finished_req)."""
        if req.req_slot is None:
            return
        ps = self.ps
        own = self.page_table[req.req_slot, : req.num_pages].copy()
        ids = req.all_ids()
        n_al = min(req.kv_len, len(ids)) // ps if insert and not getattr(self.tree, "disable", False) else req.num_cached_pages
        n_al = max(n_al, req.num_cached_pages)
        if n_al > req.num_cached_pages:
            matched = self.tree.insert(ids[: n_al * ps], own[:n_al], req.extra_key, req.priority)
            if matched > req.num_cached_pages:
                self.allocator.free(own[req.num_cached_pages:matched])
        if req.num_pages > n_al:
            self.allocator.free(own[n_al:])
        if req.last_node is not None:
            self.tree.dec_lock_ref(req.last_node)
        self.slots.free(req.req_slot)
        req.req_slot = None
        req.num_pages = 0
        req.num_cached_pages = 0
        req.last_node = None

    # ------------------------------------------------------------------ invariants (reference invariant_checker.py:65-453)
    def check_invariants(self, live_reqs: List[Any]) -> None:
        tree_pages = self.tree.all_pages()
        owned = []
        for r in live_reqs:
            if r.req_slot is not None:
                owned.append(self.page_table[r.req_slot, r.num_cached_pages: r.num_pages])
        owned_arr = np.concatenate(owned) if owned else np.empty(0, dtype=np.int32)
        used = np.concatenate([tree_pages, owned_arr])
        if np.unique(used).size != used.size:
            raise AssertionError("KV page owned twice (tree/request aliasing)")
        total = self.allocator.available_pages() + used.size
        if total != self.allocator.total_pages:
            raise AssertionError(f"KV page leak: free {self.allocator.available_pages()} + used {used.size} != total {self.allocator.total_pages}")
        if tree_pages.size != self.tree.evictable_pages() + self.tree.protected_pages():
            raise AssertionError("radix accounting mismatch")
