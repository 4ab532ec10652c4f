"""Prefix caching for state-space / hybrid models: recurrent-state snapshots attached to prefix-tree nodes.

Parity: reference ``mem_cache/mamba_radix_cache.py`` (a match is only usable where a Mamba state checkpoint exists), ``memory_pool.py``
``MambaPool`` / ``mamba_checkpoint_pool.py``, ``--mamba-track-interval``-style boundary tracking.

A transformer's prefix is reusable at any page boundary; a recurrent layer's is reusable only where its STATE was saved.  Here:

* the ``MambaPool`` gets ``--mamba-cache-slots`` extra rows; a snapshot = (conv window, SSM state) of every recurrent layer in one row;
* snapshots are taken INSIDE the prefill forward at a chunk-scan boundary: the SSD scan already materialises the state at every 64-token
  boundary, so the mixer just copies the one at the largest boundary that is also page-aligned (``ForwardBatch.mamba_snap``) — no extra pass,
  no split of the prompt into two forwards;
* ``MambaRadixCache`` nodes carry ``mamba_slot``; ``match`` walks up from the longest token match to the deepest node that has one and
  truncates the match there; admission copies the snapshot row into the request's own state row; snapshot rows are recycled LRU.
"""
from __future__ import annotations

from math import gcd
from typing import Any, Callable, List, Optional, Tuple

import numpy as np

from .allocator import PagedAllocator
from .kv_manager import KVManager
from .radix_cache import RadixCache, TreeNode

_EMPTY = np.empty(0, dtype=np.int32)


class MambaRadixCache(RadixCache):
    def __init__(self, allocator: Optional[PagedAllocator], snapshot_rows: List[int], page_size: int = 1, eviction_policy: str = "lru",
                 disable: bool = False, **kw: Any):
        self.free_rows: List[int] = list(snapshot_rows)
        self.num_rows = len(snapshot_rows)
        self.snap_nodes: List[TreeNode] = []
        self.snapshots_taken = self.snapshots_recycled = self.snapshot_hits = 0
        super().__init__(allocator, page_size, eviction_policy, disable, **kw)

    def reset(self) -> None:
        super().reset()
        for n in getattr(self, "snap_nodes", []):
            if n.mamba_slot is not None:
                self.free_rows.append(n.mamba_slot)
                n.mamba_slot = None
        self.snap_nodes = []

    # ---- snapshot rows
    def alloc_row(self) -> Optional[int]:
        if self.free_rows:
            return self.free_rows.pop()
        live = [n for n in self.snap_nodes if n.mamba_slot is not None]
        if not live:
            return None
        victim = min(live, key=lambda n: (n.last_access, n.id))      # a snapshot is only read at admission: recycling never hurts a running request
        row, victim.mamba_slot = victim.mamba_slot, None
        self.snap_nodes.remove(victim)
        self.snapshots_recycled += 1
        return row

    def release_row(self, row: int) -> None:
        self.free_rows.append(row)

    def attach(self, token_ids, n_pages: int, row: int, extra_key: Any = None) -> bool:
        """attach snapshot ``row`` to the node that ends exactly after ``n_pages`` pages of ``token_ids`` (splitting a node if needed)"""
        key = np.asarray(token_ids, dtype=np.int32)[: n_pages * self.page_size]
        pages, node = self.match_prefix(key, extra_key)
        if pages.size != n_pages or node.parent is None:
            return False
        if node.mamba_slot is not None:            # somebody else snapshotted the same prefix first
            return False
        node.mamba_slot = row
        self.snap_nodes.append(node)
        self.snapshots_taken += 1
        return True

    # ---- match: deepest snapshot on the path of the longest token match
    def match_prefix_mamba(self, token_ids, extra_key: Any = None) -> Tuple[np.ndarray, TreeNode, Optional[int]]:
        pages, node = self.match_prefix(token_ids, extra_key)
        n = int(pages.size)
        while node is not None and node.parent is not None and node.mamba_slot is None:
            n -= int(node.key.size // self.page_size)
            node = node.parent
        if node is None or node.parent is None or n <= 0:
            return _EMPTY, self._root(extra_key), None
        node.last_access = self._now()
        self.snapshot_hits += 1
        return pages[:n], node, node.mamba_slot

    def _split(self, child: TreeNode, m_pages: int) -> TreeNode:
        new = super()._split(child, m_pages)
        new.mamba_slot = None                      # the snapshot describes the END of the original node: it stays with the tail
        return new

    def evict(self, num_pages: int, on_evict=None) -> int:
        def hook(n: TreeNode) -> bool:
            if n.mamba_slot is not None:
                self.free_rows.append(n.mamba_slot)
                n.mamba_slot = None
                if n in self.snap_nodes:
                    self.snap_nodes.remove(n)
            return bool(on_evict(n)) if on_evict is not None else False

        return super().evict(num_pages, hook)


class MambaKVManager(KVManager):
    """KVManager for models with recurrent layers: prefix matches end at snapshots, admissions restore them, prefill steps plan new ones."""

    def __init__(self, allocator: PagedAllocator, tree: MambaRadixCache, page_size: int, max_reqs: int, max_context_len: int,
                 granularity: int, copy_state: Callable[[int, int], None]):
        super().__init__(allocator, tree, page_size, max_reqs, max_context_len)
        self.step_tokens = granularity * page_size // gcd(granularity, page_size)     # a snapshot boundary is a scan boundary AND a page boundary
        self.granularity = granularity
        self.copy_state = copy_state

    def match_prefix(self, req) -> int:
        ids = req.all_ids()
        limit = max(0, len(ids) - 1)
        if getattr(req, "return_logprob", False) and getattr(req, "logprob_start_len", -1) >= 0:
            limit = min(limit, req.logprob_start_len)
        pages, node, row = self.tree.match_prefix_mamba(ids[:limit], req.extra_key)
        req._match = (pages, node)
        req._mamba_row = row
        return int(pages.size) * self.ps

    def admit(self, req) -> bool:
        if getattr(req, "_match", None) is None:
            self.match_prefix(req)
        row = getattr(req, "_mamba_row", None)
        if not super().admit(req):
            return False
        req._mamba_row = None
        req._snap = None
        if req.prefix_len > 0:
            assert row is not None
            self.copy_state(row, req.req_slot)     # stream-ordered before the request's first forward
        return True

    def plan_snapshot(self, req) -> Optional[Tuple[int, int]]:
        """called once the chunk [prefix_len, fill_len) of ``req`` is fixed: (boundary relative to the chunk start, snapshot row) or None"""
        if getattr(self.tree, "disable", False):
            return None
        L = self.step_tokens
        bnd = (req.fill_len // L) * L
        rel = bnd - req.prefix_len
        if rel <= 0 or rel % self.granularity:
            return None
        row = self.tree.alloc_row()
        if row is None:
            return None
        req._snap = (bnd, row)
        return rel, row

    def _commit_snapshot(self, req) -> None:
        snap = getattr(req, "_snap", None)
        if snap is None:
            return
        req._snap = None
        bnd, row = snap
        if not self.tree.attach(req.all_ids(), bnd // self.ps, row, req.extra_key):
            self.tree.release_row(row)

    def cache_unfinished(self, req) -> None:
        super().cache_unfinished(req)
        self._commit_snapshot(req)

    def release(self, req, insert: bool = True) -> None:
        ids, key, snap = req.all_ids(), req.extra_key, getattr(req, "_snap", None)
        super().release(req, insert)
        if snap is not None:
            req._snap = None
            if not (insert and self.tree.attach(ids, snap[0] // self.ps, snap[1], key)):
                self.tree.release_row(snap[1])
