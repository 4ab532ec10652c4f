"""Sliding-window KV management: a second, smaller pool for the sliding-window layers of hybrid models.

Parity: reference ``mem_cache/swa_memory_pool.py`` (two token pools behind one allocator), ``allocator/swa.py`` and ``swa_radix_cache.py``
(tombstoned sliding-window KV inside the prefix tree, a match is valid only where the trailing window is intact), ``--swa-full-tokens-ratio``.

Models such as Gemma-2/3, GPT-OSS, Cohere2, Mistral (window) or Llama-4 (chunked) interleave full-attention layers with layers that only ever
read the last ``W`` tokens.  Keeping those layers' KV for the whole context wastes most of the cache (Gemma-3: 5 of 6 layers).  Here:

* ``HybridKVPool`` holds two ``KVPool`` planes: ``full`` (full-attention layers, sized for the whole context) and ``swa`` (sliding layers,
  sized for a window per running request plus reusable prefixes).  ``k(layer) / v(layer)`` route by layer.
* every request has TWO page-table rows (same column = same tokens): the host tables live in ``SWAKVManager``, the device copies are patched
  by the same delta kernel.  After every step ``slide()`` returns the request's own sliding pages that fell out of the window.
* the prefix tree keeps, per node, the sliding pages next to the full pages; a sliding page id of 0 is a tombstone (slid out before the
  request finished, or reclaimed under sliding-pool pressure by ``evict_swa`` without touching the full pages).  ``match_prefix`` returns the
  longest prefix whose trailing window of sliding pages is intact: full-attention layers reuse everything, sliding layers exactly what they
  can still see.
"""
from __future__ import annotations

from typing import Any, List, Optional, Tuple

import numpy as np

from .allocator import PagedAllocator
from .kv_manager import KVManager
from .radix_cache import RadixCache, TreeNode

_EMPTY = np.empty(0, dtype=np.int32)


class SWARadixCache(RadixCache):
    """Python radix tree whose nodes carry sliding-window pages next to the full-attention pages."""

    def __init__(self, allocator: Optional[PagedAllocator], swa_allocator: Optional[PagedAllocator], window: int, page_size: int = 1,
                 eviction_policy: str = "lru", disable: bool = False, **kw: Any):
        self.swa_allocator = swa_allocator
        self.window = int(window)
        super().__init__(allocator, page_size, eviction_policy, disable, **kw)
        # pages of the trailing window a reusable prefix must still hold: the window itself plus the partially filled page in front of it
        self.need_pages = (self.window + page_size - 1) // page_size + 1
        self.swa_tombstoned = 0

    # ---- accounting
    def swa_evictable_pages(self) -> int:
        n, stack = 0, list(self.roots.values())
        while stack:
            x = stack.pop()
            if x.parent is not None and x.lock_ref == 0 and x.swa_pages is not None:
                n += int((x.swa_pages > 0).sum())
            stack.extend(x.children.values())
        return n

    def all_swa_pages(self) -> np.ndarray:
        out, stack = [], list(self.roots.values())
        while stack:
            x = stack.pop()
            if x.parent is not None and x.swa_pages is not None:
                out.append(x.swa_pages[x.swa_pages > 0])
            stack.extend(x.children.values())
        return np.concatenate(out) if out else _EMPTY

    # ---- match: longest prefix whose trailing window of sliding pages is intact
    def match_prefix_swa(self, token_ids, extra_key: Any = None) -> Tuple[np.ndarray, np.ndarray, TreeNode]:
        key = np.asarray(token_ids, dtype=np.int32)
        pages, node = self.match_prefix(key, extra_key)
        if pages.size == 0:
            return _EMPTY, _EMPTY, node
        swa = self._path_swa(node, pages.size)
        ok = swa > 0
        # run[i] = number of consecutive intact sliding pages ending at page i
        run = np.zeros(ok.size, dtype=np.int64)
        c = 0
        for i, v in enumerate(ok.tolist()):
            c = c + 1 if v else 0
            run[i] = c
        P = 0
        for n in range(int(pages.size), 0, -1):
            if run[n - 1] >= min(n, self.need_pages):
                P = n
                break
        if P == pages.size:
            return pages, swa, node
        if P == 0:
            return _EMPTY, _EMPTY, self._root(extra_key)
        pages2, node2 = self.match_prefix(key[: P * self.page_size], extra_key)       # splits at the boundary, returns the proper last node
        assert pages2.size == P
        return pages2, swa[:P], node2

    def _path_swa(self, node: TreeNode, n_pages: int) -> np.ndarray:
        parts = []
        while node is not None and node.parent is not None:
            k = int(node.pages.size) if node.pages is not None else int(node.key.size // self.page_size)
            parts.append(node.swa_pages if node.swa_pages is not None else np.zeros(k, dtype=np.int32))
            node = node.parent
        out = np.concatenate(parts[::-1]) if parts else _EMPTY
        assert out.size == n_pages, (out.size, n_pages)
        return out

    # ---- insert: full pages as in the base class; sliding pages adopted where the tree has none
    def insert_swa(self, token_ids, pages, swa_pages, extra_key: Any = None, priority: int = 0) -> Tuple[int, np.ndarray]:
        """returns (leading pages already present, adopted[i] for i < that count: the tree took the caller's sliding page i because its own
        was a tombstone).  The caller frees its duplicates: full pages [cached, matched), sliding pages [cached, matched) that were NOT adopted."""
        key = np.asarray(token_ids, dtype=np.int32)
        pages = np.asarray(pages, dtype=np.int32)
        swa_pages = np.asarray(swa_pages, dtype=np.int32)
        ps = self.page_size
        n_pages = min(key.size // ps, pages.size)
        key, pages, swa_pages = key[: n_pages * ps], pages[:n_pages], swa_pages[:n_pages]
        matched = self.insert(key, pages, extra_key, priority)
        adopted = np.zeros(matched, dtype=bool)
        # walk the (now fully present) path and attach / heal the sliding pages
        node, off = self._root(extra_key), 0
        while off < key.size:
            child = node.children.get(self._child_key(key, off))
            assert child is not None
            k = int(child.key.size // ps)
            lo = off // ps
            mine = swa_pages[lo:lo + k]
            if child.swa_pages is None:
                child.swa_pages = np.zeros(k, dtype=np.int32)
            if lo >= matched:                      # freshly created by this insert: the tree owns the caller's sliding pages
                child.swa_pages = mine.copy()
            else:                                  # existed before: heal tombstones with the caller's live pages
                heal = (child.swa_pages == 0) & (mine > 0)
                if heal.any():
                    child.swa_pages = np.where(heal, mine, child.swa_pages)
                    idx = np.flatnonzero(heal) + lo
                    adopted[idx[idx < matched]] = True
            node = child
            off += k * ps
        return matched, adopted

    # ---- eviction
    def evict(self, num_pages: int, on_evict=None) -> int:
        def hook(n: TreeNode) -> bool:
            if n.swa_pages is not None and self.swa_allocator is not None:
                live = n.swa_pages[n.swa_pages > 0]
                if live.size:
                    self.swa_allocator.free(live)
                n.swa_pages = None
            return bool(on_evict(n)) if on_evict is not None else False

        return super().evict(num_pages, hook)

    def tombstone_exclusive(self, node: TreeNode, lo: int, hi: int) -> int:
        """free the sliding pages of columns [lo, hi) on the path to ``node`` where the only lock holder is the caller (lock_ref == 1)"""
        path = []
        while node is not None and node.parent is not None:
            path.append(node)
            node = node.parent
        col, freed = 0, 0
        for x in reversed(path):
            k = int(x.key.size // self.page_size)
            a, b = max(lo, col), min(hi, col + k)
            if a < b and x.lock_ref == 1 and x.swa_pages is not None:
                seg = x.swa_pages[a - col:b - col]
                live = seg[seg > 0]
                if live.size:
                    if self.swa_allocator is not None:
                        self.swa_allocator.free(live.copy())
                    x.swa_pages = x.swa_pages.copy()
                    x.swa_pages[a - col:b - col] = 0
                    freed += int(live.size)
                    self.swa_tombstoned += int(live.size)
            col += k
        return freed

    def evict_swa(self, num_pages: int) -> int:
        """reclaim >= num_pages sliding pages from unlocked nodes (policy order) WITHOUT giving up their full-attention pages"""
        if self.disable or num_pages <= 0:
            return 0
        cand, stack = [], list(self.roots.values())
        while stack:
            x = stack.pop()
            if x.parent is not None and x.lock_ref == 0 and x.swa_pages is not None and (x.swa_pages > 0).any():
                cand.append((self._policy(x), x))
            stack.extend(x.children.values())
        cand.sort(key=lambda t: (t[0], t[1].id))
        freed = 0
        for _, x in cand:
            if freed >= num_pages:
                break
            live = x.swa_pages[x.swa_pages > 0]
            if self.swa_allocator is not None:
                self.swa_allocator.free(live)
            freed += int(live.size)
            self.swa_tombstoned += int(live.size)
            x.swa_pages = np.zeros_like(x.swa_pages)
        return freed


class SWAKVManager(KVManager):
    """KVManager with a second (sliding-window) page table and allocator.  Column c of both tables covers the same tokens."""

    def __init__(self, allocator: PagedAllocator, swa_allocator: PagedAllocator, tree: SWARadixCache, page_size: int, max_reqs: int,
                 max_context_len: int, window: int, fetch_granularity: int = 128):
        super().__init__(allocator, tree, page_size, max_reqs, max_context_len)
        self.fetch_granularity = max(int(fetch_granularity), page_size)      # the attention kernels fetch whole key tiles of this many tokens
        self.swa_allocator = swa_allocator
        self.window = int(window)
        self.page_table_swa = np.zeros_like(self.page_table)
        self._swa_rows: List[np.ndarray] = []
        self._swa_cols: List[np.ndarray] = []
        self._swa_vals: List[np.ndarray] = []
        self.swa_slid_pages = 0

    # ---- accounting
    def swa_available_pages(self) -> int:
        return self.swa_allocator.available_pages() + self.tree.swa_evictable_pages()

    def _alloc_swa(self, n: int) -> Optional[np.ndarray]:
        short = n - self.swa_allocator.available_pages()
        if short > 0:
            got = self.tree.evict_swa(short)
            if got < short:                 # tombstoning was not enough: give up whole unlocked leaves (both pools)
                self.tree.evict(short - got)
        return self.swa_allocator.alloc(n)

    def _record_swa(self, rows, cols, vals) -> None:
        self._swa_rows.append(np.asarray(rows, dtype=np.int32).reshape(-1))
        self._swa_cols.append(np.asarray(cols, dtype=np.int32).reshape(-1))
        self._swa_vals.append(np.asarray(vals, dtype=np.int32).reshape(-1))

    def take_swa_deltas(self) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        if not self._swa_rows:
            return _EMPTY, _EMPTY, _EMPTY
        out = (np.concatenate(self._swa_rows), np.concatenate(self._swa_cols), np.concatenate(self._swa_vals))
        self._swa_rows, self._swa_cols, self._swa_vals = [], [], []
        return out

    # ---- admission
    def match_prefix(self, req) -> int:
        ids = req.all_ids()
        limit = max(0, len(ids) - 1)
        if getattr(req, "return_logprob", False) and getattr(req, "logprob_start_len", -1) >= 0:
            limit = min(limit, req.logprob_start_len)
        pages, swa, node = self.tree.match_prefix_swa(ids[:limit], req.extra_key)
        req._match = (pages, node)
        req._match_swa = swa
        return int(pages.size) * self.ps

    def admit(self, req) -> bool:
        if getattr(req, "_match", None) is None:
            self.match_prefix(req)
        swa = req._match_swa
        if not super().admit(req):
            return False
        n = req.num_pages
        req._match_swa = None
        req.swa_lo = 0                      # sliding pages below this column were already given back
        self.page_table_swa[req.req_slot, :] = 0
        if n:
            self.page_table_swa[req.req_slot, :n] = swa[:n]
            self._record_swa(np.full(n, req.req_slot), np.arange(n), swa[:n])
        return True

    # ---- allocation (both pools; nothing is kept on a partial failure)
    def alloc_for_extend(self, req, fill_len: int) -> bool:
        before = req.num_pages
        if not super().alloc_for_extend(req, fill_len):
            return False
        need = req.num_pages - before
        if need > 0:
            pages = self._alloc_swa(need)
            if pages is None:               # roll the full-pool allocation back
                cols = np.arange(before, req.num_pages)
                self.allocator.free(self.page_table[req.req_slot, cols].copy())
                req.num_pages = before
                return False
            cols = np.arange(before, req.num_pages)
            self.page_table_swa[req.req_slot, cols] = pages
            self._record_swa(np.full(need, req.req_slot), cols, pages)
        return True

    def slots_for_swa(self, req, start: int, end: int) -> np.ndarray:
        pos = np.arange(start, end, dtype=np.int64)
        return (self.page_table_swa[req.req_slot, pos // self.ps].astype(np.int64) * self.ps + pos % self.ps).astype(np.int32)

    def alloc_for_decode(self, reqs: List[Any], req_slots: np.ndarray, positions: np.ndarray) -> Optional[np.ndarray]:
        ps = self.ps
        need = np.flatnonzero((positions % ps) == 0)
        if need.size and need.size > self.swa_available_pages():
            return None
        out = super().alloc_for_decode(reqs, req_slots, positions)
        if out is None:
            return None
        if need.size:
            pages = self._alloc_swa(int(need.size))
            if pages is None:               # cannot happen after the availability check above unless the tree is locked solid
                raise RuntimeError("sliding-window pool exhausted after the availability check")
            cols = positions[need] // ps
            self.page_table_swa[req_slots[need], cols] = pages
            self._record_swa(req_slots[need], cols, pages)
        return out

    def decode_slots_swa(self, req_slots: np.ndarray, positions: np.ndarray) -> np.ndarray:
        ps = self.ps
        return (self.page_table_swa[req_slots, positions // ps].astype(np.int64) * ps + positions % ps).astype(np.int32)

    def pages_needed_for_decode(self, positions: np.ndarray) -> int:
        return int(((positions % self.ps) == 0).sum())

    # ---- sliding
    def slide(self, req) -> int:
        """give back the request's OWN sliding pages that no token of the next step can see (tree-owned prefix pages stay: they are locked)"""
        if req.req_slot is None:
            return 0
        first_visible = max(0, req.kv_len - self.window)           # oldest position the next token attends to
        # the attention kernels fetch whole 128-key tiles: keep everything from the tile that holds the first visible key
        gran = self.fetch_granularity
        hi = min((first_visible // gran) * gran // self.ps, req.num_pages)          # pages [.., hi) are never fetched again
        lo0 = getattr(req, "swa_lo", 0)
        if hi <= lo0:
            return 0
        freed = 0
        if lo0 < min(hi, req.num_cached_pages) and req.last_node is not None:
            # prefix pages that live in the tree: tombstone them when this request is the only one holding them
            c_hi = min(hi, req.num_cached_pages)
            freed += self.tree.tombstone_exclusive(req.last_node, lo0, c_hi)
            cols = np.arange(lo0, c_hi)
            gone = self.tree._path_swa(req.last_node, req.num_cached_pages)[lo0:c_hi] == 0
            if gone.any():
                self.page_table_swa[req.req_slot, cols[gone]] = 0
                self._record_swa(np.full(int(gone.sum()), req.req_slot), cols[gone], np.zeros(int(gone.sum()), dtype=np.int32))
        lo = max(lo0, req.num_cached_pages)
        if hi <= lo:
            req.swa_lo = max(lo0, min(hi, req.num_cached_pages))
            self.swa_slid_pages += freed
            return freed
        cols = np.arange(lo, hi)
        pages = self.page_table_swa[req.req_slot, cols]
        live = pages[pages > 0]
        if live.size:
            self.swa_allocator.free(live.copy())
        self.page_table_swa[req.req_slot, cols] = 0
        self._record_swa(np.full(cols.size, req.req_slot), cols, np.zeros(cols.size, dtype=np.int32))
        req.swa_lo = hi
        self.swa_slid_pages += int(live.size) + freed
        return int(live.size) + freed

    def take_deltas(self) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        """ONE device table [reqs, 2 * P]: columns [0, P) full-attention pages, [P, 2P) sliding pages -> one delta stream"""
        r1, c1, v1 = super().take_deltas()
        r2, c2, v2 = self.take_swa_deltas()
        if r2.size == 0:
            return r1, c1, v1
        P = self.page_table.shape[1]
        return np.concatenate([r1, r2]), np.concatenate([c1, c2 + P]).astype(np.int32), np.concatenate([v1, v2])

    # ---- caching / release
    def cache_unfinished(self, req) -> None:
        ps = self.ps
        ids = req.all_ids()
        n_al = min(req.kv_len, len(ids)) // ps
        if n_al <= req.num_cached_pages or getattr(self.tree, "disable", False):
            return
        own = self.page_table[req.req_slot, :n_al].copy()
        own_swa = self.page_table_swa[req.req_slot, :n_al].copy()
        matched, adopted = self.tree.insert_swa(ids[: n_al * ps], own, own_swa, req.extra_key, req.priority)
        c = req.num_cached_pages
        if matched > c:
            self.allocator.free(own[c:matched])
            dup = own_swa[c:matched][~adopted[c:matched]]
            dup = dup[dup > 0]
            if dup.size:
                self.swa_allocator.free(dup)
        pages, swa, node = self._path_pages(ids[: n_al * ps], req.extra_key, n_al)
        self.tree.inc_lock_ref(node)
        self.tree.dec_lock_ref(req.last_node)
        req.last_node = node
        changed = np.flatnonzero(pages != own)
        if changed.size:
            self.page_table[req.req_slot, changed] = pages[changed]
            self._record(np.full(changed.size, req.req_slot), changed, pages[changed])
        ch2 = np.flatnonzero(swa != own_swa)
        if ch2.size:
            self.page_table_swa[req.req_slot, ch2] = swa[ch2]
            self._record_swa(np.full(ch2.size, req.req_slot), ch2, swa[ch2])
        req.num_cached_pages = n_al

    def _path_pages(self, ids, extra_key, n_pages: int):
        pages, node = self.tree.match_prefix(ids, extra_key)
        assert pages.size == n_pages, (pages.size, n_pages)
        return pages, self.tree._path_swa(node, n_pages), node

    def release(self, req, insert: bool = True) -> None:
        if req.req_slot is None:
            return
        ps = self.ps
        own = self.page_table[req.req_slot, : req.num_pages].copy()
        own_swa = self.page_table_swa[req.req_slot, : req.num_pages].copy()
        ids = req.all_ids()
        n_al = min(req.kv_len, len(ids)) // ps if insert and not getattr(self.tree, "disable", False) else req.num_cached_pages
        n_al = max(n_al, req.num_cached_pages)
        c = req.num_cached_pages
        if n_al > c:
            matched, adopted = self.tree.insert_swa(ids[: n_al * ps], own[:n_al], own_swa[:n_al], req.extra_key, req.priority)
            if matched > c:
                self.allocator.free(own[c:matched])
                dup = own_swa[c:matched][~adopted[c:matched]]
                dup = dup[dup > 0]
                if dup.size:
                    self.swa_allocator.free(dup)
        if req.num_pages > n_al:
            self.allocator.free(own[n_al:])
            tail = own_swa[n_al:]
            tail = tail[tail > 0]
            if tail.size:
                self.swa_allocator.free(tail)
        if req.last_node is not None:
            self.tree.dec_lock_ref(req.last_node)
        self.page_table_swa[req.req_slot, :] = 0
        self.slots.free(req.req_slot)
        req.req_slot = None
        req.num_pages = 0
        req.num_cached_pages = 0
        req.last_node = None

    def trim(self, req, n_tokens: int) -> None:
        need = max((n_tokens + self.ps - 1) // self.ps, req.num_cached_pages)
        if req.num_pages > need:
            extra = self.page_table_swa[req.req_slot, need:req.num_pages].copy()
            extra = extra[extra > 0]
            if extra.size:
                self.swa_allocator.free(extra)
            self.page_table_swa[req.req_slot, need:req.num_pages] = 0
        super().trim(req, n_tokens)

    # ---- invariants
    def check_invariants(self, live_reqs: List[Any]) -> None:
        super().check_invariants(live_reqs)
        tree_pages = self.tree.all_swa_pages()
        owned = []
        for r in live_reqs:
            if r.req_slot is not None:
                row = self.page_table_swa[r.req_slot, r.num_cached_pages: r.num_pages]
                owned.append(row[row > 0])
        owned_arr = np.concatenate(owned) if owned else _EMPTY
        used = np.concatenate([tree_pages, owned_arr])
        if np.unique(used).size != used.size:
            raise AssertionError("sliding-window page owned twice")
        if self.swa_allocator.available_pages() + used.size != self.swa_allocator.total_pages:
            raise AssertionError(f"sliding-window page leak: free {self.swa_allocator.available_pages()} + used {used.size} != "
                                 f"total {self.swa_allocator.total_pages}")
