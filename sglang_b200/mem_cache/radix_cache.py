"""RadixAttention prefix cache (page-granular radix tree over token ids).

Semantics follow the reference protocol (SURVEY §3.5, Appendix A.6;
``python/sglang/srt/mem_cache/radix_cache.py``: ``match_prefix:377``, ``cache_finished_req:459``,
``cache_unfinished_req:516``, ``evict:593``, ``inc_lock_ref:623``, node split ``:705``; eviction
strategies ``evict_policy.py:16-53``).  Design differences (B200-first):

* values are **page ids**, not token slots: a node owning ``k`` pages covers ``k*page_size`` tokens;
  the device page table is built straight from these arrays, so attention kernels TMA-gather whole
  pages and a shared prefix is one run of identical page ids across requests.
* this file is the executable specification + fallback; the production tree is the C++
  ``sglb_rt.RadixTree`` (``csrc/runtime/radix_tree.cpp``) with the identical API
  (see ``create_radix_cache``); ``tests/test_radix_cache.py`` runs both differentially.
"""
from __future__ import annotations

import heapq
import time
from typing import Any, Callable, Dict, List, Optional, Tuple

import numpy as np

from .allocator import PagedAllocator

_EMPTY = np.empty(0, dtype=np.int32)


class TreeNode:
    __slots__ = ("children", "parent", "key", "pages", "lock_ref", "last_access", "hit_count", "created",
                 "priority", "host_pages", "id", "hashes", "swa_pages", "mamba_slot")
    _ctr = 0

    def __init__(self) -> None:
        self.children: Dict[Any, "TreeNode"] = {}
        self.parent: Optional["TreeNode"] = None
        self.key: np.ndarray = _EMPTY       # token ids, len == len(pages) * page_size
        self.pages: Optional[np.ndarray] = _EMPTY    # device page ids (None => evicted to host)
        self.host_pages: Optional[np.ndarray] = None  # HiCache L2 copy
        self.lock_ref = 0
        self.last_access = 0.0
        self.created = 0.0
        self.hit_count = 0
        self.priority = 0
        self.hashes: Optional[np.ndarray] = None     # chained 64-bit page hashes (only when KV events are enabled)
        self.mamba_slot: Optional[int] = None        # recurrent-state snapshot row taken at the END of this node (mem_cache/mamba_radix.py)
        self.swa_pages: Optional[np.ndarray] = None  # sliding-window pool pages of the same tokens (0 = tombstone: slid out / reclaimed)
        TreeNode._ctr += 1
        self.id = TreeNode._ctr

    @property
    def evicted(self) -> bool:
        return self.pages is None

    def __lt__(self, other: "TreeNode") -> bool:  # heap tie-break
        return self.id < other.id


_POLICIES: Dict[str, Callable[[TreeNode], Any]] = {
    "lru": lambda n: n.last_access,
    "lfu": lambda n: (n.hit_count, n.last_access),
    "fifo": lambda n: n.created,
    "mru": lambda n: -n.last_access,
    "filo": lambda n: -n.created,
    "priority": lambda n: (n.priority, n.last_access),
}


def _root_seed(extra_key: Any) -> int:
    """namespace seed of the hash chain (LoRA adapter / cache salt): requests in different namespaces never share a prefix"""
    if extra_key is None:
        return 0
    import hashlib

    # 63-bit keyed digest (a 32-bit CRC made a collision between two cache salts / adapter ids merge their KV namespaces)
    return (int.from_bytes(hashlib.blake2b(repr(extra_key).encode(), digest_size=8, key=b"sglb-radix-ns").digest(), "little") >> 1) | 1


def chain_hashes(tokens: np.ndarray, page_size: int, seed: int = 0) -> np.ndarray:
    """64-bit FNV-1a page hashes, each folding in its predecessor (C++ ``sglb_rt.hash_pages``; numpy fallback)"""
    try:
        from .. import sglb_rt

        return np.asarray(sglb_rt.hash_pages(np.ascontiguousarray(tokens, dtype=np.int32), page_size, int(seed)), dtype=np.uint64)
    except Exception:
        out, prev = [], int(seed)
        b = np.ascontiguousarray(tokens, dtype=np.int32).view(np.uint8)
        step = 4 * page_size
        for p in range(len(tokens) // page_size):
            h = 1469598103934665603 ^ prev
            for v in b[p * step:(p + 1) * step].tolist():
                h = ((h ^ v) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
            out.append(h)
            prev = h
        return np.asarray(out, dtype=np.uint64)


class RadixCache:
    """Python reference implementation."""

    def __init__(self, allocator: Optional[PagedAllocator], page_size: int = 1, eviction_policy: str = "lru",
                 disable: bool = False, enable_kv_events: bool = False, clock: Callable[[], float] = time.monotonic):
        self.allocator = allocator
        self.page_size = page_size
        self.disable = disable
        self.policy_name = eviction_policy
        self._policy = _POLICIES[eviction_policy]
        self._clock = clock
        self.enable_kv_events = enable_kv_events
        self.kv_events: List[Tuple[str, Any]] = []
        self.reset()

    # ------------------------------------------------------------------ basic state
    def reset(self) -> None:
        self.roots: Dict[Any, TreeNode] = {}
        self._evictable_pages = 0
        self._protected_pages = 0
        self._tick = 0.0
        if self.enable_kv_events:
            self.kv_events.append(("AllBlocksCleared", None))

    def _now(self) -> float:
        # strictly increasing logical clock so LRU order is deterministic in tests
        self._tick = max(self._clock(), self._tick + 1e-9)
        return self._tick

    def _root(self, extra_key: Any) -> TreeNode:
        r = self.roots.get(extra_key)
        if r is None:
            r = TreeNode()
            r.lock_ref = 1  # roots are never evicted
            self.roots[extra_key] = r
        return r

    def _child_key(self, key: np.ndarray, off: int = 0) -> Any:
        if self.page_size == 1:
            return int(key[off])
        return key[off:off + self.page_size].tobytes()

    def evictable_size(self) -> int:
        return self._evictable_pages * self.page_size

    def protected_size(self) -> int:
        return self._protected_pages * self.page_size

    def evictable_pages(self) -> int:
        return self._evictable_pages

    def protected_pages(self) -> int:
        return self._protected_pages

    def total_size(self) -> int:
        return (self._evictable_pages + self._protected_pages) * self.page_size

    # ------------------------------------------------------------------ match
    def _match_len_pages(self, a: np.ndarray, b: np.ndarray) -> int:
        """number of whole pages on which a and b agree"""
        n = min(a.size, b.size)
        n -= n % self.page_size
        if n == 0:
            return 0
        ne = np.flatnonzero(a[:n] != b[:n])
        m = n if ne.size == 0 else int(ne[0])
        return m // self.page_size

    def match_prefix(self, token_ids, extra_key: Any = None) -> Tuple[np.ndarray, TreeNode]:
        """Longest cached page-aligned prefix.  Returns (page ids, last node).  May split a node."""
        key = np.asarray(token_ids, dtype=np.int32)
        root = self._root(extra_key)
        if self.disable or key.size < self.page_size:
            return _EMPTY, root
        key = key[: key.size - key.size % self.page_size]
        now = self._now()
        node, off, out = root, 0, []
        ps = self.page_size
        while off < key.size:
            child = node.children.get(self._child_key(key, off))
            if child is None or child.evicted:
                break
            m = self._match_len_pages(child.key, key[off:])
            if m == 0:
                break
            if m * ps < child.key.size:
                child = self._split(child, m)
            child.last_access = now
            child.hit_count += 1
            out.append(child.pages)
            node = child
            off += m * ps
            if m * ps < child.key.size:  # pragma: no cover (split makes them equal)
                break
        pages = np.concatenate(out) if out else _EMPTY
        return pages, node

    def _split(self, child: TreeNode, m_pages: int) -> TreeNode:
        """split ``child`` so that a new parent node holds its first ``m_pages`` pages; returns it"""
        ps = self.page_size
        new = TreeNode()
        new.parent = child.parent
        new.key = child.key[: m_pages * ps]
        new.pages = None if child.pages is None else child.pages[:m_pages]   # None: host-resident (HiCache) node
        new.lock_ref = child.lock_ref
        new.last_access, new.created, new.hit_count, new.priority = (child.last_access, child.created,
                                                                     child.hit_count, child.priority)
        if child.host_pages is not None:
            new.host_pages, child.host_pages = child.host_pages[:m_pages], child.host_pages[m_pages:]
        if child.hashes is not None:
            new.hashes, child.hashes = child.hashes[:m_pages], child.hashes[m_pages:]
        if child.swa_pages is not None:
            new.swa_pages, child.swa_pages = child.swa_pages[:m_pages], child.swa_pages[m_pages:]
        child.parent.children[self._child_key(new.key)] = new
        child.key = child.key[m_pages * ps:]
        if child.pages is not None:
            child.pages = child.pages[m_pages:]
        child.parent = new
        new.children[self._child_key(child.key)] = child
        return new

    # ------------------------------------------------------------------ insert
    def insert(self, token_ids, pages, extra_key: Any = None, priority: int = 0) -> int:
        """Insert ``pages`` (one per page of ``token_ids``).  Returns the number of leading pages that
        were already present (the caller owns — and must free — its duplicates of those)."""
        if self.disable:
            return 0
        key = np.asarray(token_ids, dtype=np.int32)
        pages = np.asarray(pages, dtype=np.int32)
        ps = self.page_size
        n_pages = min(key.size // ps, pages.size)
        key, pages = key[: n_pages * ps], pages[:n_pages]
        node, off, matched = self._root(extra_key), 0, 0
        now = self._now()
        while off < key.size:
            ck = self._child_key(key, off)
            child = node.children.get(ck)
            if child is None:
                break
            m = self._match_len_pages(child.key, key[off:])
            if m == 0:  # pragma: no cover - child key starts with this page by construction
                break
            if m * ps < child.key.size:
                child = self._split(child, m)
            if child.evicted:  # re-materialise a host-only node with the fresh device pages
                child.pages = pages[off // ps: off // ps + m].copy()
                self._evictable_pages += m if child.lock_ref == 0 else 0
                self._protected_pages += m if child.lock_ref > 0 else 0
            else:
                matched += m
            child.last_access = now
            child.priority = max(child.priority, priority)
            node = child
            off += m * ps
        if off < key.size:
            new = TreeNode()
            new.parent = node
            new.key = key[off:].copy()
            new.pages = pages[off // ps:].copy()
            new.last_access = new.created = now
            new.priority = priority
            node.children[self._child_key(new.key)] = new
            self._evictable_pages += new.pages.size
            if self.enable_kv_events:
                # chained page hashes identify the PREFIX, so a fleet-level index can match requests without seeing page ids
                parent_hash = int(node.hashes[-1]) if node.hashes is not None and node.hashes.size else _root_seed(extra_key)
                new.hashes = chain_hashes(new.key, ps, parent_hash)
                self.kv_events.append(("BlockStored", {"pages": new.pages.tolist(), "hashes": new.hashes.tolist(), "parent_hash": parent_hash}))
        return matched

    # ------------------------------------------------------------------ locking
    def inc_lock_ref(self, node: TreeNode) -> None:
        while node is not None and node.parent is not None:
            if node.lock_ref == 0 and not node.evicted:
                self._evictable_pages -= node.pages.size
                self._protected_pages += node.pages.size
            node.lock_ref += 1
            node = node.parent

    def dec_lock_ref(self, node: TreeNode) -> None:
        while node is not None and node.parent is not None:
            assert node.lock_ref > 0, "dec_lock_ref underflow"
            node.lock_ref -= 1
            if node.lock_ref == 0 and not node.evicted:
                self._evictable_pages += node.pages.size
                self._protected_pages -= node.pages.size
            node = node.parent

    def lock_cost(self, node: TreeNode) -> int:
        """evictable pages that would become protected if ``node`` were locked"""
        cost = 0
        while node is not None and node.parent is not None and node.lock_ref == 0:
            if not node.evicted:
                cost += int(node.pages.size)
            node = node.parent
        return cost

    # ------------------------------------------------------------------ evict
    def _leaves(self) -> List[TreeNode]:
        out, stack = [], [r for r in self.roots.values()]
        while stack:
            n = stack.pop()
            live_children = [c for c in n.children.values() if not c.evicted]
            if not live_children and n.parent is not None and not n.evicted:
                out.append(n)
            stack.extend(live_children)
        return out

    def evict(self, num_pages: int, on_evict: Optional[Callable[[TreeNode], bool]] = None) -> int:
        """Evict >= num_pages pages of unlocked leaves (policy order).  Returns pages freed."""
        if self.disable or num_pages <= 0:
            return 0
        heap = [(self._policy(n), n) for n in self._leaves() if n.lock_ref == 0]
        heapq.heapify(heap)
        freed = 0
        while heap and freed < num_pages:
            _, n = heapq.heappop(heap)
            if n.lock_ref > 0 or n.evicted:
                continue
            keep_host = bool(on_evict(n)) if on_evict is not None else False
            if self.allocator is not None:
                self.allocator.free(n.pages)
            freed += n.pages.size
            self._evictable_pages -= n.pages.size
            if self.enable_kv_events:
                self.kv_events.append(("BlockRemoved", {"pages": n.pages.tolist(), "hashes": n.hashes.tolist() if n.hashes is not None else []}))
            parent = n.parent
            if keep_host and n.host_pages is not None:
                n.pages = None  # stays in the tree as a host-resident node
            else:
                del parent.children[self._child_key(n.key)]
            if parent.parent is not None and parent.lock_ref == 0 and not parent.evicted and \
                    not any(not c.evicted for c in parent.children.values()):
                heapq.heappush(heap, (self._policy(parent), parent))
        return freed

    # ------------------------------------------------------------------ request helpers
    def take_kv_events(self) -> List[Tuple[str, Any]]:
        ev, self.kv_events = self.kv_events, []
        return ev

    def all_pages(self) -> np.ndarray:
        out, stack = [], list(self.roots.values())
        while stack:
            n = stack.pop()
            if n.parent is not None and not n.evicted:
                out.append(n.pages)
            stack.extend(n.children.values())
        return np.concatenate(out) if out else _EMPTY

    def pretty(self) -> str:
        lines: List[str] = []

        def rec(n: TreeNode, d: int) -> None:
            for c in n.children.values():
                lines.append("  " * d + f"[{c.key.size} tok, {0 if c.evicted else c.pages.size} pg, lock={c.lock_ref}]")
                rec(c, d + 1)

        for k, r in self.roots.items():
            lines.append(f"root({k!r})")
            rec(r, 1)
        return "\n".join(lines)


def create_radix_cache(allocator, page_size: int, eviction_policy: str = "lru", disable: bool = False,
                       impl: str = "auto", enable_kv_events: bool = False):
    """Factory: C++ tree when the native runtime is built (default), Python otherwise."""
    if impl in ("auto", "cpp") and not enable_kv_events:
        try:
            from .radix_cache_cpp import RadixCacheCpp

            return RadixCacheCpp(allocator, page_size, eviction_policy, disable)
        except Exception:
            if impl == "cpp":
                raise
    return RadixCache(allocator, page_size, eviction_policy, disable, enable_kv_events)
