"""RadixCache backed by the C++ tree (``csrc/runtime/radix_tree.cpp``), same interface as
``radix_cache.RadixCache``.  Parity: reference ``RadixCacheCpp`` (``mem_cache/radix_cache_cpp.py:35``)."""
from __future__ import annotations

import zlib
from typing import Any, Tuple

import numpy as np

from .. import sglb_rt  # built by sglang_b200/build.py ; ImportError => caller falls back to python

_POLICY = {"lru": 0, "lfu": 1, "fifo": 2, "mru": 3, "filo": 4, "priority": 5}
_EMPTY = np.empty(0, dtype=np.int32)


_NS_IDS: dict = {}     # interned namespace ids: exact (a 32-bit CRC of the salt let two salts collide and share KV)


def _key(extra_key: Any) -> int:
    if extra_key is None:
        return 0
    if isinstance(extra_key, int) and not isinstance(extra_key, bool) and 0 <= extra_key < (1 << 40):
        return extra_key + 1
    k = repr(extra_key)
    v = _NS_IDS.get(k)
    if v is None:
        v = _NS_IDS[k] = (1 << 41) + len(_NS_IDS)
    return v


class _NodeHandle(int):
    """int node id; kept as a distinct type for readability"""


class RadixCacheCpp:
    def __init__(self, allocator, page_size: int = 1, eviction_policy: str = "lru", disable: bool = False):
        self.allocator = allocator
        self.page_size = page_size
        self.disable = disable
        self.policy_name = eviction_policy
        self._t = sglb_rt.RadixTree(page_size, _POLICY[eviction_policy])
        self.enable_kv_events = False

    def reset(self) -> None:
        self._t.reset()

    def match_prefix(self, token_ids, extra_key: Any = None) -> Tuple[np.ndarray, int]:
        if self.disable:
            _, node = self._t.match_prefix(_EMPTY, _key(extra_key))
            return _EMPTY, node
        return self._t.match_prefix(np.asarray(token_ids, dtype=np.int32), _key(extra_key))

    def insert(self, token_ids, pages, extra_key: Any = None, priority: int = 0) -> int:
        if self.disable:
            return 0
        return int(self._t.insert(np.asarray(token_ids, dtype=np.int32), np.asarray(pages, dtype=np.int32),
                                  _key(extra_key), int(priority)))

    def inc_lock_ref(self, node: int) -> None:
        self._t.inc_lock_ref(int(node))

    def dec_lock_ref(self, node: int) -> None:
        self._t.dec_lock_ref(int(node))

    def lock_cost(self, node: int) -> int:
        return int(self._t.lock_cost(int(node)))

    def evict(self, num_pages: int, on_evict=None) -> int:
        if self.disable or num_pages <= 0:
            return 0
        freed = self._t.evict(int(num_pages))
        if self.allocator is not None and freed.size:
            self.allocator.free(freed)
        return int(freed.size)

    def evictable_pages(self) -> int:
        return int(self._t.evictable_pages())

    def protected_pages(self) -> int:
        return int(self._t.protected_pages())

    def evictable_size(self) -> int:
        return self.evictable_pages() * self.page_size

    def protected_size(self) -> int:
        return self.protected_pages() * self.page_size

    def total_size(self) -> int:
        return self.evictable_size() + self.protected_size()

    def all_pages(self) -> np.ndarray:
        return self._t.all_pages()

    def take_kv_events(self):
        return []
