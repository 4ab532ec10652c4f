"""CPU-side page allocator for the paged KV cache.

Parity: reference ``python/sglang/srt/mem_cache/allocator/paged.py:172`` (PagedTokenToKVPoolAllocator)
and ``memory_pool.py:256`` (ReqToTokenPool).  B200-first difference: the allocator is *pure host
state* (a LIFO page stack in a numpy array, or the C++ ``sglb_rt.PageAllocator``); the device never
sees allocation — it only receives the per-step packed page table.  Page 0 is the dummy page that
graph-padded tokens write to (reference ``memory_pool.py:1955-1975``).
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np


class PagedAllocator:
    def __init__(self, num_pages: int, page_size: int):
        assert num_pages >= 2
        self.num_pages = num_pages
        self.page_size = page_size
        self.clear()

    def clear(self) -> None:
        # page 0 reserved as the dummy page
        self._free = np.arange(self.num_pages - 1, 0, -1, dtype=np.int32)  # stack: top at the end
        self._n_free = self.num_pages - 1
        self._is_free = np.ones(self.num_pages, dtype=bool)
        self._is_free[0] = False

    @property
    def total_pages(self) -> int:
        return self.num_pages - 1

    def available_pages(self) -> int:
        return self._n_free

    def available_tokens(self) -> int:
        return self._n_free * self.page_size

    def alloc(self, n: int) -> Optional[np.ndarray]:
        if n > self._n_free:
            return None
        if n == 0:
            return np.empty(0, dtype=np.int32)
        out = self._free[self._n_free - n:self._n_free][::-1].copy()
        self._n_free -= n
        self._is_free[out] = False
        return out

    def free(self, pages) -> None:
        pages = np.asarray(pages, dtype=np.int32).reshape(-1)
        n = pages.size
        if n == 0:
            return
        if self._is_free[pages].any() or (pages <= 0).any():
            bad = pages[(self._is_free[pages]) | (pages <= 0)]
            raise RuntimeError(f"double free / invalid free of KV pages {bad[:8].tolist()}")
        self._is_free[pages] = True
        self._free[self._n_free:self._n_free + n] = pages
        self._n_free += n

    def is_free(self, page: int) -> bool:
        return bool(self._is_free[page])


class ReqSlotPool:
    """Request-slot pool (index into per-request device state such as ``next_token_buf``).
    Slot 0 is the padding row for graph-padded batches (reference ``memory_pool.py:256-285``)."""

    def __init__(self, size: int):
        self.size = size
        self.clear()

    def clear(self) -> None:
        self._free: List[int] = list(range(self.size, 0, -1))

    def available(self) -> int:
        return len(self._free)

    def alloc(self) -> Optional[int]:
        return self._free.pop() if self._free else None

    def free(self, idx: int) -> None:
        assert 0 < idx <= self.size and idx not in self._free
        self._free.append(idx)
