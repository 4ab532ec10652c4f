"""Device KV pools.

Parity: reference ``python/sglang/srt/mem_cache/memory_pool.py`` (``MHATokenToKVPool:1759``,
``set_kv_buffer/get_key_buffer :1707-1719``).  Layout (B200-first): ONE allocation
``[layers, 2(K/V), num_pages*page_size, kv_heads, head_dim]`` so that (a) each (layer, K|V) plane is a
dense 3-D tensor a single TMA descriptor can tile (inner dim = head_dim, 128-B swizzled rows), (b) PD /
HiCache page transfers are strided copies of one tensor, (c) a page is contiguous per plane:
``page_size*kv_heads*head_dim`` elements.  Slot id = ``page_id*page_size + offset``.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch


class KVPool:
    def __init__(self, num_pages: int, page_size: int, num_layers: int, num_kv_heads: int, head_dim: int,
                 dtype: torch.dtype, device: str):
        self.num_pages, self.page_size = num_pages, page_size
        self.num_layers, self.num_kv_heads, self.head_dim = num_layers, num_kv_heads, head_dim
        self.dtype, self.device = dtype, device
        self.num_slots = num_pages * page_size
        self.buf = torch.zeros(num_layers, 2, self.num_slots, num_kv_heads, head_dim, dtype=dtype, device=device)
        # per-tensor scales for fp8 KV (1.0 for bf16)
        self.k_scale = 1.0
        self.v_scale = 1.0

    @staticmethod
    def bytes_per_page(page_size: int, num_layers: int, num_kv_heads: int, head_dim: int, dtype: torch.dtype) -> int:
        return 2 * num_layers * page_size * num_kv_heads * head_dim * torch.empty((), dtype=dtype).element_size()

    def k(self, layer: int) -> torch.Tensor:
        return self.buf[layer, 0]

    def v(self, layer: int) -> torch.Tensor:
        return self.buf[layer, 1]

    def move_slots(self, src: torch.Tensor, dst: torch.Tensor) -> None:
        """copy KV rows ``src -> dst`` in every layer (gather first: the two sets may overlap).  Tree speculative decoding compacts the
        accepted path of a verified token tree into consecutive positions with this (reference ``move_kv_cache`` for topk > 1)."""
        if src.numel() == 0:
            return
        rows = self.buf.index_select(2, src.long())
        self.buf.index_copy_(2, dst.long(), rows)

    # page-granular views for transfer engines (PD disaggregation / HiCache)
    def page_view(self) -> torch.Tensor:
        """[layers, 2, num_pages, page_size, kv_heads, head_dim]"""
        return self.buf.view(self.num_layers, 2, self.num_pages, self.page_size, self.num_kv_heads, self.head_dim)

    def gather_pages(self, pages: torch.Tensor) -> torch.Tensor:
        return self.page_view().index_select(2, pages)

    def scatter_pages(self, pages: torch.Tensor, data: torch.Tensor) -> None:
        self.page_view().index_copy_(2, pages, data.to(self.buf.dtype))


class HostKVPool:
    """Pinned-host mirror used by the HiCache L2 tier (reference ``memory_pool_host.py``)."""

    def __init__(self, num_pages: int, like: KVPool, pin: bool = True):
        self.num_pages = num_pages
        shape = (like.num_layers, 2, num_pages, like.page_size, like.num_kv_heads, like.head_dim)
        self.buf = torch.empty(shape, dtype=like.dtype, device="cpu")
        if pin and torch.cuda.is_available():
            self.buf = self.buf.pin_memory()
        self._free = list(range(num_pages - 1, -1, -1))

    def available(self) -> int:
        return len(self._free)

    def alloc(self, n: int) -> Optional[np.ndarray]:
        if n > len(self._free):
            return None
        out = [self._free.pop() for _ in range(n)]
        return np.asarray(out, dtype=np.int32)

    def free(self, pages) -> None:
        self._free.extend(int(p) for p in pages)
