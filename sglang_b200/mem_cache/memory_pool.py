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
                 dtype: torch.dtype, device: str, latent_v_dim: int = 0):
        """``latent_v_dim > 0``: MLA latent cache (reference ``MLATokenToKVPool``) — ONE plane of ``head_dim``-wide rows per token
        (compressed KV | rope key); the "value" the absorbed attention reads is the first ``latent_v_dim`` columns of the same row, so
        nothing is stored twice."""
        self.num_pages, self.page_size = num_pages, page_size
        self.num_layers, self.num_kv_heads, self.head_dim = num_layers, num_kv_heads, head_dim
        self.dtype, self.device = dtype, device
        self.num_slots = num_pages * page_size
        self.latent_v_dim = int(latent_v_dim)
        self.planes = 1 if self.latent_v_dim > 0 else 2
        self.buf = torch.zeros(num_layers, self.planes, self.num_slots, num_kv_heads, head_dim, dtype=dtype, device=device)
        # per-tensor scales for fp8 KV (1.0 for bf16)
        self.k_scale = 1.0
        self.v_scale = 1.0

    @staticmethod
    def bytes_per_page(page_size: int, num_layers: int, num_kv_heads: int, head_dim: int, dtype: torch.dtype, planes: int = 2) -> int:
        return planes * num_layers * page_size * num_kv_heads * head_dim * torch.empty((), dtype=dtype).element_size()

    def k(self, layer: int) -> torch.Tensor:
        return self.buf[layer, 0]

    def v(self, layer: int) -> torch.Tensor:
        if self.planes == 1:          # MLA: a column view of the latent rows (writes go through k())
            return self.buf[layer, 0][..., : self.latent_v_dim]
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
        """[layers, planes, num_pages, page_size, kv_heads, head_dim]"""
        return self.buf.view(self.num_layers, self.planes, self.num_pages, self.page_size, self.num_kv_heads, self.head_dim)

    def gather_pages(self, pages: torch.Tensor) -> torch.Tensor:
        return self.page_view().index_select(2, pages)

    def scatter_pages(self, pages: torch.Tensor, data: torch.Tensor) -> None:
        self.page_view().index_copy_(2, pages, data.to(self.buf.dtype))


class HybridKVPool:
    """Two KV planes for models that interleave full-attention and sliding-window layers (mem_cache/swa.py): ``full`` holds the
    full-attention layers for the whole context, ``swa`` the sliding layers for a window per request.  Layer ids are the model's."""

    def __init__(self, full: KVPool, swa: KVPool, sliding_layers):
        self.full, self.swa = full, swa
        self.sliding = sorted(int(i) for i in sliding_layers)
        n = full.num_layers + swa.num_layers
        self._loc = {}
        fi = si = 0
        sl = set(self.sliding)
        for i in range(n):
            if i in sl:
                self._loc[i] = (swa, si); si += 1
            else:
                self._loc[i] = (full, fi); fi += 1
        self.num_pages, self.page_size = full.num_pages, full.page_size
        self.num_layers, self.num_kv_heads, self.head_dim = n, full.num_kv_heads, full.head_dim
        self.dtype, self.device = full.dtype, full.device
        self.k_scale, self.v_scale = full.k_scale, full.v_scale
        self.mamba = None

    def is_sliding(self, layer: int) -> bool:
        return self._loc[layer][0] is self.swa

    def k(self, layer: int) -> torch.Tensor:
        p, i = self._loc[layer]
        return p.buf[i, 0]

    def v(self, layer: int) -> torch.Tensor:
        p, i = self._loc[layer]
        return p.buf[i, 1]


class HostKVPool:
    """Pinned-host mirror used by the HiCache L2 tier (reference ``memory_pool_host.py``).

    Layouts (``--hicache-mem-layout``): ``layer_first`` ``[L, 2, pages, ps, H, D]`` mirrors the device pool; ``page_first``
    ``[pages, L, 2, ps, H, D]`` makes every host page ONE contiguous record, which is what the L3 storage back ends read and write.
    On CUDA pages move with ``csrc/kernels/kvio.cu`` (one launch for all layers, the SMs address the pinned pool directly over UVA);
    on CPU (tests) with index copies."""

    def __init__(self, num_pages: int, like: KVPool, pin: bool = True, layout: str = "layer_first"):
        assert layout in ("layer_first", "page_first"), layout
        self.num_pages, self.layout = num_pages, layout
        L, ps, H, D = like.num_layers, like.page_size, like.num_kv_heads, like.head_dim
        P = getattr(like, "planes", 2)            # 1 for the MLA latent pool
        shape = (L, P, num_pages, ps, H, D) if layout == "layer_first" else (num_pages, L, P, ps, H, D)
        self.buf = torch.empty(shape, dtype=like.dtype, device="cpu")
        if pin and torch.cuda.is_available():
            self.buf = self.buf.pin_memory()
        self.planes = L * P
        self.chunk_bytes = ps * H * D * self.buf.element_size()
        self._free = list(range(num_pages - 1, -1, -1))
        self.kernel_transfers = 0

    # ---- page access independent of the layout
    def page(self, hp: int) -> torch.Tensor:
        """[L, 2, ps, H, D] view of one host page"""
        return self.buf[:, :, hp] if self.layout == "layer_first" else self.buf[hp]

    def _strides(self):
        cb = self.chunk_bytes
        if self.layout == "layer_first":
            return self.num_pages * cb, cb          # plane stride, page stride
        return cb, self.planes * cb

    def _use_kernel(self, like: KVPool) -> bool:
        if not (like.buf.is_cuda and self.buf.is_pinned()):
            return False
        from .. import ops

        return ops.load_native()

    def write_from_device(self, like: KVPool, dev_pages: np.ndarray, host_pages: np.ndarray) -> None:
        """device pages -> host pages (all layers, K and V)"""
        if self._use_kernel(like):
            dev = like.buf.device
            dp = torch.from_numpy(np.ascontiguousarray(dev_pages, dtype=np.int32)).to(dev)
            hp = torch.from_numpy(np.ascontiguousarray(host_pages, dtype=np.int32)).to(dev)
            ps_, hs_ = self._strides()
            torch.ops.sglb.kv_transfer_pages(like.buf, self.buf, dp, hp, self.planes, like.num_slots * like.num_kv_heads * like.head_dim *
                                             like.buf.element_size(), ps_, hs_, self.chunk_bytes, True)
            torch.cuda.current_stream().synchronize()      # the host copy must be complete before the device pages are recycled
            self.kernel_transfers += 1
            return
        data = like.gather_pages(torch.from_numpy(np.asarray(dev_pages, dtype=np.int64)).to(like.buf.device)).to("cpu")   # [L, 2, n, ps, H, D]
        idx = torch.from_numpy(np.asarray(host_pages, dtype=np.int64))
        if self.layout == "layer_first":
            self.buf[:, :, idx] = data
        else:
            self.buf[idx] = data.permute(2, 0, 1, 3, 4, 5)

    def read_to_device(self, like: KVPool, host_pages: np.ndarray, dev_pages: np.ndarray) -> None:
        """host pages -> device pages"""
        if self._use_kernel(like):
            dev = like.buf.device
            dp = torch.from_numpy(np.ascontiguousarray(dev_pages, dtype=np.int32)).to(dev)
            hp = torch.from_numpy(np.ascontiguousarray(host_pages, dtype=np.int32)).to(dev)
            ps_, hs_ = self._strides()
            torch.ops.sglb.kv_transfer_pages(like.buf, self.buf, dp, hp, self.planes, like.num_slots * like.num_kv_heads * like.head_dim *
                                             like.buf.element_size(), ps_, hs_, self.chunk_bytes, False)
            self.kernel_transfers += 1
            return
        idx = torch.from_numpy(np.asarray(host_pages, dtype=np.int64))
        data = self.buf[:, :, idx] if self.layout == "layer_first" else self.buf[idx].permute(1, 2, 0, 3, 4, 5)
        like.scatter_pages(torch.from_numpy(np.asarray(dev_pages, dtype=np.int64)).to(like.buf.device), data.to(like.buf.device))

    def available(self) -> int:
        return len(self._free)

    def alloc(self, n: int) -> Optional[np.ndarray]:
        if n > len(self._free):
            return None
        out = [self._free.pop() for _ in range(n)]
        return np.asarray(out, dtype=np.int32)

    def free(self, pages) -> None:
        self._free.extend(int(p) for p in pages)
