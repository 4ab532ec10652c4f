"""Hierarchical KV cache: L1 device pages -> L2 pinned host pages -> L3 storage (files).

Parity: reference ``python/sglang/srt/mem_cache/hiradix_cache.py:77`` (HiRadixCache), ``managers/cache_controller.py``
(HiCacheController: write-through / write-back, load-back on hit), ``memory_pool_host.py`` (host pools),
``hicache_storage.py:150,361`` (storage backends; ``HiCacheFile``), page transfer kernels ``kernels/aot/csrc/kvcacheio/transfer.cu``.

Design: the radix node keeps ``host_pages`` next to ``pages``.  Eviction of a device node with a host copy only drops
the device pages (the node stays in the tree as *host-resident*); a later prefix match that reaches a host-resident node
re-allocates device pages and loads the KV back with one strided H2D copy per node (the KV pool is a single tensor whose
page dimension is contiguous per (layer, K|V) plane, so a page set moves as one ``index_copy_``).  L3 spills host pages
to files named by the chained page hash (``sglb_rt.hash_pages``), giving cross-restart prefix reuse.
"""
from __future__ import annotations

import json
import logging
from typing import Any, List, Optional

import numpy as np
import torch

from .memory_pool import HostKVPool, KVPool
from .radix_cache import _EMPTY, RadixCache, TreeNode

logger = logging.getLogger(__name__)


from .hicache_storage import FileStorage, HiCacheStorage, create_storage  # noqa: E402,F401  (FileStorage re-exported for callers)


class HiRadixCache(RadixCache):
    def __init__(self, allocator: Any, page_size: int, kv_pool: KVPool, host_pages: int, eviction_policy: str = "lru",
                 write_policy: str = "write_through", storage: Optional[FileStorage] = None, mem_layout: str = "layer_first"):
        super().__init__(allocator, page_size, eviction_policy)
        self.kv_pool = kv_pool
        self.host = HostKVPool(host_pages, kv_pool, pin=kv_pool.buf.is_cuda, layout=mem_layout)
        if write_policy not in ("write_through", "write_through_selective", "write_back"):
            raise ValueError(f"unknown hicache write policy {write_policy!r}")
        self.write_policy = write_policy
        self.selective_hits = 2
        self.storage = storage
        self.stats = {"host_writes": 0, "host_hits_pages": 0, "l3_writes": 0, "l3_hits_pages": 0, "host_evictions": 0}

    # ---- L1 -> L2
    def _backup(self, node: TreeNode) -> bool:
        if node.host_pages is not None:
            return True
        n = int(node.pages.size)
        hp = self.host.alloc(n)
        if hp is None:
            self._evict_host(n)
            hp = self.host.alloc(n)
            if hp is None:
                return False
        self.host.write_from_device(self.kv_pool, node.pages, hp)      # kvio.cu page mover on CUDA (one launch, all layers)
        node.host_pages = hp
        self.stats["host_writes"] += n
        if self.storage is not None:
            self._spill_l3(node)
        return True

    def _spill_l3(self, node: TreeNode) -> None:
        from .. import sglb_rt

        key = self._full_key(node)
        hashes = sglb_rt.hash_pages(key, self.page_size)
        n = int(node.host_pages.size)
        hs = [int(h) for h in hashes[-n:]]
        have = self.storage.batch_exists(hs)                      # one round trip for the whole node, then one for the missing pages
        todo = [(h, hp) for h, hp, ok in zip(hs, node.host_pages.tolist(), have) if not ok]
        if todo:
            self.storage.batch_put([h for h, _ in todo], [self.host.page(hp) for _, hp in todo])
            self.stats["l3_writes"] += len(todo)

    def _full_key(self, node: TreeNode) -> np.ndarray:
        parts = []
        while node is not None and node.parent is not None:
            parts.append(node.key)
            node = node.parent
        return np.concatenate(parts[::-1]) if parts else _EMPTY

    def _evict_host(self, n: int) -> None:
        """drop host copies of the coldest host-only nodes"""
        cands = []
        stack = list(self.roots.values())
        while stack:
            x = stack.pop()
            stack.extend(x.children.values())
            if x.parent is not None and x.evicted and x.host_pages is not None and not x.children and x.lock_ref == 0:
                cands.append(x)
        cands.sort(key=lambda x: x.last_access)
        freed = 0
        for x in cands:
            self.host.free(x.host_pages)
            freed += x.host_pages.size
            x.host_pages = None
            del x.parent.children[self._child_key(x.key)]
            self.stats["host_evictions"] += 1
            if freed >= n:
                break

    def insert(self, token_ids, pages, extra_key: Any = None, priority: int = 0) -> int:
        m = super().insert(token_ids, pages, extra_key, priority)
        if self.write_policy in ("write_through", "write_through_selective"):
            # write_through: every new node is copied to the host tier right away; write_through_selective: only nodes that were HIT at
            # least ``selective_hits`` times (hot prefixes) — one-off prompts are backed up lazily at eviction like write_back
            _, node = RadixCache.match_prefix(self, token_ids, extra_key)
            while node is not None and node.parent is not None and node.host_pages is None and not node.evicted:
                if self.write_policy == "write_through" or node.hit_count >= self.selective_hits:
                    self._backup(node)
                node = node.parent
        return m

    def evict(self, num_pages: int, on_evict=None) -> int:
        def keep(node: TreeNode) -> bool:
            return self._backup(node)

        return super().evict(num_pages, on_evict=keep)

    # ---- L2 / L3 -> L1
    def match_prefix(self, token_ids, extra_key: Any = None):
        pages, node = super().match_prefix(token_ids, extra_key)
        key = np.asarray(token_ids, dtype=np.int32)
        key = key[: key.size - key.size % self.page_size]
        off = int(pages.size) * self.page_size
        out = [pages] if pages.size else []
        while off < key.size:
            child = node.children.get(self._child_key(key, off))
            if child is None or not child.evicted or child.host_pages is None:
                break
            m = self._match_len_pages(child.key, key[off:])
            if m == 0:
                break
            if m * self.page_size < child.key.size:
                child = self._split(child, m)
            if not self._load_back(child):
                break
            out.append(child.pages)
            node = child
            off += m * self.page_size
        if off < key.size and self.storage is not None:
            node, more = self._load_from_l3(node, key, off, extra_key)
            if more is not None:
                out.append(more)
        return (np.concatenate(out) if out else _EMPTY), node

    def _load_back(self, node: TreeNode) -> bool:
        n = int(node.host_pages.size)
        short = n - self.allocator.available_pages()
        if short > 0:
            super().evict(short, on_evict=lambda x: self._backup(x))
        dev_pages = self.allocator.alloc(n)
        if dev_pages is None:
            return False
        self.host.read_to_device(self.kv_pool, node.host_pages, dev_pages)
        node.pages = dev_pages
        self._evictable_pages += n if node.lock_ref == 0 else 0
        self._protected_pages += n if node.lock_ref > 0 else 0
        node.last_access = self._now()
        self.stats["host_hits_pages"] += n
        return True

    def _load_from_l3(self, node: TreeNode, key: np.ndarray, off: int, extra_key: Any):
        from .. import sglb_rt

        hashes = sglb_rt.hash_pages(key, self.page_size)
        first = off // self.page_size
        hs = [int(h) for h in hashes[first:]]
        have = self.storage.batch_exists(hs) if hs else []
        n_ok = next((i for i, ok in enumerate(have) if not ok), len(have))          # longest stored run from the first missing page
        got: List[torch.Tensor] = []
        for t in (self.storage.batch_get(hs[:n_ok]) if n_ok else []):
            if t is None:
                break
            got.append(t)
        if not got:
            return node, None
        n = len(got)
        dev_pages = self.allocator.alloc(n)
        if dev_pages is None:
            return node, None
        data = torch.stack(got, dim=2).to(self.kv_pool.buf.device)
        self.kv_pool.scatter_pages(torch.from_numpy(dev_pages.astype(np.int64)).to(self.kv_pool.buf.device), data)
        end = off + n * self.page_size
        RadixCache.insert(self, key[:end], np.concatenate([self._pages_along(key[:off], extra_key), dev_pages]), extra_key)
        self.stats["l3_hits_pages"] += n
        pages, node2 = RadixCache.match_prefix(self, key[:end], extra_key)
        return node2, dev_pages

    def _pages_along(self, key: np.ndarray, extra_key: Any) -> np.ndarray:
        if key.size == 0:
            return _EMPTY
        p, _ = RadixCache.match_prefix(self, key, extra_key)
        return p


def attach_hicache(scheduler: Any, runner: Any, args: Any) -> HiRadixCache:
    dev_pages = scheduler.allocator.total_pages
    host_pages = max(8, int(dev_pages * args.hicache_ratio))
    if runner.is_cuda:
        # bound pinned memory to something sane for a dev box; production sizes it from --hicache-ratio
        per_page = runner.kv_pool.buf[:, :, :1].numel() * runner.kv_pool.buf.element_size() * runner.page_size
        host_pages = min(host_pages, max(8, int(64 * (1 << 30) // max(1, per_page))))
    extra = json.loads(args.hicache_storage_backend_extra_config) if getattr(args, "hicache_storage_backend_extra_config", None) else {}
    if args.hicache_storage_dir:
        extra.setdefault("dir", args.hicache_storage_dir)
    storage = create_storage(args.hicache_storage_backend, extra)
    tree = HiRadixCache(scheduler.allocator, scheduler.page_size, runner.kv_pool, host_pages, args.radix_eviction_policy,
                        args.hicache_write_policy, storage, getattr(args, "hicache_mem_layout", "layer_first"))
    scheduler.tree = tree
    scheduler.kv.tree = tree
    return tree
