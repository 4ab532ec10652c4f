"""Radix prefix cache: python spec vs C++ tree, differential + semantics (reference test strategy:
test/registered/unit/mem_cache/test_radix_cache_unit.py)."""
import numpy as np
import pytest

from sglang_b200.mem_cache.allocator import PagedAllocator
from sglang_b200.mem_cache.radix_cache import RadixCache
from sglang_b200.mem_cache.radix_cache_cpp import RadixCacheCpp

IMPLS = [RadixCache, RadixCacheCpp]


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("ps", [1, 4])
def test_insert_match_split(impl, ps):
    alloc = PagedAllocator(64, ps)
    t = impl(alloc, ps)
    toks = list(range(100, 100 + 8 * ps))
    pages = alloc.alloc(8)
    assert t.insert(toks, pages) == 0
    got, node = t.match_prefix(toks)
    assert got.tolist() == pages.tolist()
    # diverge after 3 pages -> split
    other = toks[: 3 * ps] + [7] * (2 * ps)
    got2, node2 = t.match_prefix(other)
    assert got2.tolist() == pages[:3].tolist()
    p2 = alloc.alloc(2)
    matched = t.insert(other, np.concatenate([pages[:3], p2]))
    assert matched == 3
    assert t.evictable_pages() == 10 and t.protected_pages() == 0
    # partial page never matches
    got3, _ = t.match_prefix(toks[: 2 * ps + (ps - 1 if ps > 1 else 0)])
    assert got3.size == 2 if ps > 1 else got3.size == 2


@pytest.mark.parametrize("impl", IMPLS)
def test_lock_and_evict_lru(impl):
    ps = 2
    alloc = PagedAllocator(32, ps)
    t = impl(alloc, ps)
    a = [1, 2, 3, 4, 5, 6]
    b = [1, 2, 9, 9, 8, 8]
    pa = alloc.alloc(3)
    t.insert(a, pa)
    pb = alloc.alloc(2)
    t.insert(b, np.concatenate([pa[:1], pb]))
    assert t.evictable_pages() == 5
    _, na = t.match_prefix(a)
    t.inc_lock_ref(na)
    assert t.protected_pages() == 3 and t.evictable_pages() == 2
    free0 = alloc.available_pages()
    assert t.evict(100) == 2            # only b's private tail is evictable
    assert alloc.available_pages() == free0 + 2
    assert t.match_prefix(b)[0].size == 1
    t.dec_lock_ref(na)
    assert t.evictable_pages() == 3
    assert t.evict(1) >= 1
    assert t.evict(100) >= 0
    assert t.evictable_pages() == 0 and alloc.available_pages() == 31


@pytest.mark.parametrize("impl", IMPLS)
def test_extra_key_namespaces(impl):
    alloc = PagedAllocator(16, 1)
    t = impl(alloc, 1)
    t.insert([1, 2, 3], alloc.alloc(3), extra_key="lora-a")
    assert t.match_prefix([1, 2, 3], extra_key="lora-b")[0].size == 0
    assert t.match_prefix([1, 2, 3], extra_key="lora-a")[0].size == 3
    assert t.match_prefix([1, 2, 3])[0].size == 0


def test_differential_random():
    rng = np.random.default_rng(0)
    for ps in (1, 2, 4):
        a1, a2 = PagedAllocator(4096, ps), PagedAllocator(4096, ps)
        t1, t2 = RadixCache(a1, ps), RadixCacheCpp(a2, ps)
        locked = []
        for step in range(400):
            base = rng.integers(0, 3, size=rng.integers(1, 6) * ps).tolist()
            op = rng.integers(0, 10)
            if op < 5:
                n = len(base) // ps
                m1, _ = t1.match_prefix(base)
                m2, _ = t2.match_prefix(base)
                assert m1.size == m2.size
                new1, new2 = a1.alloc(n - m1.size), a2.alloc(n - m2.size)
                r1 = t1.insert(base, np.concatenate([m1, new1]))
                r2 = t2.insert(base, np.concatenate([m2, new2]))
                assert r1 == r2
            elif op < 7:
                m1, n1 = t1.match_prefix(base)
                m2, n2 = t2.match_prefix(base)
                assert m1.size == m2.size
                t1.inc_lock_ref(n1); t2.inc_lock_ref(n2)
                locked.append((n1, n2))
            elif op < 8 and locked:
                n1, n2 = locked.pop(rng.integers(0, len(locked)))
                t1.dec_lock_ref(n1); t2.dec_lock_ref(n2)
            else:
                k = int(rng.integers(1, 6))
                assert t1.evict(k) == t2.evict(k)
            assert t1.evictable_pages() == t2.evictable_pages()
            assert t1.protected_pages() == t2.protected_pages()
            assert a1.available_pages() == a2.available_pages()


def test_eviction_policies():
    for pol in ("lru", "lfu", "fifo", "mru", "filo", "priority"):
        alloc = PagedAllocator(16, 1)
        t = RadixCache(alloc, 1, pol)
        t.insert([1], alloc.alloc(1), priority=5)
        t.insert([2], alloc.alloc(1), priority=1)
        t.match_prefix([1]); t.match_prefix([1])
        t.evict(1)
        left = [k for k in ([1], [2]) if t.match_prefix(k)[0].size]
        expect = {"lru": [[1]], "lfu": [[1]], "fifo": [[2]], "mru": [[2]], "filo": [[1]], "priority": [[1]]}[pol]
        assert left == expect, (pol, left)


def test_kv_events():
    alloc = PagedAllocator(16, 1)
    t = RadixCache(alloc, 1, enable_kv_events=True)
    t.insert([1, 2], alloc.alloc(2))
    t.evict(2)
    kinds = [e[0] for e in t.take_kv_events()]
    assert kinds == ["AllBlocksCleared", "BlockStored", "BlockRemoved"]
