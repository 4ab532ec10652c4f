"""Sliding-window KV management (mem_cache/swa.py): tombstoning prefix tree, two page tables, sliding release — and the engine on hybrid
sliding-window models giving the same tokens with and without the second pool (reference: test/registered/... swa radix cache / hybrid model tests)."""
import numpy as np
import pytest

import sglang_b200 as sgl
from sglang_b200.mem_cache.allocator import PagedAllocator
from sglang_b200.mem_cache.swa import SWARadixCache


def _tree(ps=4, window=8):
    full, swa = PagedAllocator(64, ps), PagedAllocator(32, ps)
    return SWARadixCache(full, swa, window, ps), full, swa


def test_match_needs_an_intact_trailing_window():
    t, full, swa = _tree(ps=4, window=8)          # need_pages = 8/4 + 1 = 3
    ids = list(range(100, 140))                   # 10 pages
    fp, sp = full.alloc(10), swa.alloc(10)
    sp_t = sp.copy(); sp_t[:4] = 0                # the first four pages slid out before the request finished
    swa.free(sp[:4])
    matched, adopted = t.insert_swa(ids, fp, sp_t)
    assert matched == 0 and adopted.size == 0
    pages, swap, node = t.match_prefix_swa(ids)
    assert pages.size == 10 and (swap[4:] == sp[4:]).all() and (swap[:4] == 0).all()
    # a shorter query ending inside the tombstoned region is not reusable by the sliding layers ...
    pages, swap, _ = t.match_prefix_swa(ids[:16])          # 4 pages, all sliding pages gone
    assert pages.size == 0
    # ... while 7 pages end on three intact sliding pages (4, 5, 6)
    pages, swap, node7 = t.match_prefix_swa(ids[:28])
    assert pages.size == 7 and (swap[4:7] > 0).all()
    # reclaim sliding pages under pressure: full pages stay, the prefix stops being reusable
    freed = t.evict_swa(3)
    assert freed >= 3 and t.evictable_pages() == 10
    assert swa.available_pages() == 31 - 6 + freed
    pages, _, _ = t.match_prefix_swa(ids)
    assert pages.size < 10


def test_insert_heals_tombstones_and_reports_adopted_pages():
    t, full, swa = _tree(ps=4, window=8)
    ids = list(range(40))
    fp, sp = full.alloc(10), swa.alloc(10)
    t.insert_swa(ids, fp, sp)
    t.evict_swa(10)                                    # everything tombstoned
    assert t.match_prefix_swa(ids)[0].size == 0
    fp2, sp2 = full.alloc(10), swa.alloc(10)
    sp2_t = sp2.copy(); sp2_t[:6] = 0; swa.free(sp2[:6])
    matched, adopted = t.insert_swa(ids, fp2, sp2_t)
    assert matched == 10 and adopted.tolist() == [False] * 6 + [True] * 4
    full.free(fp2)                                     # the caller's duplicates
    pages, swap, _ = t.match_prefix_swa(ids)
    assert pages.size == 10 and (swap[6:] == sp2[6:]).all()
    # evicting the leaf gives both kinds of pages back
    t.evict(100)
    assert full.available_pages() == 63 and swa.available_pages() == 31


@pytest.mark.parametrize("model", ["tiny-gemma2", "tiny-gpt-oss"])
def test_engine_with_sliding_pool_matches_single_pool(model):
    rng = np.random.default_rng(3)
    shared = rng.integers(5, 500, size=70).tolist()
    prompts = [shared + rng.integers(5, 500, size=int(n)).tolist() for n in (3, 9, 40)] + [rng.integers(5, 500, size=int(n)).tolist() for n in (5, 33, 120)]
    sp = {"max_new_tokens": 24, "temperature": 0, "ignore_eos": True}
    kw = dict(model_path=f"dummy:{model}", page_size=4, chunked_prefill_size=32, log_level="error", device="cpu", max_running_requests=8)
    ref = sgl.Engine(**kw)
    want = [o["output_ids"] for o in ref.generate(input_ids=prompts, sampling_params=sp)]
    want2 = [o["output_ids"] for o in ref.generate(input_ids=prompts[:3], sampling_params=sp)]      # second round: prefix hits
    ref.shutdown()
    eng = sgl.Engine(swa_full_tokens_ratio=0.5, **kw)
    assert eng.scheduler.swa is not None and eng.runner.kv_pool.swa.num_pages < eng.runner.kv_pool.full.num_pages
    got = [o["output_ids"] for o in eng.generate(input_ids=prompts, sampling_params=sp)]
    eng.scheduler.check_invariants()
    got2 = [o["output_ids"] for o in eng.generate(input_ids=prompts[:3], sampling_params=sp)]
    eng.scheduler.check_invariants()
    kv = eng.scheduler.kv
    assert got == want and got2 == want2
    assert kv.swa_slid_pages > 0, "long requests must have returned sliding pages while running"
    assert eng.scheduler.num_cached_tokens > 0, "the second round must reuse cached prefixes"
    eng.shutdown()


def test_engine_sliding_pool_under_memory_pressure():
    """tiny pools: the sliding allocator has to tombstone cached prefixes (and the scheduler to retract) while answers stay the same"""
    rng = np.random.default_rng(5)
    prompts = [rng.integers(5, 500, size=int(n)).tolist() for n in rng.integers(20, 90, size=14)]
    sp = {"max_new_tokens": 30, "temperature": 0, "ignore_eos": True}
    kw = dict(model_path="dummy:tiny-gemma2", page_size=4, chunked_prefill_size=32, log_level="error", device="cpu", max_running_requests=6)
    ref = sgl.Engine(**kw)
    want = [o["output_ids"] for o in ref.generate(input_ids=prompts, sampling_params=sp)]
    ref.shutdown()
    eng = sgl.Engine(swa_full_tokens_ratio=0.3, max_total_tokens=700, **kw)
    for _ in range(2):
        got = [o["output_ids"] for o in eng.generate(input_ids=prompts, sampling_params=sp)]
        eng.scheduler.check_invariants()
        assert got == want
    assert eng.scheduler.tree.swa_tombstoned > 0 or eng.scheduler.kv.swa_slid_pages > 0
    eng.shutdown()


def test_request_that_can_never_fit_is_rejected_not_retried_forever():
    """a sliding pool of a handful of pages: prompts that fit the admission check but can never be placed must come back as aborted
    (the scheduler used to spin at 20 k iterations / s without progress)"""
    rng = np.random.default_rng(8)
    prompts = [rng.integers(5, 500, size=int(n)).tolist() for n in (20, 24, 22)]
    eng = sgl.Engine(model_path="dummy:tiny-gemma2", page_size=4, log_level="error", device="cpu", max_running_requests=4, swa_full_tokens_ratio=0.5,
                     max_total_tokens=40)          # 10 full pages, 8 sliding pages
    outs = eng.generate(input_ids=prompts, sampling_params={"max_new_tokens": 30, "temperature": 0, "ignore_eos": True})
    kinds = [o["meta_info"]["finish_reason"]["type"] for o in outs]
    assert len(outs) == 3 and all(k in ("length", "abort") for k in kinds) and "abort" in kinds
    assert any("can not be scheduled" in (o["meta_info"]["finish_reason"].get("message") or "") for o in outs)
    # the engine stays usable: a request that fits still runs
    ok = eng.generate(input_ids=[prompts[0][:8]], sampling_params={"max_new_tokens": 4, "temperature": 0, "ignore_eos": True})
    assert ok[0]["meta_info"]["finish_reason"]["type"] == "length"
    eng.scheduler.check_invariants()
    eng.shutdown()
