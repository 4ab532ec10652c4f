"""Prefix caching for state-space / hybrid models through recurrent-state snapshots (mem_cache/mamba_radix.py; reference
mem_cache/mamba_radix_cache.py): a match ends where a snapshot exists, admission restores it, answers equal the uncached engine."""
import numpy as np
import pytest

import sglang_b200 as sgl
from sglang_b200.mem_cache.allocator import PagedAllocator
from sglang_b200.mem_cache.mamba_radix import MambaRadixCache


def test_match_ends_at_the_deepest_snapshot_and_rows_recycle():
    alloc = PagedAllocator(64, 4)
    t = MambaRadixCache(alloc, [100, 101], 4)
    ids = list(range(48))                                  # 12 pages
    t.insert(ids, alloc.alloc(12))
    assert t.match_prefix_mamba(ids)[0].size == 0          # tokens match, but no state to resume from
    r1 = t.alloc_row()
    assert t.attach(ids, 8, r1)                            # snapshot after 8 pages (splits the node)
    pages, node, row = t.match_prefix_mamba(ids)
    assert pages.size == 8 and row == r1 and node.mamba_slot == r1
    assert t.match_prefix_mamba(ids[:28])[0].size == 0     # 7 pages: shorter than the only snapshot
    r2 = t.alloc_row()
    assert t.attach(ids, 12, r2) and t.match_prefix_mamba(ids)[2] == r2
    r3 = t.alloc_row()                                     # no free row: the least recently used snapshot is recycled
    assert r3 in (r1, r2) and t.snapshots_recycled == 1
    t.release_row(r3)
    t.evict(100)                                           # evicting nodes returns their rows
    assert sorted(t.free_rows) == [100, 101] and alloc.available_pages() == 63


@pytest.mark.parametrize("model", ["tiny-mamba2", "tiny-nemotron-h"])
def test_hybrid_engine_reuses_prefixes_through_snapshots(model):
    rng = np.random.default_rng(2)
    shared = rng.integers(5, 500, size=150).tolist()        # > 2 scan boundaries of 64 tokens
    prompts = [shared + rng.integers(5, 500, size=int(n)).tolist() for n in (3, 20, 70)] + [rng.integers(5, 500, size=90).tolist()]
    sp = {"max_new_tokens": 12, "temperature": 0, "ignore_eos": True}
    kw = dict(model_path=f"dummy:{model}", page_size=4, log_level="error", device="cpu", max_running_requests=8)
    ref = sgl.Engine(**kw)
    assert ref.scheduler.tree.disable                       # without snapshot rows the recurrent model cannot share prefixes
    want = [o["output_ids"] for o in ref.generate(input_ids=prompts, sampling_params=sp)]
    ref.shutdown()
    eng = sgl.Engine(mamba_cache_slots=6, **kw)
    got1 = [o["output_ids"] for o in eng.generate(input_ids=prompts[:1], sampling_params=sp)]      # takes a snapshot at token 128
    got2 = [o["output_ids"] for o in eng.generate(input_ids=prompts, sampling_params=sp)]          # the others resume from it
    eng.scheduler.check_invariants()
    tree = eng.scheduler.tree
    assert got1 == want[:1] and got2 == want
    assert tree.snapshots_taken >= 1 and tree.snapshot_hits >= 2 and eng.scheduler.num_cached_tokens >= 2 * 128
    eng.shutdown()


@pytest.mark.parametrize("name", ["mamba1", "jamba", "qwen3-next"])
def test_snapshot_prefix_cache_on_other_recurrent_families(name, tmp_path):
    """Mamba-1 selective scan, Jamba (Mamba-1 + attention + MoE) and Qwen3-Next (gated delta rule) export boundary states too: a tiny random HF
    checkpoint served with snapshot rows answers exactly like the uncached engine while later requests resume from a snapshot"""
    import torch
    import transformers as tf

    from test_hf_parity import CASES

    cfg_cls, model_cls, kw = CASES[name]
    if not hasattr(tf, cfg_cls) or not hasattr(tf, model_cls):
        pytest.skip(f"transformers has no {model_cls}")
    torch.manual_seed(0)
    try:
        hf = getattr(tf, model_cls)(getattr(tf, cfg_cls)(**kw)).float().eval()
    except Exception as e:
        pytest.skip(f"HF {model_cls} unusable here: {e}")
    with torch.no_grad():
        for n, p in hf.named_parameters():
            if p.dim() > 1:
                p.normal_(0.0, 0.08)
    d = str(tmp_path / name)
    hf.save_pretrained(d, safe_serialization=True)
    rng = np.random.default_rng(4)
    shared = rng.integers(5, 500, size=140).tolist()
    prompts = [shared + rng.integers(5, 500, size=int(n)).tolist() for n in (4, 25)] + [rng.integers(5, 500, size=70).tolist()]
    sp = {"max_new_tokens": 8, "temperature": 0, "ignore_eos": True}
    kw2 = dict(model_path=d, device="cpu", page_size=4, skip_tokenizer_init=True, log_level="error", max_running_requests=4)
    ref = sgl.Engine(**kw2)
    want = [o["output_ids"] for o in ref.generate(input_ids=prompts, sampling_params=sp)]
    ref.shutdown()
    eng = sgl.Engine(mamba_cache_slots=4, **kw2)
    got1 = [o["output_ids"] for o in eng.generate(input_ids=prompts[:1], sampling_params=sp)]
    got2 = [o["output_ids"] for o in eng.generate(input_ids=prompts, sampling_params=sp)]
    eng.scheduler.check_invariants()
    t = eng.scheduler.tree
    assert got1 == want[:1] and got2 == want
    assert t.snapshots_taken >= 1 and t.snapshot_hits >= 2 and eng.scheduler.num_cached_tokens >= 2 * 128
    eng.shutdown()
