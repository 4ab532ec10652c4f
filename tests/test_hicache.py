"""HiCache: device eviction -> host tier -> load-back on a later hit; L3 file tier survives a flush (reference: test/registered/hicache/*)."""
import numpy as np

import sglang_b200 as sgl


def test_host_tier_roundtrip_and_l3(tmp_path):
    eng = sgl.Engine(model_path="dummy:tiny-llama", page_size=4, max_total_tokens=256, enable_hierarchical_cache=True,
                     hicache_ratio=4.0, hicache_storage_backend="file", hicache_storage_dir=str(tmp_path), radix_cache_impl="python")
    rng = np.random.default_rng(0)
    sp = {"max_new_tokens": 4, "temperature": 0, "ignore_eos": True}
    A = rng.integers(3, 200, size=100).tolist()
    ref = eng.generate(input_ids=A, sampling_params=sp)
    for _ in range(6):   # push A out of the 256-token device pool
        eng.generate(input_ids=rng.integers(3, 200, size=100).tolist(), sampling_params=sp)
    tree = eng.scheduler.tree
    assert tree.stats["host_writes"] > 0
    again = eng.generate(input_ids=A, sampling_params=sp)
    assert again["output_ids"] == ref["output_ids"]
    assert again["meta_info"]["cached_tokens"] >= 64 and tree.stats["host_hits_pages"] > 0
    eng.scheduler.check_invariants()
    # L3: wipe the tree + device + host state, the file tier still serves the prefix
    assert tree.stats["l3_writes"] > 0
    assert eng.flush_cache()
    tree.host._free = list(range(tree.host.num_pages - 1, -1, -1))
    third = eng.generate(input_ids=A, sampling_params=sp)
    assert third["output_ids"] == ref["output_ids"]
    assert tree.stats["l3_hits_pages"] > 0 and third["meta_info"]["cached_tokens"] > 0
