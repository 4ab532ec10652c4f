"""HiCache: device eviction -> host tier -> load-back on a later hit; L3 file tier survives a flush (reference: test/registered/hicache/*)."""
import numpy as np

import sglang_b200 as sgl


import pytest


@pytest.mark.parametrize("backend,layout", [("file", "layer_first"), ("shm", "layer_first"), ("memory", "page_first"), ("remote", "layer_first"),
                                            ("file", "page_first")])
def test_host_tier_roundtrip_and_l3(tmp_path, backend, layout):
    import json

    from sglang_b200.mem_cache.hicache_storage import PageServer

    extra, srv = {}, None
    if backend == "remote":
        srv = PageServer("127.0.0.1", 0).start()
        extra = {"port": srv.port}
    elif backend == "shm":
        extra = {"name": f"sglb_test_{tmp_path.name}"}
    eng = sgl.Engine(model_path="dummy:tiny-llama", page_size=4, max_total_tokens=256, enable_hierarchical_cache=True,
                     hicache_ratio=4.0, hicache_storage_backend=backend, hicache_storage_dir=str(tmp_path) if backend == "file" else None,
                     hicache_storage_backend_extra_config=json.dumps(extra), radix_cache_impl="python", hicache_mem_layout=layout)
    assert eng.scheduler.tree.host.layout == layout
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
    assert tree.storage.stats().get("backend", backend) in (backend, "memory")
    if backend == "remote":
        # a SECOND engine sharing the page server gets the prefix without ever having computed it
        eng2 = sgl.Engine(model_path="dummy:tiny-llama", page_size=4, max_total_tokens=256, enable_hierarchical_cache=True, hicache_ratio=4.0,
                          hicache_storage_backend=f"remote://127.0.0.1:{srv.port}", radix_cache_impl="python")
        o = eng2.generate(input_ids=A, sampling_params=sp)
        assert o["output_ids"] == ref["output_ids"] and o["meta_info"]["cached_tokens"] > 0
        eng2.shutdown()
        srv.shutdown()
    if backend == "shm":
        import shutil

        shutil.rmtree(tree.storage.root, ignore_errors=True)
    eng.shutdown()


def test_runtime_attach_detach_storage_backend():
    from fastapi.testclient import TestClient

    from sglang_b200.entrypoints.http_server import create_app

    eng = sgl.Engine(model_path="dummy:tiny-llama", page_size=4, max_total_tokens=256, enable_hierarchical_cache=True, hicache_ratio=2.0,
                     radix_cache_impl="python", log_level="error")
    with TestClient(create_app(eng)) as c:
        assert c.get("/hicache/storage-backend").json()["attached"] is False
        assert c.put("/hicache/storage-backend", json={"backend": "memory", "config": {"capacity_pages": 64}}).json()["success"]
        assert c.get("/hicache/storage-backend").json() == {"attached": True, "stats": {"backend": "memory", "pages": 0, "hits": 0, "misses": 0}}
        assert c.put("/hicache/storage-backend", json={"backend": "nope"}).status_code == 400
        assert c.delete("/hicache/storage-backend").json()["success"]
        assert c.get("/hicache/storage-backend").json()["attached"] is False
    eng.shutdown()


def test_write_policies():
    """write_through backs every prefix up at insert; write_through_selective only prefixes that were hit again; write_back only at eviction"""
    rng = np.random.default_rng(1)
    sp = {"max_new_tokens": 2, "temperature": 0, "ignore_eos": True}
    A = rng.integers(3, 200, size=40).tolist()
    writes = {}
    for pol in ("write_through", "write_through_selective", "write_back"):
        eng = sgl.Engine(model_path="dummy:tiny-llama", page_size=4, max_total_tokens=4096, enable_hierarchical_cache=True, hicache_ratio=2.0,
                         hicache_write_policy=pol, radix_cache_impl="python", log_level="error")
        eng.generate(input_ids=A, sampling_params=sp)
        first = eng.scheduler.tree.stats["host_writes"]
        eng.generate(input_ids=A + [7, 8, 9, 10, 11, 12, 13, 14], sampling_params=sp)       # hits the A prefix twice over
        eng.generate(input_ids=A + [20, 21, 22, 23, 24, 25, 26, 27], sampling_params=sp)
        writes[pol] = (first, eng.scheduler.tree.stats["host_writes"])
        eng.shutdown()
    assert writes["write_through"][0] > 0
    assert writes["write_through_selective"][0] == 0 and writes["write_through_selective"][1] > 0
    assert writes["write_back"] == (0, 0)


@pytest.mark.parametrize("layout", ["layer_first", "page_first"])
def test_mla_latent_pool_is_one_plane_and_survives_the_host_tier(layout):
    """DeepSeek MLA: the KV pool is ONE 576-wide latent plane (value = column view of the same rows); evicting to the host tier and loading
    back (both host layouts) reproduces the generation"""
    import numpy as np

    rng = np.random.default_rng(3)
    A = rng.integers(3, 250, size=61).tolist()
    fillers = [rng.integers(3, 250, size=70).tolist() for _ in range(4)]
    sp = {"max_new_tokens": 6, "temperature": 0, "ignore_eos": True}
    eng = sgl.Engine(model_path="dummy:tiny-deepseek-v3", device="cpu", page_size=4, max_total_tokens=200, enable_hierarchical_cache=True,
                     hicache_ratio=4.0, radix_cache_impl="python", hicache_mem_layout=layout, log_level="error")
    try:
        pool = eng.runner.kv_pool
        cfg = eng.model_config
        assert pool.planes == 1 and pool.buf.shape[1] == 1 and pool.buf.shape[-1] == cfg.kv_lora_rank + cfg.qk_rope_head_dim
        assert pool.v(0).shape[-1] == cfg.kv_lora_rank and pool.v(0).data_ptr() == pool.k(0).data_ptr()
        ref = eng.generate(input_ids=A, sampling_params=sp)
        for f in fillers:                       # push A's pages out of the device pool
            eng.generate(input_ids=f, sampling_params=sp)
        again = eng.generate(input_ids=A, sampling_params=sp)
        tree = eng.scheduler.kv.tree
        assert again["output_ids"] == ref["output_ids"]
        assert tree.stats["host_writes"] > 0 and tree.stats["host_hits_pages"] > 0 and again["meta_info"]["cached_tokens"] > 0
    finally:
        eng.shutdown()
