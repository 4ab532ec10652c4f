"""Mock-model e2e (reference test/registered/mock_model/*): token oracle dictates outputs, KV canary checks every KV slot
read; a perturbation (corrupting the page table) must make the canary fire."""
import numpy as np
import pytest

import sglang_b200 as sgl
from sglang_b200.utils.kv_canary import KVCanaryError, TokenOracle, install_canary


def _run(eng, n=10, seed=0):
    rng = np.random.default_rng(seed)
    base = rng.integers(300, 500, size=16).tolist()
    prompts = [base + rng.integers(300, 500, size=int(k)).tolist() for k in rng.integers(3, 50, size=n)]
    return prompts, eng.generate(input_ids=prompts, sampling_params={"max_new_tokens": 12, "temperature": 0, "ignore_eos": True})


@pytest.mark.parametrize("overlap", [True, False])
def test_oracle_and_canary_clean(overlap):
    eng = sgl.Engine(model_path="dummy:tiny-llama", page_size=4, chunked_prefill_size=32, max_total_tokens=600,
                     disable_overlap_schedule=not overlap, max_running_requests=4)
    oracle = TokenOracle(512, lambda rid, pos: 260 + (pos * 7) % 200)
    oracle.install(eng)
    canary = install_canary(eng, "raise")
    prompts, outs = _run(eng)
    for p, o in zip(prompts, outs):
        want = [260 + ((len(p) + i) * 7) % 200 for i in range(12)]
        assert o["output_ids"] == want
    prompts, outs = _run(eng, seed=0)       # second pass: radix hits, canary still clean
    assert canary.checked > 0 and not canary.violations
    eng.scheduler.check_invariants()


def test_canary_detects_page_table_corruption():
    eng = sgl.Engine(model_path="dummy:tiny-llama", page_size=4, disable_overlap_schedule=True)
    canary = install_canary(eng, "raise")
    inner = eng.runner.forward_batch_generation.__closure__   # the canary wrapper is installed; now corrupt mid-run
    sch = eng.scheduler
    step = {"n": 0}
    orig = sch.run_batch

    def evil(batch):
        step["n"] += 1
        if step["n"] == 4 and batch.forward_mode.is_decode():
            r = batch.reqs[0]
            sch.kv.page_table[r.req_slot, 0] = sch.kv.page_table[r.req_slot, 1]    # alias page 0 to page 1
        return orig(batch)

    sch.run_batch = evil
    with pytest.raises(KVCanaryError):
        eng.generate(input_ids=list(range(300, 330)), sampling_params={"max_new_tokens": 10, "temperature": 0, "ignore_eos": True})
