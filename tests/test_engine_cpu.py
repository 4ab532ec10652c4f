"""BASELINE config #1: radix prefix cache + continuous batching correctness on a 2-layer transformer, CPU, world=1.
Batched/chunked/retracted/cached generation must reproduce one-request-at-a-time greedy outputs token for token
(reference strategy: mock-model e2e + scheduler kits)."""
import numpy as np
import pytest

import sglang_b200 as sgl
from sglang_b200.environ import envs


def _prompts(n, rng, lo=5, hi=60, shared=0):
    base = rng.integers(3, 200, size=shared).tolist()
    return [base + rng.integers(3, 200, size=int(rng.integers(lo, hi))).tolist() for _ in range(n)]


@pytest.fixture(scope="module")
def ref_engine():
    eng = sgl.Engine(model_path="dummy:tiny-llama", page_size=4, disable_radix_cache=True, chunked_prefill_size=-1,
                     disable_overlap_schedule=True, max_running_requests=1)
    yield eng
    eng.shutdown()


def _ref(ref_engine, prompts, max_new):
    return [ref_engine.generate(input_ids=p, sampling_params={"max_new_tokens": m, "temperature": 0, "ignore_eos": True})["output_ids"]
            for p, m in zip(prompts, max_new)]


@pytest.mark.parametrize("overlap", [False, True])
@pytest.mark.parametrize("page_size", [1, 4])
def test_continuous_batching_matches_sequential(ref_engine, overlap, page_size):
    rng = np.random.default_rng(1)
    prompts = _prompts(12, rng, shared=9)
    max_new = rng.integers(1, 20, size=12).tolist()
    want = _ref(ref_engine, prompts, max_new)
    eng = sgl.Engine(model_path="dummy:tiny-llama", page_size=page_size, disable_overlap_schedule=not overlap,
                     chunked_prefill_size=32, max_running_requests=5)
    sps = [{"max_new_tokens": m, "temperature": 0, "ignore_eos": True} for m in max_new]
    got = eng.generate(input_ids=prompts, sampling_params=sps)
    assert [g["output_ids"] for g in got] == want
    # second pass: everything page-aligned is served from the radix cache, outputs unchanged
    got2 = eng.generate(input_ids=prompts, sampling_params=sps)
    assert [g["output_ids"] for g in got2] == want
    assert sum(g["meta_info"]["cached_tokens"] for g in got2) > 0
    eng.scheduler.check_invariants()
    st = eng.scheduler.get_internal_state()
    assert st["available_pages"] == st["total_pages"]    # every page is free or evictable once idle


def test_chunked_prefill_long_prompt(ref_engine):
    rng = np.random.default_rng(2)
    prompts = _prompts(3, rng, lo=150, hi=300)
    want = _ref(ref_engine, prompts, [6] * 3)
    eng = sgl.Engine(model_path="dummy:tiny-llama", page_size=4, chunked_prefill_size=64, scripted_scheduler=True)
    got = eng.generate(input_ids=prompts, sampling_params={"max_new_tokens": 6, "temperature": 0, "ignore_eos": True})
    assert [g["output_ids"] for g in got] == want
    recs = eng.scheduler.batch_records
    assert any(r["chunked_rid"] for r in recs), "a long prompt must have been chunked"
    assert max(r["tokens"] for r in recs if r["mode"] == "EXTEND") <= 64


def test_retract_decode_recovers(ref_engine):
    rng = np.random.default_rng(3)
    prompts = _prompts(6, rng)
    want = _ref(ref_engine, prompts, [24] * 6)
    with envs.SGLB_TEST_RETRACT.override(True), envs.SGLB_TEST_RETRACT_INTERVAL.override(4):
        eng = sgl.Engine(model_path="dummy:tiny-llama", page_size=4)
        got = eng.generate(input_ids=prompts, sampling_params={"max_new_tokens": 24, "temperature": 0, "ignore_eos": True})
    assert eng.scheduler.num_retracted > 0
    assert [g["output_ids"] for g in got] == want
    eng.scheduler.check_invariants()


def test_kv_pressure_eviction_and_oom_retract(ref_engine):
    rng = np.random.default_rng(4)
    prompts = _prompts(10, rng, lo=20, hi=40)
    want = _ref(ref_engine, prompts, [30] * 10)
    eng = sgl.Engine(model_path="dummy:tiny-llama", page_size=4, max_total_tokens=400, schedule_conservativeness=0.3)
    got = eng.generate(input_ids=prompts, sampling_params={"max_new_tokens": 30, "temperature": 0, "ignore_eos": True})
    assert [g["output_ids"] for g in got] == want
    eng.scheduler.check_invariants()


def test_stop_conditions_and_stream():
    eng = sgl.Engine(model_path="dummy:tiny-llama", page_size=4)
    full = eng.generate("abc", {"max_new_tokens": 12, "temperature": 0, "ignore_eos": True})
    ids = full["output_ids"]
    out = eng.generate("abc", {"max_new_tokens": 12, "temperature": 0, "ignore_eos": True, "stop_token_ids": [ids[3]]})
    assert out["output_ids"] == ids[: ids.index(ids[3]) + 1] and out["meta_info"]["finish_reason"]["type"] == "stop"
    chunks = list(eng.generate("abc", {"max_new_tokens": 12, "temperature": 0, "ignore_eos": True}, stream=True))
    assert chunks[-1]["output_ids"] == ids and chunks[-1]["meta_info"]["finish_reason"]["type"] == "length"
    assert len(chunks) >= 2


def test_sampling_seeded_and_logprobs():
    eng = sgl.Engine(model_path="dummy:tiny-llama", page_size=4)
    sp = {"max_new_tokens": 8, "temperature": 0.9, "top_p": 0.9, "top_k": 50, "sampling_seed": 7, "ignore_eos": True}
    a = eng.generate("hello", dict(sp))
    g = eng.generate("hello", {"max_new_tokens": 4, "temperature": 0, "ignore_eos": True}, return_logprob=True, top_logprobs_num=3,
                     logprob_start_len=0)
    lps = g["meta_info"]["output_token_logprobs"]
    assert len(lps) == 4 and all(lp <= 0 for lp, _, _ in lps)
    tops = g["meta_info"]["output_top_logprobs"]
    assert all(abs(t[0][0] - lp[0]) < 1e-4 for t, lp in zip(tops, lps))   # greedy token is the top-1
    assert len(g["meta_info"]["input_token_logprobs"]) == g["meta_info"]["prompt_tokens"]
    assert len(a["output_ids"]) == 8


def test_abort_and_priority():
    eng = sgl.Engine(model_path="dummy:tiny-llama", page_size=4, enable_priority_scheduling=True, max_running_requests=1,
                     scripted_scheduler=True)
    outs = eng.generate(["low low low", "high high high"], {"max_new_tokens": 3, "temperature": 0, "ignore_eos": True},
                        priority=None, rid=["lo", "hi"])
    assert len(outs) == 2
    from sglang_b200.managers import io_struct as ios

    eng.scheduler.handle_request(eng._tokenize(ios.GenerateReqInput(text="to be aborted", rid="ab", sampling_params={"max_new_tokens": 50})))
    eng.abort_request("ab")
    while eng.scheduler.has_work():
        eng.scheduler.step()
    eng.scheduler.check_invariants()
    assert eng.flush_cache()


def test_flush_and_abort_see_the_unmerged_prefill_batch():
    """between a prefill step and the next iteration the batch lives only in scheduler.last_batch.  flush_cache() used to succeed
    there (tree / allocator reset under live requests -> 'dec_lock_ref underflow' kills the scheduler) and abort_request() missed it."""
    from sglang_b200.managers import io_struct as ios

    eng = sgl.Engine(model_path="dummy:tiny-llama", page_size=4, disable_overlap_schedule=True, scripted_scheduler=True)
    sch = eng.scheduler
    sch.handle_request(eng._tokenize(ios.GenerateReqInput(text="the quick brown fox jumps", rid="lb", sampling_params={"max_new_tokens": 40, "ignore_eos": True})))
    sch.step()                                           # prefill ran; the request now sits in last_batch only
    assert sch.last_batch is not None and not sch.running_batch.reqs
    assert eng.flush_cache() is False                    # refused: KV pages are held by the unmerged batch
    sch.step(); sch.step()
    sch.check_invariants()
    # abort in the same window
    sch.handle_request(eng._tokenize(ios.GenerateReqInput(text="another prompt to abort", rid="second", sampling_params={"max_new_tokens": 40, "ignore_eos": True})))
    eng.abort_request("lb")
    for _ in range(10):
        sch.step()
        if sch.last_batch is not None and sch.last_batch.forward_mode.is_extend() and any(r.rid == "second" for r in sch.last_batch.reqs):
            break
    else:
        raise AssertionError("the second request never prefilled")
    eng.abort_request("second")                             # "second" is in last_batch right now
    n = 0
    while sch.has_work():
        sch.step()
        n += 1
        assert n < 20, "aborted requests kept generating"
    sch.check_invariants()
    assert eng.flush_cache() is True
    eng.shutdown()


def test_seeded_sampling_is_scheduling_independent_and_min_new_tokens_holds_under_overlap():
    """a per-request seed must give the same text whatever the batch composition / step the scheduler picks (the RNG counter is the
    request's own output position), and min_new_tokens must suppress EOS exactly that long in overlap mode too."""
    eng = sgl.Engine(model_path="dummy:tiny-llama", page_size=4)
    assert eng.scheduler.enable_overlap
    sp = {"max_new_tokens": 12, "temperature": 1.3, "top_p": 0.95, "top_k": 40, "sampling_seed": 1234, "ignore_eos": True}
    alone = eng.generate("seeded request", dict(sp))["output_ids"]
    eng.flush_cache()
    noise = {"max_new_tokens": 5, "temperature": 1.0, "ignore_eos": True}
    outs = eng.generate(["a b c d e f", "seeded request", "other text here", "zzz"], [noise, dict(sp), noise, dict(noise, max_new_tokens=17)])
    assert outs[1]["output_ids"] == alone
    again = eng.generate("seeded request", dict(sp))["output_ids"]     # radix hit this time: different prefill shape, same stream
    assert again == alone
    # min_new_tokens: make every token an EOS candidate by declaring the greedy continuation a stop token
    g = eng.generate("stop me", {"max_new_tokens": 10, "temperature": 0, "ignore_eos": True})["output_ids"]
    first = g[0]
    o = eng.generate("stop me", {"max_new_tokens": 10, "temperature": 0, "stop_token_ids": [first], "min_new_tokens": 4})
    ids = o["output_ids"]
    assert len(ids) >= 4 and first not in ids[:4], ids
    eng.shutdown()


def test_request_waiting_and_running_timeouts():
    """--request-waiting-timeout aborts what sat in the queue too long; --request-running-timeout ends a generation that runs too long"""
    eng = sgl.Engine(model_path="dummy:tiny-llama", device="cpu", page_size=4, log_level="error", max_running_requests=1,
                     request_waiting_timeout=0.4, request_running_timeout=1.2)
    try:
        eng.scheduler.forward_sleep_time = 0.03
        sp = {"max_new_tokens": 400, "temperature": 0, "ignore_eos": True}
        outs = eng.generate(input_ids=[[5, 6, 7], [8, 9, 10]], sampling_params=sp)
        a, b = outs
        assert a["meta_info"]["finish_reason"]["type"] == "abort" and "ran longer" in a["meta_info"]["finish_reason"]["message"]
        assert 5 < a["meta_info"]["completion_tokens"] < 400                                  # it did run for a while
        assert b["meta_info"]["finish_reason"]["type"] == "abort" and "waited longer" in b["meta_info"]["finish_reason"]["message"]
        assert b["meta_info"]["completion_tokens"] == 0
        eng.scheduler.forward_sleep_time = 0.0
        ok = eng.generate(input_ids=[5, 6, 7], sampling_params={"max_new_tokens": 4, "temperature": 0, "ignore_eos": True})
        assert ok["meta_info"]["finish_reason"]["type"] == "length"
        eng.scheduler.check_invariants()
    finally:
        eng.shutdown()
