"""--enable-mixed-chunk: running requests ride along with prefill chunks as one-token extends (reference ScheduleBatch.mix_with_running /
test/srt/test_chunked_prefill.py mixed variants): same tokens as the plain schedule, fewer forward passes, no KV leak under pressure."""
import numpy as np
import pytest

import sglang_b200 as sgl

RNG = np.random.default_rng(0)
PROMPTS = [RNG.integers(3, 250, size=n).tolist() for n in (5, 9, 70, 3, 120, 33, 7, 64)]
SP = [{"max_new_tokens": m, "temperature": 0, "ignore_eos": True} for m in (20, 12, 6, 30, 5, 9, 25, 8)]


def _run(**kw):
    eng = sgl.Engine(model_path="dummy:tiny-llama", device="cpu", page_size=4, chunked_prefill_size=16, log_level="error", **kw)
    try:
        out = eng.generate(input_ids=PROMPTS, sampling_params=SP)
        eng.scheduler.check_invariants()
        st = eng.scheduler.get_internal_state()
        return [o["output_ids"] for o in out], st, eng.scheduler.forward_ct, eng.scheduler.kv.available_pages(), eng.scheduler.kv.total_pages()
    finally:
        eng.shutdown()


@pytest.mark.parametrize("overlap", [True, False])
def test_mixed_chunk_same_tokens_fewer_steps(overlap):
    ref, _, steps_ref, _, _ = _run(disable_overlap_schedule=not overlap)
    got, st, steps, free, total = _run(enable_mixed_chunk=True, disable_overlap_schedule=not overlap)
    assert got == ref
    assert st["num_mixed_decode_tokens"] > 20 and steps < steps_ref          # decode tokens really rode along with prefill chunks


def test_failed_mixed_attempt_returns_its_pages(monkeypatch):
    """every other mixed attempt is made to fail AFTER the riding requests took the page for their extra token: the step must degrade to a
    plain prefill, hand those pages back (the decode path would otherwise allocate the same column again) and produce the same tokens"""
    ref, _, _, _, _ = _run()
    monkeypatch.setenv("SGLB_STRICT_MEM_CHECK", "1")
    eng = sgl.Engine(model_path="dummy:tiny-llama", device="cpu", page_size=4, chunked_prefill_size=16, log_level="error", enable_mixed_chunk=True)
    try:
        sch, n = eng.scheduler, {"k": 0}
        orig = sch._prepare_for_extend

        def flaky(batch):
            if batch.mixed_ids:
                n["k"] += 1
                if n["k"] % 2 == 1:
                    for r in batch.reqs:
                        if id(r) in batch.mixed_ids:
                            sch.kv.alloc_for_extend(r, r.fill_len)
                    return False
            return orig(batch)

        sch._prepare_for_extend = flaky
        out = eng.generate(input_ids=PROMPTS, sampling_params=SP)
        sch.check_invariants()
        st = sch.get_internal_state()
        assert [o["output_ids"] for o in out] == ref
        assert st["num_mixed_fallbacks"] >= 3 and st["num_mixed_decode_tokens"] > 0
    finally:
        eng.shutdown()


def test_mixed_chunk_leaves_grammar_and_logprob_requests_on_the_decode_path():
    eng = sgl.Engine(model_path="dummy:tiny-llama", device="cpu", page_size=4, chunked_prefill_size=16, enable_mixed_chunk=True, log_level="error")
    try:
        sp = {"max_new_tokens": 10, "temperature": 0, "ignore_eos": True}
        o = eng.generate(input_ids=[PROMPTS[0], PROMPTS[4], PROMPTS[2]], sampling_params=sp, return_logprob=[True, False, False])
        assert len(o[0]["meta_info"]["output_token_logprobs"]) == 10 and all(len(x["output_ids"]) == 10 for x in o)
        eng.scheduler.check_invariants()
    finally:
        eng.shutdown()
    plain = sgl.Engine(model_path="dummy:tiny-llama", device="cpu", page_size=4, chunked_prefill_size=16, log_level="error")
    try:
        p = plain.generate(input_ids=[PROMPTS[0], PROMPTS[4], PROMPTS[2]], sampling_params=sp, return_logprob=[True, False, False])
    finally:
        plain.shutdown()
    assert [x["output_ids"] for x in p] == [x["output_ids"] for x in o]
    assert np.allclose([v for v, _, _ in p[0]["meta_info"]["output_token_logprobs"]], [v for v, _, _ in o[0]["meta_info"]["output_token_logprobs"]], atol=1e-4)
