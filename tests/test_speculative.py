"""Speculative decoding: greedy outputs identical to plain decoding; drafts get accepted on repetitive text
(reference: test/registered/spec/* + spec_decoding_kit accept-length gate)."""
import numpy as np
import pytest

import sglang_b200 as sgl


def _prompts():
    rng = np.random.default_rng(0)
    base = rng.integers(3, 200, size=12).tolist()
    return [base * 4, rng.integers(3, 200, size=30).tolist(), (base[:5] * 9)[:40]]


@pytest.mark.parametrize("algo", ["NGRAM", "STANDALONE"])
def test_spec_greedy_equals_plain(algo):
    prompts = _prompts()
    sp = {"max_new_tokens": 40, "temperature": 0, "ignore_eos": True}
    plain = sgl.Engine(model_path="dummy:tiny-llama", page_size=4)
    want = [o["output_ids"] for o in plain.generate(input_ids=prompts, sampling_params=sp)]
    kw = dict(speculative_draft_model_path="dummy:tiny-llama") if algo == "STANDALONE" else {}
    eng = sgl.Engine(model_path="dummy:tiny-llama", page_size=4, speculative_algorithm=algo, speculative_num_draft_tokens=4, **kw)
    outs = eng.generate(input_ids=prompts, sampling_params=sp)
    assert [o["output_ids"] for o in outs] == want
    w = eng.scheduler.spec_worker
    assert w.total_verify > 0
    if algo == "STANDALONE":
        # the draft model IS the target model (same seed offset differs -> different weights), acceptance is not guaranteed
        assert w.total_drafted > 0
    m = outs[0]["meta_info"]
    assert m["spec_verify_ct"] > 0 and m["spec_accept_length"] >= 1.0
    eng.scheduler.check_invariants()


def test_ngram_accepts_on_repetition():
    # a model that has already generated this continuation once: the corpus must predict it the second time
    eng = sgl.Engine(model_path="dummy:tiny-llama", page_size=4, speculative_algorithm="NGRAM", speculative_num_draft_tokens=6,
                     disable_radix_cache=True)
    p = np.random.default_rng(5).integers(3, 200, size=20).tolist()
    sp = {"max_new_tokens": 48, "temperature": 0, "ignore_eos": True}
    a = eng.generate(input_ids=p, sampling_params=sp)
    w = eng.scheduler.spec_worker
    v1 = w.total_verify
    b = eng.generate(input_ids=p, sampling_params=sp)
    assert a["output_ids"] == b["output_ids"]
    steps2 = w.total_verify - v1
    assert steps2 < 48 * 0.6, f"second pass should need far fewer target steps ({steps2})"
    assert b["meta_info"]["spec_accept_length"] > 1.5


def test_eagle_chain_drafts_exact_and_feature_window():
    """EAGLE worker: outputs identical to plain greedy decoding whatever the draft quality; features tapped from prefill + verify"""
    prompts = _prompts()
    sp = {"max_new_tokens": 24, "temperature": 0, "ignore_eos": True}
    plain = sgl.Engine(model_path="dummy:tiny-llama", page_size=4)
    want = [o["output_ids"] for o in plain.generate(input_ids=prompts, sampling_params=sp)]
    eng = sgl.Engine(model_path="dummy:tiny-llama", page_size=4, speculative_algorithm="EAGLE", speculative_num_draft_tokens=3,
                     chunked_prefill_size=16)
    outs = eng.generate(input_ids=prompts, sampling_params=sp)
    assert [o["output_ids"] for o in outs] == want
    w = eng.scheduler.spec_worker
    assert type(w).__name__ == "EagleSpecWorker" and w.total_verify > 0 and w.total_drafted > 0
    eng.scheduler.check_invariants()
    # a draft that IS a perfect predictor: make the draft layer output the target feature of the next position by construction is not
    # possible with random weights, so check the mechanics instead — drafts are proposed from step one and bounded by num_draft
    assert w.total_drafted <= w.total_verify * 3


def test_ngram_tree_drafts_exact_and_kv_compaction():
    """token-TREE drafts (branch 3): verification uses a tree mask, the accepted path is often not the first branch, so its KV rows are
    compacted; greedy output must still equal plain decoding, also after the same prompts ran once (corpus full of branching history)"""
    prompts = _prompts() + [[7, 8, 9, 10] * 6 + [7, 8, 11, 12] * 3 + [7, 8]]
    sp = {"max_new_tokens": 40, "temperature": 0, "ignore_eos": True}
    plain = sgl.Engine(model_path="dummy:tiny-llama", page_size=4)
    want = [o["output_ids"] for o in plain.generate(input_ids=prompts, sampling_params=sp)]
    plain.shutdown()
    eng = sgl.Engine(model_path="dummy:tiny-llama", page_size=4, speculative_algorithm="NGRAM", speculative_num_draft_tokens=8,
                     speculative_eagle_topk=3, disable_radix_cache=True)
    for _ in range(2):
        outs = eng.generate(input_ids=prompts, sampling_params=sp)
        assert [o["output_ids"] for o in outs] == want
    w = eng.scheduler.spec_worker
    assert w.total_tree_verifies > 0 and w.total_accepted > 0
    # an oracle that hides the true continuation in the SECOND branch at every depth: the accepted path is never the leading slots,
    # so every accepted token's KV row has to be moved
    truth = {tuple(p): o for p, o in zip(prompts, want)}

    def oracle_tree(req, max_nodes):
        o = truth[tuple(req.origin_input_ids)]
        k = len(req.output_ids)
        nxt = o[k: k + 3]
        toks, par, parent = [], [], -1
        for t in nxt:
            if len(toks) + 2 > max_nodes:
                break
            toks += [(t + 1) % 500, t]                 # wrong sibling first
            par += [parent, parent]
            parent = len(toks) - 1
        return toks, par

    w.propose_tree = oracle_tree
    acc0 = w.total_accepted
    outs = eng.generate(input_ids=prompts, sampling_params=sp)
    assert [o["output_ids"] for o in outs] == want
    assert w.total_kv_moves > 0 and w.total_accepted - acc0 > 20
    eng.scheduler.check_invariants()
    eng.shutdown()


def test_eagle_tree_drafts_exact():
    """EAGLE with topk 3: the draft layer expands the best frontier nodes under a tree mask; the verified output is exact and the tree is
    a proper tree (several children per node, depth > 1)"""
    prompts = _prompts()
    sp = {"max_new_tokens": 24, "temperature": 0, "ignore_eos": True}
    plain = sgl.Engine(model_path="dummy:tiny-llama", page_size=4)
    want = [o["output_ids"] for o in plain.generate(input_ids=prompts, sampling_params=sp)]
    plain.shutdown()
    eng = sgl.Engine(model_path="dummy:tiny-llama", page_size=4, speculative_algorithm="EAGLE", speculative_num_draft_tokens=8,
                     speculative_eagle_topk=3, speculative_num_steps=3)
    w = eng.scheduler.spec_worker
    seen = []
    orig = w.propose_tree

    def spy(req, n):
        t, p = orig(req, n)
        seen.append((list(t), list(p)))
        return t, p

    w.propose_tree = spy
    outs = eng.generate(input_ids=prompts, sampling_params=sp)
    assert [o["output_ids"] for o in outs] == want
    full = [p for t, p in seen if len(t) == 8]
    assert full, "trees of the requested size were proposed"
    assert any(p.count(-1) == 3 for p in full) and any(max(p) > 0 for p in full)        # 3 children of the root, deeper levels exist
    assert w.total_tree_verifies > 0
    eng.scheduler.check_invariants()
    eng.shutdown()
