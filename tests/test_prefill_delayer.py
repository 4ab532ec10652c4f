"""PrefillDelayer / MinFreeSlotsDelayer policy (reference test: test/registered/scheduler/test_prefill_delayer.py)."""
from sglang_b200 import Engine
from sglang_b200.managers.prefill_delayer import MinFreeSlotsDelayer, PrefillDelayer, resolve_min_free_slots


def _mk(rows_seq, **kw):
    it = iter(rows_seq)
    return PrefillDelayer(gather=lambda row: next(it), dp_size=2, dp_attention=True, **kw)


def test_mixed_waits_then_times_out():
    mixed = [[1, 0, 4, 2, 3], [0, 0, 4, 0, 0]]
    d = _mk([mixed] * 5, max_delay_passes=3)
    out = [d.negotiate(True, 0.9, 4, 3, 64) for _ in range(4)]
    assert [o.allow for o in out] == [False, False, True, False]
    assert out[2].reason == "wait_timeout" and out[2].waited_passes == 2


def test_mixed_then_all_is_wait_success_and_watermark_forces():
    mixed = [[1, 0, 1, 2, 3], [0, 0, 1, 0, 0]]
    allp = [[1, 0, 1, 2, 3], [1, 0, 1, 2, 1]]
    d = _mk([mixed, allp], max_delay_passes=10)
    assert not d.negotiate(True, 0.9, 1, 3, 64).allow
    o = d.negotiate(True, 0.9, 1, 3, 64)
    assert o.allow and o.reason == "wait_success" and o.waited_passes == 1
    forced = [[1, 1, 0, 2, 3], [0, 0, 4, 0, 0]]
    o = _mk([forced], max_delay_passes=10, token_usage_low_watermark=0.1).negotiate(True, 0.01, 0, 3, 64)
    assert o.allow and o.reason == "token_watermark"


def test_slot_and_queue_triggers():
    full = [[1, 0, 62, 4, 2], [1, 0, 60, 4, 2]]      # 64 - 62 < 4 -> slot condition
    d = _mk([full] * 4, max_delay_passes=3)
    got = [d.negotiate(True, 0.9, 62, 2, 64).reason for _ in range(4)]
    assert got == ["no_wait", "delay", "delay", "wait_timeout"]     # first one is skipped (skip_first), then bounded delay
    short = [[1, 0, 32, 8, 2], [1, 0, 32, 8, 1]]      # queue 2 < min(32*0.25, 8)
    longq = [[1, 0, 32, 8, 9], [1, 0, 32, 8, 1]]
    d = _mk([short, short, longq], max_delay_passes=100, queue_min_ratio=0.25)
    assert [d.negotiate(True, 0.9, 32, 2, 64).reason for _ in range(3)] == ["no_wait", "delay", "wait_success"]


def test_min_free_slots():
    assert resolve_min_free_slots(None, 64) is None and resolve_min_free_slots(1, 64) is None and resolve_min_free_slots(100, 8) == 8
    m = MinFreeSlotsDelayer(4)
    assert m.should_delay(10, 3) and not m.should_delay(10, 4) and not m.should_delay(0, 0)


def test_engine_with_delayers_matches_plain():
    prompts = [[5, 6, 7, 8], [9, 10, 11], [3] * 20, [4, 5], [6] * 9, [8, 9, 10]]
    sp = [{"max_new_tokens": 3 + 2 * i, "temperature": 0, "ignore_eos": True} for i in range(len(prompts))]
    e = Engine(model_path="dummy:tiny-llama", device="cpu", page_size=4)
    ref = [o["output_ids"] for o in e.generate(input_ids=prompts, sampling_params=sp)]
    e.shutdown()
    e = Engine(model_path="dummy:tiny-llama", device="cpu", page_size=4, max_running_requests=4, enable_prefill_delayer=True,
               prefill_delayer_max_delay_passes=4, prefill_delayer_queue_min_ratio=0.5)
    got = [o["output_ids"] for o in e.generate(input_ids=prompts, sampling_params=sp)]
    st = e.scheduler.prefill_delayer.stats
    e.shutdown()
    assert got == ref
    assert st["delay"] > 0 and (st["wait_timeout"] + st["wait_success"]) > 0
    e = Engine(model_path="dummy:tiny-llama", device="cpu", page_size=4, max_running_requests=4, min_free_slots=2)
    got = [o["output_ids"] for o in e.generate(input_ids=prompts, sampling_params=sp)]
    e.shutdown()
    assert got == ref
