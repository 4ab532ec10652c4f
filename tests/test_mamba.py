"""Mamba-2 / hybrid decoders: SSD chunk scan == step recurrence; the engine's chunked prefill + decode == one-shot prefill of the whole
sequence (state carried in the MambaPool across chunks and decode steps); batching does not mix states."""
import numpy as np
import torch

import sglang_b200 as sgl
from sglang_b200.models.mamba import ssd_chunk_scan, ssm_step


def test_ssd_chunk_scan_equals_recurrence():
    torch.manual_seed(0)
    T, H, P, N = 37, 4, 8, 16
    x, dt = torch.randn(T, H, P), torch.rand(T, H) * 0.5 + 0.01
    A = -(torch.rand(H) * 2 + 0.2)
    B, C = torch.randn(T, H, N), torch.randn(T, H, N)
    st0 = torch.randn(H, P, N)
    st = st0.clone()[None]
    ref = []
    for t in range(T):
        ref.append(ssm_step(x[t:t + 1], dt[t:t + 1], A, B[t:t + 1], C[t:t + 1], st)[0])
    ref = torch.stack(ref)
    for chunk in (4, 8, 64):
        y, fin = ssd_chunk_scan(x, dt, A, B, C, st0, chunk)
        torch.testing.assert_close(y, ref, atol=2e-4, rtol=2e-4)
        torch.testing.assert_close(fin, st[0], atol=2e-4, rtol=2e-4)


def _lp(o):
    return [v for v, _, _ in o["meta_info"]["output_token_logprobs"]]


def _run(preset, **kw):
    rng = np.random.default_rng(3)
    prompts = [rng.integers(3, 400, size=n).tolist() for n in (21, 5, 33, 9)]
    sp = {"max_new_tokens": 8, "temperature": 0, "ignore_eos": True}
    e = sgl.Engine(model_path=f"dummy:{preset}", device="cpu", page_size=4, **kw)
    outs = e.generate(input_ids=prompts, sampling_params=sp, return_logprob=True)
    assert e.scheduler.tree.disabled if hasattr(e.scheduler.tree, "disabled") else True
    e.scheduler.check_invariants()
    return e, prompts, outs


def test_hybrid_engine_chunked_and_batched_equals_one_by_one():
    for preset in ("tiny-nemotron-h", "tiny-mamba2"):
        e, prompts, outs = _run(preset, chunked_prefill_size=8, max_running_requests=3)       # chunks of 8, queueing, mixed batches
        e.shutdown()
        e1 = sgl.Engine(model_path=f"dummy:{preset}", device="cpu", page_size=4, chunked_prefill_size=512, max_running_requests=1)
        sp = {"max_new_tokens": 8, "temperature": 0, "ignore_eos": True}
        for p, o in zip(prompts, outs):
            r = e1.generate(input_ids=p, sampling_params=sp, return_logprob=True)
            assert r["output_ids"] == o["output_ids"], preset
            assert max(abs(a - b) for a, b in zip(_lp(r), _lp(o))) < 2e-3
        # teacher forcing: prompt + generated tokens prefilled in ONE pass must reproduce the decode-time logprobs of those tokens
        p, o = prompts[0], outs[0]
        full = e1.generate(input_ids=p + o["output_ids"], sampling_params={"max_new_tokens": 1, "temperature": 0}, return_logprob=True,
                           logprob_start_len=len(p) - 1)
        tf = [v for v, _, _ in full["meta_info"]["input_token_logprobs"][1:]]
        assert max(abs(a - b) for a, b in zip(tf, _lp(o))) < 2e-3, preset
        e1.shutdown()


def test_retraction_reprefills_state():
    """a retracted request is re-prefilled from position 0: its recurrent state must be rebuilt, not reused"""
    import os

    e, prompts, outs = _run("tiny-nemotron-h", chunked_prefill_size=16)
    e.shutdown()
    os.environ["SGLB_TEST_RETRACT"] = "1"
    try:
        from sglang_b200.environ import envs

        envs.SGLB_TEST_RETRACT.set(True)
        e2, _, outs2 = _run("tiny-nemotron-h", chunked_prefill_size=16)
        assert [o["output_ids"] for o in outs2] == [o["output_ids"] for o in outs]
        e2.shutdown()
    finally:
        os.environ.pop("SGLB_TEST_RETRACT", None)
        envs.SGLB_TEST_RETRACT.set(False)


def test_gated_delta_chunked_equals_recurrence():
    from sglang_b200.models.qwen3_next import gated_delta_chunked, gated_delta_step

    torch.manual_seed(0)
    T, H, Dk, Dv = 150, 3, 8, 16
    q, k, v = torch.randn(T, H, Dk), torch.randn(T, H, Dk), torch.randn(T, H, Dv)
    g, beta = -torch.rand(T, H) * 0.5, torch.rand(T, H)
    st0 = torch.randn(H, Dk, Dv) * 0.3
    st = st0.clone()[None]
    ref = torch.stack([gated_delta_step(q[t:t + 1], k[t:t + 1], v[t:t + 1], g[t:t + 1], beta[t:t + 1], st)[0] for t in range(T)])
    for chunk in (16, 64):
        o, fin = gated_delta_chunked(q, k, v, g, beta, st0, chunk)
        torch.testing.assert_close(o, ref, atol=2e-4, rtol=2e-4)
        torch.testing.assert_close(fin, st[0], atol=2e-4, rtol=2e-4)
