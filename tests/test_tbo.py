"""Two-batch overlap: prefill and decode batches are split into two micro-batches that walk the layers alternately; results equal the plain
forward (single process: dense + MoE; multi-process DP-attention + EP all-to-all runs in test_parallel_cpu)."""
import numpy as np

import sglang_b200 as sgl


def _run(preset, **kw):
    rng = np.random.default_rng(7)
    prompts = [rng.integers(3, 400, size=n).tolist() for n in (17, 4, 9, 26, 6)]
    sp = {"max_new_tokens": 6, "temperature": 0, "ignore_eos": True}
    e = sgl.Engine(model_path=f"dummy:{preset}", device="cpu", page_size=4, log_level="error", **kw)
    outs = e.generate(input_ids=prompts, sampling_params=sp, return_logprob=True, logprob_start_len=0)
    n = getattr(e.runner, "tbo_forwards", 0)
    e.scheduler.check_invariants()
    e.shutdown()
    return outs, n


def test_tbo_matches_plain_forward():
    for preset in ("tiny-llama", "tiny-mixtral", "tiny-gemma2", "tiny-qwen3-moe"):
        ref, n0 = _run(preset)
        got, n1 = _run(preset, enable_two_batch_overlap=True)
        assert n0 == 0 and n1 > 3, preset               # prefill batches and every multi-request decode step were split
        for a, b in zip(ref, got):
            assert a["output_ids"] == b["output_ids"], preset
            la = [v for v, _, _ in a["meta_info"]["input_token_logprobs"][1:]]
            lb = [v for v, _, _ in b["meta_info"]["input_token_logprobs"][1:]]
            assert max(abs(x - y) for x, y in zip(la, lb)) < 1e-4, preset
