"""Diffusion-LLM block decoding: the engine (paged KV, custom block mask on the verify path) against a dense recompute
oracle that runs the same denoising rule on the Hugging Face model with an explicit block-causal 4-D mask."""
import numpy as np
import pytest
import torch

import sglang_b200 as sgl
from sglang_b200.dllm import JointThreshold, LowConfidence

MASK = 511


def _hf(tmp_path):
    import transformers as tf

    torch.manual_seed(0)
    cfg = tf.LlamaConfig(vocab_size=512, hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4,
                         num_key_value_heads=2, max_position_embeddings=256, eos_token_id=2, bos_token_id=1)
    m = tf.LlamaForCausalLM(cfg).eval().float()
    with torch.no_grad():
        for p in m.parameters():
            p.mul_(3.0)                                   # sharper distributions so several positions clear the threshold
    d = tmp_path / "dllm"
    m.save_pretrained(str(d), safe_serialization=True)
    return m, str(d)


def _oracle(hf, prompt, n_new, bs, algo):
    """Full recompute every forward.  ``ends[i]`` = last position token i may attend to (its block's end)."""
    with torch.no_grad():
        first = int(hf(input_ids=torch.tensor([prompt])).logits[0, -1].argmax())
    seq, ends = list(prompt) + [first], list(range(len(prompt) + 1))
    forwards = 0
    while len(seq) - len(prompt) < n_new:
        b = min(bs, n_new - (len(seq) - len(prompt)))
        L = len(seq)
        ids = torch.full((1, b), MASK, dtype=torch.long)
        e = torch.tensor(ends + [L + b - 1] * b)
        allow = torch.arange(L + b)[None, :] <= e[:, None]
        m4 = torch.zeros(1, 1, L + b, L + b).masked_fill(~allow[None, None], float("-inf"))
        state = algo.init_state(ids)
        for _ in range(algo.max_steps(b) + 1):
            with torch.no_grad():
                lg = hf(input_ids=torch.tensor([seq + ids[0].tolist()]), attention_mask=m4).logits[0, L:]
            forwards += 1
            new, fin = algo.step(ids, lg[None], state)
            done = bool(fin[0]) or (not bool((ids == MASK).any()) and bool((new == ids).all()))
            if done:
                break
            ids = new
        seq += ids[0].tolist()
        ends += [L + b - 1] * b
    return seq[len(prompt):], forwards


@pytest.mark.parametrize("algo_name,cfg", [("LowConfidence", {"threshold": 0.3}), ("JointThreshold", {"threshold": 0.3, "edit_threshold": 0.6, "max_post_edit_steps": 2})])
def test_block_diffusion_decode_matches_dense_oracle(tmp_path, algo_name, cfg):
    hf, d = _hf(tmp_path)
    prompts = [np.random.default_rng(s).integers(3, 500, size=n).tolist() for s, n in ((1, 11), (2, 7))]
    import json

    eng = sgl.Engine(model_path=d, device="cpu", page_size=4, skip_tokenizer_init=True, log_level="error", dllm_algorithm=algo_name,
                     dllm_algorithm_config=json.dumps(cfg), dllm_block_size=4, dllm_mask_id=MASK, disable_radix_cache=True)
    outs = eng.generate(input_ids=prompts, sampling_params={"max_new_tokens": 10, "temperature": 0, "ignore_eos": True})
    w = eng.scheduler.spec_worker
    blocks, fwd = w.total_blocks, w.total_forwards
    eng.shutdown()
    cls = {"LowConfidence": LowConfidence, "JointThreshold": JointThreshold}[algo_name]
    for p, o in zip(prompts, outs):
        ref, _ = _oracle(hf, p, 10, 4, cls(MASK, cfg))
        assert o["output_ids"] == ref
        assert MASK not in o["output_ids"]
    assert blocks >= 3 and fwd >= blocks                                   # 9 tokens after the prefill token = 3 blocks per request


def test_low_confidence_rule_commits_above_threshold_or_best_one():
    a = LowConfidence(MASK, {"threshold": 0.9})
    ids = torch.tensor([[MASK, MASK, 7], [MASK, MASK, MASK]])
    lg = torch.full((2, 3, 8), -10.0)
    lg[0, 0, 3] = 10.0          # confident
    lg[0, 1, :] = 0.0           # flat: not committed while another position clears the bar
    lg[1, :, :] = 0.0
    lg[1, 2, 5] = 0.5           # nothing clears 0.9 -> only the best position commits
    new, done = a.step(ids, lg, {})
    assert new.tolist() == [[3, MASK, 7], [MASK, MASK, 5]] and done.tolist() == [False, False]
    _, done = a.step(torch.tensor([[1, 2, 3]]), lg[:1], {})
    assert done.tolist() == [True]


def test_args_validation_and_plugin_registry():
    from sglang_b200.dllm import ALGORITHMS, DllmAlgorithm, DllmConfig, register_algorithm
    from sglang_b200.server_args import ServerArgs

    with pytest.raises(ValueError):
        ServerArgs(model_path="dummy:tiny", device="cpu", dllm_algorithm="LowConfidence", speculative_algorithm="NGRAM")
    a = ServerArgs(model_path="dummy:tiny", device="cpu", dllm_algorithm="LowConfidence", dllm_algorithm_config='{"threshold": 0.5, "block_size": 8}')
    assert a.disable_overlap_schedule
    c = DllmConfig.from_args(a, "SDARForCausalLM")
    assert (c.block_size, c.mask_id, c.algorithm_config) == (8, 151669, {"threshold": 0.5})
    with pytest.raises(ValueError):
        DllmConfig.from_args(a, "LlamaForCausalLM")
    register_algorithm("Mine", type("Mine", (DllmAlgorithm,), {}))
    assert "Mine" in ALGORITHMS
    ALGORITHMS.pop("Mine")
