"""Transformers fallback (models/transformers_fallback.py): architectures without a native class run the Hugging Face module tree with the
engine's paged attention plugged in; logprobs must match plain Hugging Face (reference ``test/srt/models/test_transformers_models.py``)."""
import numpy as np
import pytest
import torch
import transformers as tf

import sglang_b200 as sgl

BASE = dict(hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2, vocab_size=512,
            max_position_embeddings=512)
CASES = {
    "solar-open": ("SolarOpenConfig", "SolarOpenForCausalLM", dict(BASE), "auto"),
    "bitnet": ("BitNetConfig", "BitNetForCausalLM", dict(BASE), "auto"),
    "afmoe": ("AfmoeConfig", "AfmoeForCausalLM", dict(BASE), "auto"),
    "llama-forced": ("LlamaConfig", "LlamaForCausalLM", dict(BASE), "transformers"),
}


def _save(name, tmp_path):
    cfg_cls, model_cls, kw, impl = CASES[name]
    if not hasattr(tf, cfg_cls) or not hasattr(tf, model_cls):
        pytest.skip(f"transformers has no {model_cls}")
    torch.manual_seed(0)
    hf = getattr(tf, model_cls)(getattr(tf, cfg_cls)(**kw)).float().eval()
    with torch.no_grad():
        for n, p in hf.named_parameters():
            if p.dim() > 1:
                p.normal_(0.0, 0.08)
    d = tmp_path / name
    hf.save_pretrained(str(d), safe_serialization=True)
    return hf, str(d), impl


@pytest.mark.parametrize("name", list(CASES))
def test_fallback_matches_hf(name, tmp_path):
    hf, d, impl = _save(name, tmp_path)
    ids = np.random.default_rng(1).integers(3, 500, size=29).tolist()
    eng = sgl.Engine(model_path=d, device="cpu", page_size=4, skip_tokenizer_init=True, chunked_prefill_size=16, log_level="error", model_impl=impl)
    try:
        assert eng.model_config.architecture == "TransformersForCausalLM"
        o = eng.generate(input_ids=ids, sampling_params={"max_new_tokens": 6, "temperature": 0, "ignore_eos": True}, return_logprob=True,
                         logprob_start_len=0)
        # a second request over the same prefix takes the radix-cache path: same continuation, prefix served from the paged pool
        a = eng.generate(input_ids=ids, sampling_params={"max_new_tokens": 6, "temperature": 0, "ignore_eos": True})
        assert a["output_ids"] == o["output_ids"] and a["meta_info"]["cached_tokens"] >= 24
    finally:
        eng.shutdown()
    full = ids + o["output_ids"]
    with torch.no_grad():
        lp = torch.log_softmax(hf(input_ids=torch.tensor([full])).logits[0].float(), -1)
    ref = [float(lp[t, full[t + 1]]) for t in range(len(ids) - 1)]
    got = [v for v, _, _ in o["meta_info"]["input_token_logprobs"][1:]]
    assert max(abs(a - b) for a, b in zip(ref, got)) < 3e-3
    for j, tok in enumerate(o["output_ids"]):
        assert abs(float(lp[len(ids) - 1 + j, tok]) - o["meta_info"]["output_token_logprobs"][j][0]) < 3e-3


def test_model_impl_sglang_refuses_unknown_architecture(tmp_path):
    _, d, _ = _save("solar-open", tmp_path)
    with pytest.raises(Exception, match="unsupported architecture"):
        sgl.Engine(model_path=d, device="cpu", skip_tokenizer_init=True, log_level="error", model_impl="sglang")


def test_dummy_weights_through_fallback(tmp_path):
    _, d, _ = _save("solar-open", tmp_path)
    eng = sgl.Engine(model_path=d, device="cpu", page_size=4, skip_tokenizer_init=True, load_format="dummy", log_level="error")
    try:
        o = eng.generate(input_ids=[5, 6, 7, 8], sampling_params={"max_new_tokens": 4, "temperature": 0, "ignore_eos": True})
        assert len(o["output_ids"]) == 4
    finally:
        eng.shutdown()
