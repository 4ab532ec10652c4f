"""Qwen2-Audio against Hugging Face: padded log-mel window + key-padding mask from the clip's true length, pooled encoder rows spliced over
the expanded audio placeholders, prompt logprobs and greedy continuation."""
import numpy as np
import pytest
import torch

import sglang_b200 as sgl
from sglang_b200.multimodal.audio import process_audio_llm

tf = pytest.importorskip("transformers")


def test_engine_matches_hf(tmp_path):
    if not hasattr(tf, "Qwen2AudioForConditionalGeneration"):
        pytest.skip("transformers has no Qwen2Audio")
    torch.manual_seed(0)
    cfg = tf.Qwen2AudioConfig(audio_config=dict(d_model=32, encoder_layers=2, encoder_attention_heads=2, encoder_ffn_dim=48, num_mel_bins=16,
                                                max_source_positions=60),
                              text_config=dict(model_type="qwen2", hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4,
                                               num_key_value_heads=2, vocab_size=512, max_position_embeddings=2048, tie_word_embeddings=False),
                              audio_token_index=500)
    hf = tf.Qwen2AudioForConditionalGeneration(cfg).float().eval()
    for p in hf.parameters():
        if p.dim() > 1:
            torch.nn.init.normal_(p, std=0.08)
    d = str(tmp_path / "qwen2audio")
    hf.save_pretrained(d, safe_serialization=True)
    rng = np.random.default_rng(0)
    clips = [(rng.standard_normal(16000 * 1 // 2).astype(np.float32) * 0.1, 16000), (rng.standard_normal(16000 * 3 // 10).astype(np.float32) * 0.1, 16000)]
    ids = [5, 6, 500, 7, 8, 9, 500, 10, 11]
    eng = sgl.Engine(model_path=d, device="cpu", page_size=4, skip_tokenizer_init=True, chunked_prefill_size=16, disable_radix_cache=True, log_level="error")
    assert type(eng.runner.model).__name__ == "Qwen2AudioForConditionalGeneration"
    o = eng.generate(input_ids=ids, audio_data=clips, sampling_params={"max_new_tokens": 6, "temperature": 0, "ignore_eos": True},
                     return_logprob=True, logprob_start_len=0)
    exp_ids, mm = process_audio_llm(ids, clips, eng.model_config)
    eng.shutdown()
    hf_ids = list(exp_ids)
    for off, n in zip(mm["image_offsets"], mm["image_token_counts"]):
        hf_ids[off:off + n] = [500] * n
    assert mm["audio_frames"] == [50, 30] and mm["image_token_counts"] == [12, 7] and len(hf_ids) == o["meta_info"]["prompt_tokens"]
    feats = torch.from_numpy(mm["audio_features"])
    fmask = torch.zeros(feats.shape[0], feats.shape[-1], dtype=torch.long)
    for i, f in enumerate(mm["audio_frames"]):
        fmask[i, :f] = 1
    full = hf_ids + o["output_ids"]
    with torch.no_grad():
        out = hf(input_ids=torch.tensor([full]), input_features=feats, feature_attention_mask=fmask)
    lp = torch.log_softmax(out.logits[0].float(), -1)
    ref = [float(lp[t, full[t + 1]]) for t in range(len(hf_ids) - 1)]
    got = [v for v, _, _ in o["meta_info"]["input_token_logprobs"][1:]]
    text_next = [t for t in range(len(hf_ids) - 1) if hf_ids[t + 1] != 500]
    assert max(abs(ref[t] - got[t]) for t in text_next) < 2e-3
    n = len(hf_ids)
    for j, tok in enumerate(o["output_ids"]):
        assert int(lp[n - 1 + j].argmax()) == tok and abs(float(lp[n - 1 + j, tok]) - o["meta_info"]["output_token_logprobs"][j][0]) < 2e-3


def test_voxtral_matches_hf(tmp_path):
    """Voxtral: fixed windows (a clip longer than one window becomes two), 4 encoder frames per soft token, bias-free projector"""
    if not hasattr(tf, "VoxtralForConditionalGeneration"):
        pytest.skip("transformers has no Voxtral")
    torch.manual_seed(0)
    cfg = tf.VoxtralConfig(audio_config=dict(hidden_size=32, num_hidden_layers=2, num_attention_heads=2, intermediate_size=128, num_mel_bins=16,
                                             max_source_positions=60, vocab_size=64),
                           text_config=dict(model_type="llama", hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4,
                                            num_key_value_heads=2, head_dim=16, vocab_size=512, max_position_embeddings=2048, tie_word_embeddings=False),
                           audio_token_id=500)
    hf = tf.VoxtralForConditionalGeneration(cfg).float().eval()
    for p in hf.parameters():
        if p.dim() > 1:
            torch.nn.init.normal_(p, std=0.08)
    d = str(tmp_path / "voxtral")
    hf.save_pretrained(d, safe_serialization=True)
    rng = np.random.default_rng(0)
    # window = 120 mel frames = 19200 samples: the first clip needs two windows, the second one
    clips = [(rng.standard_normal(25000).astype(np.float32) * 0.1, 16000), (rng.standard_normal(9000).astype(np.float32) * 0.1, 16000)]
    ids = [5, 6, 500, 7, 8, 9, 500, 10, 11]
    eng = sgl.Engine(model_path=d, device="cpu", page_size=4, skip_tokenizer_init=True, chunked_prefill_size=16, disable_radix_cache=True, log_level="error")
    assert type(eng.runner.model).__name__ == "VoxtralForConditionalGeneration"
    o = eng.generate(input_ids=ids, audio_data=clips, sampling_params={"max_new_tokens": 6, "temperature": 0, "ignore_eos": True},
                     return_logprob=True, logprob_start_len=0)
    exp_ids, mm = process_audio_llm(ids, clips, eng.model_config)
    eng.shutdown()
    assert mm["image_token_counts"] == [30, 15] and mm["audio_features"].shape == (3, 16, 120)
    hf_ids = list(exp_ids)
    for off, n in zip(mm["image_offsets"], mm["image_token_counts"]):
        hf_ids[off:off + n] = [500] * n
    full = hf_ids + o["output_ids"]
    with torch.no_grad():
        out = hf(input_ids=torch.tensor([full]), input_features=torch.from_numpy(mm["audio_features"]))
    lp = torch.log_softmax(out.logits[0].float(), -1)
    ref = [float(lp[t, full[t + 1]]) for t in range(len(hf_ids) - 1)]
    got = [v for v, _, _ in o["meta_info"]["input_token_logprobs"][1:]]
    text_next = [t for t in range(len(hf_ids) - 1) if hf_ids[t + 1] != 500]
    assert max(abs(ref[t] - got[t]) for t in text_next) < 2e-3
    n = len(hf_ids)
    for j, tok in enumerate(o["output_ids"]):
        assert int(lp[n - 1 + j].argmax()) == tok and abs(float(lp[n - 1 + j, tok]) - o["meta_info"]["output_token_logprobs"][j][0]) < 2e-3
