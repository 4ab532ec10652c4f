"""Whisper: log-mel front end vs the HF feature extractor; encoder + cross-attention decoder vs HF on a tiny random checkpoint (prompt
logprobs and greedy continuation); two requests with different audio in one batch stay separate."""
import numpy as np
import pytest
import torch

import sglang_b200 as sgl
from sglang_b200.multimodal import audio as A

tf = pytest.importorskip("transformers")


def _wave(seed, seconds=1.3):
    rng = np.random.default_rng(seed)
    t = np.arange(int(seconds * 16000)) / 16000.0
    return (0.3 * np.sin(2 * np.pi * (220 + 40 * seed) * t) + 0.05 * rng.standard_normal(t.size)).astype(np.float32)


def test_log_mel_matches_hf_extractor():
    fe = tf.WhisperFeatureExtractor(feature_size=80)
    w = _wave(1, 2.5)
    ref = fe(w, sampling_rate=16000, return_tensors="np")["input_features"][0]
    got = A.log_mel_spectrogram(w)
    assert got.shape == ref.shape == (80, 3000)
    assert np.abs(got - ref).max() < 2e-3


def test_wav_bytes_round_trip(tmp_path):
    import wave

    w = _wave(2, 0.5)
    p = tmp_path / "a.wav"
    with wave.open(str(p), "wb") as f:
        f.setnchannels(2)
        f.setsampwidth(2)
        f.setframerate(32000)
        st = np.repeat((np.repeat(w, 2) * 32767).astype("<i2")[:, None], 2, 1)      # 32 kHz stereo
        f.writeframes(st.tobytes())
    x = A.load_audio(str(p))
    assert abs(x.shape[0] - w.shape[0]) <= 1 and np.abs(x[:w.shape[0] - 1] - w[: x.shape[0]][: w.shape[0] - 1]).mean() < 0.02


def test_engine_matches_hf(tmp_path):
    torch.manual_seed(0)
    cfg = tf.WhisperConfig(vocab_size=512, d_model=64, encoder_layers=2, decoder_layers=2, encoder_attention_heads=4, decoder_attention_heads=4,
                           encoder_ffn_dim=128, decoder_ffn_dim=128, num_mel_bins=80, max_source_positions=1500, max_target_positions=64,
                           pad_token_id=0, bos_token_id=1, eos_token_id=2, decoder_start_token_id=3, suppress_tokens=None, begin_suppress_tokens=None)
    hf = tf.WhisperForConditionalGeneration(cfg).float().eval()
    with torch.no_grad():
        for n, p in hf.named_parameters():
            if p.dim() > 1 and "embed_positions" not in n:
                p.normal_(0, 0.08)
    d = str(tmp_path / "whisper")
    hf.save_pretrained(d, safe_serialization=True)
    waves = [_wave(3), _wave(4, 0.7)]
    prompts = [[3, 10, 11, 12], [3, 20, 21]]
    eng = sgl.Engine(model_path=d, device="cpu", page_size=4, skip_tokenizer_init=True, log_level="error")
    assert eng.scheduler.tree.__class__.__name__ and eng.server_args.disable_radix_cache
    outs = eng.generate(input_ids=prompts, audio_data=waves, sampling_params={"max_new_tokens": 5, "temperature": 0, "ignore_eos": True},
                        return_logprob=True, logprob_start_len=0)
    eng.shutdown()
    for w, ids, o in zip(waves, prompts, outs):
        feats = torch.from_numpy(A.log_mel_spectrogram(w))[None]
        full = ids + o["output_ids"]
        with torch.no_grad():
            lp = torch.log_softmax(hf(input_features=feats, decoder_input_ids=torch.tensor([full])).logits[0].float(), -1)
        ref = [float(lp[t, full[t + 1]]) for t in range(len(ids) - 1)]
        got = [v for v, _, _ in o["meta_info"]["input_token_logprobs"][1:]]
        assert max(abs(a - b) for a, b in zip(ref, got)) < 3e-3
        dec = [v for v, _, _ in o["meta_info"]["output_token_logprobs"]]
        for j, tok in enumerate(o["output_ids"]):
            assert abs(float(lp[len(ids) - 1 + j, tok]) - dec[j]) < 3e-3


def test_transcriptions_endpoint_multipart_and_json():
    import base64
    import io
    import wave

    from fastapi.testclient import TestClient

    from sglang_b200.entrypoints.http_server import create_app

    eng = sgl.Engine(model_path="dummy:tiny-whisper", device="cpu", page_size=4, log_level="error")
    w = _wave(5, 0.6)
    buf = io.BytesIO()
    with wave.open(buf, "wb") as f:
        f.setnchannels(1)
        f.setsampwidth(2)
        f.setframerate(16000)
        f.writeframes((w * 32767).astype("<i2").tobytes())
    wav = buf.getvalue()
    direct = eng.generate(input_ids=[3, 10, 11], audio_data=wav, sampling_params={"max_new_tokens": 6, "temperature": 0})
    with TestClient(create_app(eng)) as c:
        r = c.post("/v1/audio/transcriptions", files={"file": ("a.wav", wav, "audio/wav")},
                   data={"model": "whisper", "prompt_ids": "[3, 10, 11]", "max_tokens": "6"})
        assert r.status_code == 200, r.text
        assert r.json()["output_ids"] == direct["output_ids"] and isinstance(r.json()["text"], str)
        r2 = c.post("/v1/audio/transcriptions", json={"audio": base64.b64encode(wav).decode(), "prompt_ids": [3, 10, 11], "max_tokens": 6,
                                                      "response_format": "text"})
        assert r2.status_code == 200 and r2.text == r.json()["text"]
        assert c.post("/v1/audio/transcriptions", json={"prompt_ids": [3]}).status_code == 400
    eng.shutdown()
