"""Encoder-only embedding / cross-encoder models and decoder classifiers against Hugging Face (tiny random checkpoints)."""
import numpy as np
import pytest
import torch

import sglang_b200 as sgl

tf = pytest.importorskip("transformers")


def _randomise(hf):
    with torch.no_grad():
        for n, p in hf.named_parameters():
            if p.dim() > 1:
                p.normal_(0, 0.08)
            else:
                p.add_(torch.randn_like(p) * 0.05)
    return hf.float().eval()


ENC = dict(vocab_size=512, hidden_size=64, num_hidden_layers=2, num_attention_heads=4, intermediate_size=128, max_position_embeddings=128, pad_token_id=1)


@pytest.mark.parametrize("name", ["BertModel", "XLMRobertaModel"])
def test_encoder_embeddings_match_hf(name, tmp_path):
    torch.manual_seed(0)
    cfg = (tf.BertConfig if name.startswith("Bert") else tf.XLMRobertaConfig)(**ENC)
    hf = _randomise(getattr(tf, name)(cfg, add_pooling_layer=False))
    d = str(tmp_path / name)
    hf.save_pretrained(d, safe_serialization=True)
    rng = np.random.default_rng(0)
    seqs = [rng.integers(5, 500, size=n).tolist() for n in (9, 23, 4)]
    e = sgl.Engine(model_path=d, device="cpu", page_size=4, skip_tokenizer_init=True, log_level="error", chunked_prefill_size=8)
    assert e.server_args.chunked_prefill_size is None and e.server_args.disable_radix_cache
    outs = e.encode(input_ids=seqs)
    e.shutdown()
    for ids, o in zip(seqs, outs):
        with torch.no_grad():
            ref = hf(input_ids=torch.tensor([ids])).last_hidden_state[0, 0]
        ref = torch.nn.functional.normalize(ref, dim=-1)
        assert torch.allclose(torch.tensor(o["embedding"]), ref, atol=2e-4)


@pytest.mark.parametrize("name", ["BertForSequenceClassification", "XLMRobertaForSequenceClassification"])
def test_cross_encoder_scores_match_hf(name, tmp_path):
    torch.manual_seed(0)
    cfg = (tf.BertConfig if name.startswith("Bert") else tf.XLMRobertaConfig)(**ENC, num_labels=1, sep_token_id=3)
    hf = _randomise(getattr(tf, name)(cfg))
    d = str(tmp_path / name)
    hf.save_pretrained(d, safe_serialization=True)
    seqs = [[2, 10, 11, 12, 3, 40, 41, 3], [2, 20, 3, 50, 51, 52, 53, 3]]          # <cls> query <sep> document <sep>
    e = sgl.Engine(model_path=d, device="cpu", page_size=4, skip_tokenizer_init=True, log_level="error")
    outs = e.encode(input_ids=seqs)
    e.shutdown()
    for ids, o in zip(seqs, outs):
        tt = None
        if name.startswith("Bert"):
            first = ids.index(3)
            tt = torch.tensor([[0] * (first + 1) + [1] * (len(ids) - first - 1)])
        with torch.no_grad():
            ref = hf(input_ids=torch.tensor([ids]), token_type_ids=tt).logits[0]
        assert torch.allclose(torch.tensor(o["embedding"]), ref, atol=2e-4)


@pytest.mark.parametrize("name,cfgname", [("LlamaForSequenceClassification", "LlamaConfig"), ("Qwen2ForSequenceClassification", "Qwen2Config")])
def test_decoder_classifier_matches_hf(name, cfgname, tmp_path):
    torch.manual_seed(0)
    cfg = getattr(tf, cfgname)(hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2, vocab_size=512,
                               max_position_embeddings=512, num_labels=3, pad_token_id=0)
    hf = _randomise(getattr(tf, name)(cfg))
    d = str(tmp_path / name)
    hf.save_pretrained(d, safe_serialization=True)
    seqs = [[5, 6, 7, 8, 9, 10], [11, 12, 13]]
    e = sgl.Engine(model_path=d, device="cpu", page_size=4, skip_tokenizer_init=True, log_level="error")
    outs = e.encode(input_ids=seqs)
    e.shutdown()
    for ids, o in zip(seqs, outs):
        with torch.no_grad():
            ref = hf(input_ids=torch.tensor([ids])).logits[0]
        assert torch.allclose(torch.tensor(o["embedding"]), ref, atol=2e-3)


def test_rerank_uses_the_cross_encoder_head(tmp_path):
    """/v1/rerank on a ...ForSequenceClassification model scores (query, document) PAIRS in one pass each"""
    torch.manual_seed(1)
    cfg = tf.XLMRobertaConfig(**ENC, num_labels=1)
    hf = _randomise(tf.XLMRobertaForSequenceClassification(cfg))
    d = str(tmp_path / "rr")
    hf.save_pretrained(d, safe_serialization=True)

    class PairTok:          # minimal pair tokenizer: bytes + 5, <s>=0 </s>=2
        sep_token_id = 2

        def encode(self, t):
            return [0] + [5 + (b % 200) for b in t.encode()] + [2]

        def __call__(self, a, b=None, **kw):
            ids = self.encode(a) + ([2] + self.encode(b)[1:] if b is not None else [])
            return {"input_ids": ids}

        def decode(self, ids, **kw):
            return ""

    e = sgl.Engine(model_path=d, device="cpu", page_size=4, skip_tokenizer_init=True, log_level="error")
    e.tokenizer = PairTok()
    docs = ["alpha beta", "gamma", "delta epsilon zeta"]
    got = e.rerank("what is alpha", docs)
    e.shutdown()
    tk = PairTok()
    ref = []
    for dd in docs:
        with torch.no_grad():
            ref.append(float(hf(input_ids=torch.tensor([tk("what is alpha", dd)["input_ids"]])).logits[0, 0]))
    assert [r["index"] for r in got] == sorted(range(3), key=lambda i: -ref[i])
    assert all(abs(r["score"] - ref[r["index"]]) < 2e-4 for r in got)
