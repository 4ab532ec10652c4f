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


@pytest.mark.parametrize("arch", ["Qwen2ForRewardModel", "InternLM2ForRewardModel", "LlamaForSequenceClassificationWithNormal_Weights"])
def test_reward_heads(arch, tmp_path):
    """reward-model heads that Hugging Face ships as remote code: build the checkpoint by hand (decoder + head tensors) and check the score
    against the decoder's last hidden state pushed through the same head"""
    import json

    from safetensors.torch import load_file, save_file

    torch.manual_seed(0)
    base = {"Qwen2ForRewardModel": "Qwen2", "InternLM2ForRewardModel": "Llama", "LlamaForSequenceClassificationWithNormal_Weights": "Llama"}[arch]
    cfg = getattr(tf, base + "Config")(hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2,
                                       vocab_size=512, max_position_embeddings=512, pad_token_id=0)
    hf = _randomise(getattr(tf, base + "ForCausalLM")(cfg))
    d = str(tmp_path / arch)
    hf.save_pretrained(d, safe_serialization=True)
    sd = load_file(d + "/model.safetensors")
    H = 64
    if arch == "Qwen2ForRewardModel":
        head = {"score.0.weight": torch.randn(H, H) * 0.1, "score.0.bias": torch.randn(H) * 0.1, "score.2.weight": torch.randn(1, H) * 0.1,
                "score.2.bias": torch.randn(1) * 0.1}
        score = lambda h: torch.relu(h @ head["score.0.weight"].T + head["score.0.bias"]) @ head["score.2.weight"].T + head["score.2.bias"]
    elif arch == "InternLM2ForRewardModel":
        head = {"v_head.weight": torch.randn(1, H) * 0.1}
        score = lambda h: h @ head["v_head.weight"].T
        ren = {}
        G = 2
        for k, v in list(sd.items()):          # InternLM2 checkpoint names: fused wqkv per KV group, w1/w3/w2, attention_norm / ffn_norm
            if ".self_attn.q_proj." in k:
                pre = k.split(".self_attn.")[0]
                q, kk, vv = v, sd[pre + ".self_attn.k_proj.weight"], sd[pre + ".self_attn.v_proj.weight"]
                w = torch.cat([q.view(2, G, 16, H), kk.view(2, 1, 16, H), vv.view(2, 1, 16, H)], 1).reshape(-1, H)
                ren[pre.replace("model.layers", "model.layers") + ".attention.wqkv.weight"] = w
        for k, v in sd.items():
            if ".self_attn.q_proj." in k or ".self_attn.k_proj." in k or ".self_attn.v_proj." in k:
                continue
            k2 = k.replace(".self_attn.o_proj.", ".attention.wo.").replace(".mlp.gate_proj.", ".feed_forward.w1.").replace(".mlp.up_proj.", ".feed_forward.w3.") \
                .replace(".mlp.down_proj.", ".feed_forward.w2.").replace(".input_layernorm.", ".attention_norm.").replace(".post_attention_layernorm.", ".ffn_norm.") \
                .replace("model.embed_tokens.", "model.tok_embeddings.").replace("lm_head.weight", "output.weight")
            ren[k2] = v
        sd = ren
    else:
        head = {"score.weight": torch.randn(3, H) * 0.1, "weights.weight": torch.randn(3, H) * 0.1}
        score = lambda h: ((h @ head["score.weight"].T) * (h @ head["weights.weight"].T)).sum(-1, keepdim=True)
    save_file({**{k: v.contiguous() for k, v in sd.items()}, **head}, d + "/model.safetensors")
    c = json.load(open(d + "/config.json"))
    c["architectures"] = [arch]
    if arch.endswith("Normal_Weights"):
        c["num_labels"] = 3
        c["id2label"] = {str(i): str(i) for i in range(3)}
    json.dump(c, open(d + "/config.json", "w"))
    seqs = [[5, 6, 7, 8, 9, 10], [11, 12, 13]]
    e = sgl.Engine(model_path=d, device="cpu", page_size=4, skip_tokenizer_init=True, log_level="error")
    outs = e.encode(input_ids=seqs)
    e.shutdown()
    for ids, o in zip(seqs, outs):
        with torch.no_grad():
            h = hf.model(input_ids=torch.tensor([ids])).last_hidden_state[0, -1]
        assert torch.allclose(torch.tensor(o["embedding"]), score(h), atol=2e-3), arch


def test_contriever_mean_pooling_matches_hf(tmp_path):
    """facebook/contriever layout: a BERT checkpoint whose architecture is ``Contriever``; the embedding is the un-normalised token mean"""
    import json

    torch.manual_seed(0)
    hf = _randomise(tf.BertModel(tf.BertConfig(**ENC), add_pooling_layer=False))
    d = str(tmp_path / "contriever")
    hf.save_pretrained(d, safe_serialization=True)
    c = json.load(open(d + "/config.json"))
    c["architectures"] = ["Contriever"]
    json.dump(c, open(d + "/config.json", "w"))
    rng = np.random.default_rng(0)
    seqs = [rng.integers(5, 500, size=n).tolist() for n in (9, 23, 4)]
    e = sgl.Engine(model_path=d, device="cpu", page_size=4, skip_tokenizer_init=True, log_level="error")
    assert type(e.runner.model).__name__ == "Contriever"
    outs = e.encode(input_ids=seqs)
    e.shutdown()
    for ids, o in zip(seqs, outs):
        with torch.no_grad():
            ref = hf(input_ids=torch.tensor([ids])).last_hidden_state[0].mean(0)
        assert torch.allclose(torch.tensor(o["embedding"]), ref, atol=2e-4)
