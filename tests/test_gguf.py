"""GGUF container + block de-quantisers + engine load (reference: test/registered/quant/test_gguf.py loads a real GGUF from the hub;
offline we write the file ourselves and check the K-quant de-quantisers against a direct scalar transcription of the block specs)."""
import os
import struct

import numpy as np
import pytest
import torch

from sglang_b200 import Engine
from sglang_b200.model_loader import gguf as G


def _h(b, o):
    return float(np.frombuffer(bytes(b[o:o + 2]), dtype=np.float16)[0])


def _k4(j, q):
    if j < 4:
        return q[j] & 63, q[j + 4] & 63
    return (q[j + 4] & 0xF) | ((q[j - 4] >> 6) << 4), (q[j + 4] >> 4) | ((q[j] >> 6) << 4)


def _ref_block(name, b):
    """scalar, loop-for-loop reading of one 256-element super-block"""
    b = [int(v) for v in b]
    y = []
    if name == "Q4_K":
        d, dmin, sc, q = _h(b, 0), _h(b, 2), b[4:16], b[16:144]
        for j in range(4):
            s1, m1 = _k4(2 * j, sc)
            s2, m2 = _k4(2 * j + 1, sc)
            y += [d * s1 * (q[32 * j + l] & 0xF) - dmin * m1 for l in range(32)]
            y += [d * s2 * (q[32 * j + l] >> 4) - dmin * m2 for l in range(32)]
    elif name == "Q5_K":
        d, dmin, sc, qh, q = _h(b, 0), _h(b, 2), b[4:16], b[16:48], b[48:176]
        for j in range(4):
            s1, m1 = _k4(2 * j, sc)
            s2, m2 = _k4(2 * j + 1, sc)
            u1, u2 = 1 << (2 * j), 2 << (2 * j)
            y += [d * s1 * ((q[32 * j + l] & 0xF) + (16 if qh[l] & u1 else 0)) - dmin * m1 for l in range(32)]
            y += [d * s2 * ((q[32 * j + l] >> 4) + (16 if qh[l] & u2 else 0)) - dmin * m2 for l in range(32)]
    elif name == "Q6_K":
        ql, qh, sc, d = b[0:128], b[128:192], [v - 256 if v > 127 else v for v in b[192:208]], _h(b, 208)
        for n in range(2):
            out = [0.0] * 128
            for l in range(32):
                i = l // 16
                h = qh[32 * n + l]
                q1 = ((ql[64 * n + l] & 0xF) | (((h >> 0) & 3) << 4)) - 32
                q2 = ((ql[64 * n + l + 32] & 0xF) | (((h >> 2) & 3) << 4)) - 32
                q3 = ((ql[64 * n + l] >> 4) | (((h >> 4) & 3) << 4)) - 32
                q4 = ((ql[64 * n + l + 32] >> 4) | (((h >> 6) & 3) << 4)) - 32
                out[l], out[l + 32] = d * sc[8 * n + i] * q1, d * sc[8 * n + i + 2] * q2
                out[l + 64], out[l + 96] = d * sc[8 * n + i + 4] * q3, d * sc[8 * n + i + 6] * q4
            y += out
    elif name == "Q2_K":
        scales, q, d, dmin = b[0:16], b[16:80], _h(b, 80), _h(b, 82)
        i = 0
        for n in range(2):
            for j in range(4):
                for half in range(2):
                    s = scales[i]
                    i += 1
                    y += [d * (s & 0xF) * ((q[32 * n + 16 * half + l] >> (2 * j)) & 3) - dmin * (s >> 4) for l in range(16)]
    elif name == "Q3_K":
        hm, q, s12, d = b[0:32], b[32:96], b[96:108], _h(b, 108)
        aux = list(struct.unpack("<3I", bytes(s12)))
        k1, k2 = 0x03030303, 0x0f0f0f0f
        tmp = aux[2]
        a = [(aux[0] & k2) | (((tmp >> 0) & k1) << 4), (aux[1] & k2) | (((tmp >> 2) & k1) << 4),
             ((aux[0] >> 4) & k2) | (((tmp >> 4) & k1) << 4), ((aux[1] >> 4) & k2) | (((tmp >> 6) & k1) << 4)]
        sc = [v - 256 if v > 127 else v for v in struct.pack("<4I", *a)]
        i, m = 0, 1
        for n in range(2):
            for j in range(4):
                for half in range(2):
                    dl = d * (sc[i] - 32)
                    i += 1
                    y += [dl * (((q[32 * n + 16 * half + l] >> (2 * j)) & 3) - (0 if hm[16 * half + l] & m else 4)) for l in range(16)]
                m <<= 1
    return np.asarray(y, np.float32)


@pytest.mark.parametrize("name", ["Q4_K", "Q5_K", "Q6_K", "Q2_K", "Q3_K"])
def test_k_quant_dequant_matches_scalar_spec(name):
    typ = G._NAME_TO_TYPE[name]
    _, be, bb = G.GGML_TYPES[typ]
    rng = np.random.default_rng(0)
    nb = 6
    raw = rng.integers(0, 256, size=(nb, bb), dtype=np.uint8)
    # keep the fp16 scale fields finite and small
    for off in {"Q4_K": (0, 2), "Q5_K": (0, 2), "Q6_K": (208,), "Q2_K": (80, 82), "Q3_K": (108,)}[name]:
        raw[:, off:off + 2] = rng.uniform(0.01, 0.2, size=(nb, 1)).astype(np.float16).view(np.uint8)
    got = G.dequantize(raw.reshape(-1), typ, (nb, be)).numpy()
    ref = np.stack([_ref_block(name, raw[i]) for i in range(nb)])
    np.testing.assert_allclose(got, ref, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("name,tol", [("F32", 0), ("F16", 1e-3), ("BF16", 1e-2), ("Q8_0", 0.02), ("Q4_0", 0.25), ("Q4_1", 0.2)])
def test_legacy_quant_round_trip(name, tol, tmp_path):
    x = torch.randn(8, 64)
    p = str(tmp_path / "t.gguf")
    G.write_gguf(p, {"general.architecture": "llama", "k": [1, 2, 3], "s": "hi", "f": 0.5, "names": ["a", "b"]}, {"w": (x, name)})
    g = G.GGUFFile(p)
    assert g.metadata["k"] == [1, 2, 3] and g.metadata["s"] == "hi" and g.metadata["names"] == ["a", "b"] and g.tensors["w"][0] == (8, 64)
    y = g.get("w").float()
    assert (y - x).abs().max().item() <= tol * max(1.0, x.abs().max().item()) + 1e-7


def test_q5_dequant_bit_placement():
    # Q5_0 block: d = 1, element e gets high bit (qh >> e) & 1
    qh = 0b1010_0000_0000_0001_0000_0000_0000_0110
    qs = bytes((i & 0xF) | (((15 - i) & 0xF) << 4) for i in range(16))
    blk = np.frombuffer(np.float16(1.0).tobytes() + struct.pack("<I", qh) + qs, dtype=np.uint8)
    y = G.dequantize(blk, G._NAME_TO_TYPE["Q5_0"], (32,)).numpy()
    ref = [((i & 0xF) | (((qh >> i) & 1) << 4)) - 16 for i in range(16)] + [(((15 - i) & 0xF) | (((qh >> (i + 16)) & 1) << 4)) - 16 for i in range(16)]
    assert y.tolist() == [float(v) for v in ref]


def _export_llama(engine, path, qtype):
    """what llama.cpp's converter does for a Llama checkpoint: rename, permute q/k, quantise 2-D weights"""
    cfg = engine.model_config
    sd = {}
    for n, p in engine.runner.model.named_parameters():
        sd[n] = p.detach().float()
    H, hd, nh, nkv = cfg.hidden_size, cfg.head_dim, cfg.num_attention_heads, cfg.num_key_value_heads
    out = {"token_embd.weight": (sd["model.embed_tokens.weight"], "F16"), "output_norm.weight": (sd["model.norm.weight"], "F32")}
    if "lm_head.weight" in sd:
        out["output.weight"] = (sd["lm_head.weight"], qtype)
    for i in range(cfg.num_hidden_layers):
        pre = f"model.layers.{i}."
        q, k, v = sd[pre + "self_attn.qkv_proj.weight"].split([nh * hd, nkv * hd, nkv * hd])
        g, u = sd[pre + "mlp.gate_up_proj.weight"].chunk(2)
        out.update({f"blk.{i}.attn_q.weight": (G.permute_qk(q, nh), qtype), f"blk.{i}.attn_k.weight": (G.permute_qk(k, nkv), qtype),
                    f"blk.{i}.attn_v.weight": (v, qtype), f"blk.{i}.attn_output.weight": (sd[pre + "self_attn.o_proj.weight"], qtype),
                    f"blk.{i}.ffn_gate.weight": (g, qtype), f"blk.{i}.ffn_up.weight": (u, qtype),
                    f"blk.{i}.ffn_down.weight": (sd[pre + "mlp.down_proj.weight"], qtype),
                    f"blk.{i}.attn_norm.weight": (sd[pre + "input_layernorm.weight"], "F32"),
                    f"blk.{i}.ffn_norm.weight": (sd[pre + "post_attention_layernorm.weight"], "F32")})
    md = {"general.architecture": "llama", "llama.block_count": cfg.num_hidden_layers, "llama.embedding_length": H,
          "llama.feed_forward_length": cfg.intermediate_size, "llama.attention.head_count": nh, "llama.attention.head_count_kv": nkv,
          "llama.attention.layer_norm_rms_epsilon": float(cfg.rms_norm_eps), "llama.rope.freq_base": float(cfg.rope_theta),
          "llama.context_length": cfg.max_position_embeddings, "llama.vocab_size": cfg.vocab_size, "llama.attention.key_length": hd}
    G.write_gguf(path, md, out)


def test_engine_loads_gguf(tmp_path):
    prompts = [[5, 6, 7, 8, 9], [3, 4, 5, 6, 7, 8, 9, 10]]
    sp = {"max_new_tokens": 6, "temperature": 0, "ignore_eos": True}
    kw = dict(device="cpu", page_size=4)
    e = Engine(model_path="dummy:tiny-llama", **kw)
    ref = e.generate(input_ids=prompts, sampling_params=sp, return_logprob=True, logprob_start_len=0)
    f16, q8 = str(tmp_path / "m-f16.gguf"), str(tmp_path / "m-q8.gguf")
    _export_llama(e, f16, "BF16")
    _export_llama(e, q8, "Q8_0")
    e.shutdown()

    def lps(outs):
        return [[x[0] for x in o["meta_info"]["input_token_logprobs"][1:]] for o in outs]

    for path, tol in ((f16, 2e-2), (q8, 8e-2)):
        e2 = Engine(model_path=path, skip_tokenizer_init=True, **kw)
        assert e2.model_config.architecture == "LlamaForCausalLM" and e2.model_config.num_hidden_layers == e.model_config.num_hidden_layers
        got = e2.generate(input_ids=prompts, sampling_params=sp, return_logprob=True, logprob_start_len=0)
        e2.shutdown()
        err = max(abs(a - b) for x, y in zip(lps(ref), lps(got)) for a, b in zip(x, y))
        assert err < tol, (path, err)
        if path == f16:
            assert [o["output_ids"] for o in got] == [o["output_ids"] for o in ref]


def test_gguf_tokenizer_bpe():
    # byte-level BPE vocabulary: single bytes + two merges
    from tokenizers import pre_tokenizers

    alphabet = sorted(pre_tokenizers.ByteLevel.alphabet())
    toks = alphabet + ["he", "ll", "hell"]
    md = {"tokenizer.ggml.model": "gpt2", "tokenizer.ggml.tokens": toks, "tokenizer.ggml.merges": ["h e", "l l", "he ll"],
          "tokenizer.ggml.eos_token_id": 0}
    tk = G.GGUFTokenizer(md)
    ids = tk.encode("hello hell")
    assert tk.decode(ids) == "hello hell"
    assert toks.index("hell") in ids and len(tk) == len(toks)
