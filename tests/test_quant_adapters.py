"""Foreign quantised checkpoint formats (AWQ, GPTQ, compressed-tensors, fp8, NVFP4, MXFP4, bitsandbytes) de-quantise exactly.
Each test packs a known integer/fp4 grid the way the format's own packer does and checks the adapter's output against the dense formula."""
import json

import torch

from sglang_b200.model_loader import quant_adapters as Q


def _pack_last(x, bits, order=None):
    n = 32 // bits
    x = x.to(torch.int64).reshape(x.shape[:-1] + (x.shape[-1] // n, n))
    if order is not None:
        x = x[..., order]
    out = torch.zeros(x.shape[:-1], dtype=torch.int64)
    for i in range(n):
        out |= x[..., i] << (bits * i)
    return torch.where(out >= 2 ** 31, out - 2 ** 32, out).to(torch.int32)


def _pack_first(x, bits):
    return _pack_last(x.t().contiguous(), bits).t().contiguous()


def test_awq():
    K, N, g = 256, 64, 128
    q = torch.randint(0, 16, (K, N))
    z = torch.randint(0, 16, (K // g, N))
    s = torch.rand(K // g, N) * 0.1 + 0.01
    ref = ((q - z.repeat_interleave(g, 0)).float() * s.repeat_interleave(g, 0)).t()
    got = Q.dequant_awq(_pack_last(q, 4, Q._AWQ_ORDER), _pack_last(z, 4, Q._AWQ_ORDER), s.half())
    torch.testing.assert_close(got, ((q - z.repeat_interleave(g, 0)).float() * s.half().float().repeat_interleave(g, 0)).t())
    assert (got - ref).abs().max() < 1e-2


def test_gptq_with_act_order_and_8bit():
    K, N, g = 256, 32, 64
    for bits in (4, 8):
        q = torch.randint(0, 2 ** bits, (K, N))
        z = torch.randint(1, 2 ** bits, (K // g, N))
        s = (torch.rand(K // g, N) * 0.1 + 0.01).half()
        g_idx = torch.randperm(K) // g
        ref = ((q - z[g_idx]).float() * s.float()[g_idx]).t()
        got = Q.dequant_gptq(_pack_first(q, bits), _pack_last(z - 1, bits), s, g_idx.int(), bits, zero_offset=1)
        torch.testing.assert_close(got, ref)
        got2 = Q.dequant_gptq(_pack_first(q, bits), _pack_last(z, bits), s, None, bits, zero_offset=0)
        torch.testing.assert_close(got2, ((q - z.repeat_interleave(g, 0)).float() * s.float().repeat_interleave(g, 0)).t())


def test_compressed_tensors_int4_and_channel_int8_and_fp8():
    N, K, g = 48, 256, 128
    q = torch.randint(-8, 8, (N, K))
    s = (torch.rand(N, K // g) * 0.1 + 0.01).half()
    got = Q.dequant_compressed_int(_pack_last(q + 8, 4), s, None, torch.tensor([N, K]), 4)
    torch.testing.assert_close(got, q.float() * s.float().repeat_interleave(g, 1))
    w8 = torch.randint(-128, 128, (N, K), dtype=torch.int8)
    sc = torch.rand(N, 1) * 0.01
    torch.testing.assert_close(Q.dequant_scaled(w8, sc), w8.float() * sc)
    wf = torch.randn(N, K).to(torch.float8_e4m3fn)
    torch.testing.assert_close(Q.dequant_scaled(wf, torch.tensor(0.5)), wf.float() * 0.5)
    blk = torch.rand(1, 2) + 0.5
    torch.testing.assert_close(Q.dequant_scaled(wf, blk), wf.float() * blk.repeat_interleave(128, 1)[:, :K])


def test_nvfp4_and_mxfp4():
    N, K = 8, 64
    idx = torch.randint(0, 16, (N, K))
    packed = (idx[:, 0::2] | (idx[:, 1::2] << 4)).to(torch.uint8)
    sc = (torch.rand(N, K // 16) + 0.5).to(torch.float8_e4m3fn)
    s2 = torch.tensor(0.25)
    torch.testing.assert_close(Q.dequant_nvfp4(packed, sc, s2), Q._FP4[idx] * (sc.float() * 0.25).repeat_interleave(16, 1))
    E = 3
    idx = torch.randint(0, 16, (E, N, K))
    blocks = (idx[..., 0::2] | (idx[..., 1::2] << 4)).to(torch.uint8).reshape(E, N, K // 32, 16)
    e8 = torch.randint(120, 134, (E, N, K // 32), dtype=torch.uint8)
    ref = Q._FP4[idx] * torch.exp2(e8.float() - 127).repeat_interleave(32, -1)
    torch.testing.assert_close(Q.dequant_mxfp4(blocks, e8), ref)


def test_bitsandbytes_nf4_double_quant():
    N, K, bs = 16, 128, 64
    qmap = torch.linspace(-1, 1, 16)
    idx = torch.randint(0, 16, (N * K,))
    packed = ((idx[0::2] << 4) | idx[1::2]).to(torch.uint8).reshape(-1, 1)
    nblocks = N * K // bs
    nested_map = torch.linspace(-1, 1, 256)
    am_q = torch.randint(0, 256, (nblocks,), dtype=torch.uint8)
    nested_am = torch.rand((nblocks + 255) // 256) + 0.5
    state = {"blocksize": bs, "shape": [N, K], "nested_blocksize": 256, "nested_offset": 0.3, "quant_type": "nf4"}
    am = nested_map[am_q.long()] * nested_am.repeat_interleave(256)[:nblocks] + 0.3
    ref = (qmap[idx] * am.repeat_interleave(bs)).reshape(N, K)
    torch.testing.assert_close(Q.dequant_bnb_4bit(packed, am_q, qmap, state, nested_am, nested_map), ref)


def test_adapter_stream_mixed_checkpoint():
    """a checkpoint with AWQ linears, a plain norm, an fp8 per-tensor linear, block-fp8 and mxfp4 experts, in shuffled order"""
    K, N, g = 128, 32, 128
    q, z, s = torch.randint(0, 16, (K, N)), torch.randint(0, 16, (1, N)), (torch.rand(1, N) * 0.1 + 0.01).half()
    awq = ((q - z.repeat_interleave(g, 0)).float() * s.float().repeat_interleave(g, 0)).t()
    wf = torch.randn(N, K).to(torch.float8_e4m3fn)
    blk = torch.rand(1, 1) + 0.5
    idx = torch.randint(0, 16, (2, N, 64))
    blocks = (idx[..., 0::2] | (idx[..., 1::2] << 4)).to(torch.uint8).reshape(2, N, 2, 16)
    e8 = torch.full((2, N, 2), 126, dtype=torch.uint8)
    stream = [("m.l0.q_proj.scales", s), ("m.norm.weight", torch.ones(K)), ("m.l0.q_proj.qweight", _pack_last(q, 4, Q._AWQ_ORDER)),
              ("m.l0.o_proj.weight", wf), ("m.l0.q_proj.qzeros", _pack_last(z, 4, Q._AWQ_ORDER)), ("m.l0.o_proj.weight_scale", torch.tensor(2.0)),
              ("m.l0.o_proj.input_scale", torch.tensor(1.0)), ("m.l0.up.weight_scale_inv", blk), ("m.l0.up.weight", wf),
              ("m.l0.experts.w13_blocks", blocks), ("m.l0.experts.w13_scales", e8), ("m.l0.q_proj.bias", torch.zeros(N)),
              ("m.embed.weight", torch.randn(4, K))]
    out = dict(Q.adapt_quantized_checkpoint(iter(stream), {"quant_method": "awq", "bits": 4, "group_size": 128}, torch.float32))
    torch.testing.assert_close(out["m.l0.q_proj.weight"], awq)
    torch.testing.assert_close(out["m.l0.o_proj.weight"], wf.float() * 2.0)
    torch.testing.assert_close(out["m.l0.up.weight"], wf.float() * blk)
    torch.testing.assert_close(out["m.l0.experts.w13"], (Q._FP4[idx] * 0.5).transpose(1, 2))
    assert out["m.norm.weight"].shape == (K,) and "m.l0.q_proj.bias" in out and out["m.embed.weight"].shape == (4, K)
    assert not any(k.endswith(("qweight", "qzeros", "scales", "weight_scale", "input_scale")) for k in out)
    kept = dict(Q.adapt_quantized_checkpoint(iter(stream), {"quant_method": "fp8"}, torch.float32, keep_block_fp8=True))
    assert kept["m.l0.up.weight"].dtype == torch.float8_e4m3fn and "m.l0.up.weight_scale_inv" in kept


def test_engine_loads_gptq_checkpoint(tmp_path):
    """end to end: write a GPTQ-format safetensors directory for tiny-llama (exact 8-bit grid) and serve it"""
    from safetensors.torch import save_file

    from sglang_b200 import Engine

    e = Engine(model_path="dummy:tiny-llama", device="cpu", page_size=4)
    cfg = e.model_config
    sd = {n: p.detach().float().clone() for n, p in e.runner.model.named_parameters()}
    prompts = [[5, 6, 7, 8, 9], [3, 4, 5, 6, 7, 8, 9, 10]]
    sp = {"max_new_tokens": 5, "temperature": 0, "ignore_eos": True}
    bits, g = 8, 32
    out = {}

    def add_linear(name, w):          # w [N, K] -> gptq tensors; snap w to the grid so the round trip is exact
        N, K = w.shape
        wg = w.t().reshape(K // g, g, N)
        mn, mx = wg.min(1).values, wg.max(1).values
        s = ((mx - mn) / (2 ** bits - 1)).clamp(min=1e-8).half()
        z = torch.round(-mn / s.float()).clamp(0, 2 ** bits - 1).to(torch.int64)
        q = torch.round(wg / s.float()[:, None] + z[:, None]).clamp(0, 2 ** bits - 1).to(torch.int64).reshape(K, N)
        out[name + ".qweight"], out[name + ".qzeros"] = _pack_first(q, bits), _pack_last(z - 1, bits)
        out[name + ".scales"], out[name + ".g_idx"] = s, (torch.arange(K) // g).int()
        return ((q - z.repeat_interleave(g, 0)).float() * s.float().repeat_interleave(g, 0)).t()

    H, hd, nh, nkv = cfg.hidden_size, cfg.head_dim, cfg.num_attention_heads, cfg.num_key_value_heads
    snapped = {}
    for i in range(cfg.num_hidden_layers):
        pre = f"model.layers.{i}."
        qw, kw_, vw = sd[pre + "self_attn.qkv_proj.weight"].split([nh * hd, nkv * hd, nkv * hd])
        gw, uw = sd[pre + "mlp.gate_up_proj.weight"].chunk(2)
        for nm, w in (("self_attn.q_proj", qw), ("self_attn.k_proj", kw_), ("self_attn.v_proj", vw), ("self_attn.o_proj", sd[pre + "self_attn.o_proj.weight"]),
                      ("mlp.gate_proj", gw), ("mlp.up_proj", uw), ("mlp.down_proj", sd[pre + "mlp.down_proj.weight"])):
            snapped[pre + nm] = add_linear(pre + nm, w)
        out[pre + "input_layernorm.weight"] = sd[pre + "input_layernorm.weight"]
        out[pre + "post_attention_layernorm.weight"] = sd[pre + "post_attention_layernorm.weight"]
    out["model.embed_tokens.weight"], out["model.norm.weight"] = sd["model.embed_tokens.weight"], sd["model.norm.weight"]
    if "lm_head.weight" in sd:
        out["lm_head.weight"] = sd["lm_head.weight"]
    d = tmp_path / "gptq"
    d.mkdir()
    save_file({k: v.contiguous() for k, v in out.items()}, str(d / "model.safetensors"))
    hf = {"architectures": ["LlamaForCausalLM"], "hidden_size": H, "intermediate_size": cfg.intermediate_size, "num_hidden_layers": cfg.num_hidden_layers,
          "num_attention_heads": nh, "num_key_value_heads": nkv, "vocab_size": cfg.vocab_size, "max_position_embeddings": cfg.max_position_embeddings,
          "rms_norm_eps": cfg.rms_norm_eps, "rope_theta": cfg.rope_theta, "torch_dtype": "float32", "tie_word_embeddings": "lm_head.weight" not in sd,
          "quantization_config": {"quant_method": "gptq", "bits": bits, "group_size": g, "desc_act": False}}
    (d / "config.json").write_text(json.dumps(hf))
    # reference = the same engine with the snapped (exactly representable) weights
    e.shutdown()
    got = Engine(model_path=str(d), device="cpu", page_size=4, skip_tokenizer_init=True)
    o1 = got.generate(input_ids=prompts, sampling_params=sp, return_logprob=True, logprob_start_len=0)
    w = dict(got.runner.model.named_parameters())
    q0 = w["model.layers.0.self_attn.qkv_proj.weight"][: nh * hd].float()
    got.shutdown()
    torch.testing.assert_close(q0, snapped["model.layers.0.self_attn.q_proj"], atol=1e-6, rtol=0)
    assert all(len(o["output_ids"]) == 5 for o in o1)
