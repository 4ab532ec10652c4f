"""Multi-adapter LoRA batching (reference: test/registered/lora/*)."""
import sglang_b200 as sgl


def test_lora_adapters_route_per_request():
    eng = sgl.Engine(model_path="dummy:tiny-llama", page_size=4, enable_lora=True, lora_paths=["a=dummy:rank=4", "b=dummy:rank=8"])
    sp = {"max_new_tokens": 8, "temperature": 0, "ignore_eos": True}
    p = "the quick brown fox jumps"
    base = eng.generate(p, sp)["output_ids"]
    mixed = eng.generate([p, p, p], sp, lora_path=["a", None, "b"])
    assert mixed[1]["output_ids"] == base                      # no adapter -> base model
    assert mixed[0]["output_ids"] != base and mixed[2]["output_ids"] != base and mixed[0]["output_ids"] != mixed[2]["output_ids"]
    again = eng.generate(p, sp, lora_path="a")
    assert again["output_ids"] == mixed[0]["output_ids"]
    # KV computed under adapter "a" must never be served to the base model or adapter "b"
    assert eng.generate(p, sp)["output_ids"] == base
    eng.load_lora_adapter("c", "dummy:rank=2")
    assert eng.generate(p, sp, lora_path="c")["output_ids"] != base
    eng.unload_lora_adapter("c")
    eng.scheduler.check_invariants()


def test_peft_adapter_equals_merged_weights(tmp_path):
    """a PEFT-format adapter served through the LoRA path must equal the base model with ``W + (alpha / r) B A`` merged in; also an adapter
    trained on the fused qkv / gate_up projections"""
    import json

    import torch
    from safetensors.torch import save_file

    base = sgl.Engine(model_path="dummy:tiny-llama", page_size=4, device="cpu")
    cfg = base.model_config
    r, alpha = 4, 8.0
    g = torch.Generator().manual_seed(0)
    sd, fused = {}, {}
    H, hd, nh, nkv, I = cfg.hidden_size, cfg.head_dim, cfg.num_attention_heads, cfg.num_key_value_heads, cfg.intermediate_size
    dims = {"q_proj": (nh * hd, H), "k_proj": (nkv * hd, H), "v_proj": (nkv * hd, H), "o_proj": (H, nh * hd), "gate_proj": (I, H), "up_proj": (I, H),
            "down_proj": (H, I)}
    for li in range(cfg.num_hidden_layers):
        for sub, (o, i) in dims.items():
            blk = "self_attn" if sub in ("q_proj", "k_proj", "v_proj", "o_proj") else "mlp"
            A, B = torch.randn(r, i, generator=g) * 0.1, torch.randn(o, r, generator=g) * 0.1
            sd[f"base_model.model.model.layers.{li}.{blk}.{sub}.lora_A.weight"] = A
            sd[f"base_model.model.model.layers.{li}.{blk}.{sub}.lora_B.weight"] = B
    d = tmp_path / "peft"
    d.mkdir()
    save_file(sd, str(d / "adapter_model.safetensors"))
    (d / "adapter_config.json").write_text(json.dumps({"r": r, "lora_alpha": alpha, "target_modules": list(dims)}))
    # merged reference
    P = dict(base.runner.model.named_parameters())
    sc = alpha / r
    for li in range(cfg.num_hidden_layers):
        pre = f"model.layers.{li}."
        dl = lambda blk, sub: sc * sd[f"base_model.model.model.layers.{li}.{blk}.{sub}.lora_B.weight"] @ sd[f"base_model.model.model.layers.{li}.{blk}.{sub}.lora_A.weight"]
        P[pre + "self_attn.qkv_proj.weight"].data += torch.cat([dl("self_attn", "q_proj"), dl("self_attn", "k_proj"), dl("self_attn", "v_proj")], 0)
        P[pre + "self_attn.o_proj.weight"].data += dl("self_attn", "o_proj")
        P[pre + "mlp.gate_up_proj.weight"].data += torch.cat([dl("mlp", "gate_proj"), dl("mlp", "up_proj")], 0)
        P[pre + "mlp.down_proj.weight"].data += dl("mlp", "down_proj")
    prompts = [[5, 6, 7, 8, 9, 10], [11, 12, 13]]
    sp = {"max_new_tokens": 6, "temperature": 0, "ignore_eos": True}
    want = base.generate(input_ids=prompts, sampling_params=sp, return_logprob=True, logprob_start_len=0)
    base.shutdown()
    eng = sgl.Engine(model_path="dummy:tiny-llama", page_size=4, device="cpu", enable_lora=True, lora_paths=[f"x={d}"])
    got = eng.generate(input_ids=prompts, sampling_params=sp, lora_path=["x", "x"], return_logprob=True, logprob_start_len=0)
    plain = eng.generate(input_ids=prompts, sampling_params=sp)
    eng.shutdown()
    lp = lambda o: [v for v, _, _ in o["meta_info"]["input_token_logprobs"][1:]]
    for a, b, c in zip(want, got, plain):
        assert a["output_ids"] == b["output_ids"]
        assert max(abs(x - y) for x, y in zip(lp(a), lp(b))) < 2e-3
    assert any(a["output_ids"] != c["output_ids"] for a, c in zip(want, plain))         # the adapter actually changes the model
