"""GPT-OSS: the engine (paged attention + lse-based sinks, per-layer sliding windows, YaRN tables, clamped-SwiGLU MoE with biases) against a
dense, from-the-formulas PyTorch implementation on the same synthetic checkpoint."""
import math

import numpy as np
import torch

import sglang_b200 as sgl
from sglang_b200.ops.torch_ref import build_cos_sin_cache


def _reference_logprobs(cfg, W, ids):
    T, H, hd, nh, nkv = len(ids), cfg.hidden_size, cfg.head_dim, cfg.num_attention_heads, cfg.num_key_value_heads
    x = W["model.embed_tokens.weight"][ids].float()
    cs = build_cos_sin_cache(hd, T, cfg.rope_theta, cfg.rope_scaling)
    cos, sin = cs[:, : hd // 2], cs[:, hd // 2:]

    def rms(v, w):
        return v * torch.rsqrt(v.pow(2).mean(-1, keepdim=True) + cfg.rms_norm_eps) * w.float()

    def rope(t):          # [T, h, hd] neox halves
        a, b = t[..., : hd // 2], t[..., hd // 2:]
        return torch.cat([a * cos[:, None] - b * sin[:, None], b * cos[:, None] + a * sin[:, None]], -1)

    pos = torch.arange(T)
    for i in range(cfg.num_hidden_layers):
        p = f"model.layers.{i}."
        h = rms(x, W[p + "input_layernorm.weight"])
        q = (h @ W[p + "self_attn.q_proj.weight"].float().t() + W[p + "self_attn.q_proj.bias"].float()).view(T, nh, hd)
        k = (h @ W[p + "self_attn.k_proj.weight"].float().t() + W[p + "self_attn.k_proj.bias"].float()).view(T, nkv, hd)
        v = (h @ W[p + "self_attn.v_proj.weight"].float().t() + W[p + "self_attn.v_proj.bias"].float()).view(T, nkv, hd)
        q, k = rope(q), rope(k)
        k, v = k.repeat_interleave(nh // nkv, 1), v.repeat_interleave(nh // nkv, 1)
        s = torch.einsum("qhd,khd->hqk", q, k) / math.sqrt(hd)
        mask = pos[None, :] > pos[:, None]
        if cfg.layer_types[i] == "sliding_attention":
            mask |= pos[None, :] <= pos[:, None] - cfg.sliding_window
        s = s.masked_fill(mask[None], float("-inf"))
        sink = W[p + "self_attn.sinks"].float()[:, None, None].expand(nh, T, 1)
        pr = torch.softmax(torch.cat([s, sink], -1), -1)[..., :-1]          # the sink takes probability mass, contributes no value
        a = torch.einsum("hqk,khd->qhd", pr, v).reshape(T, nh * hd)
        x = x + a @ W[p + "self_attn.o_proj.weight"].float().t() + W[p + "self_attn.o_proj.bias"].float()
        h = rms(x, W[p + "post_attention_layernorm.weight"])
        lg = h @ W[p + "mlp.router.weight"].float().t() + W[p + "mlp.router.bias"].float()
        tv, ti = torch.topk(lg, cfg.num_experts_per_tok, -1)
        tw = torch.softmax(tv, -1)
        gu, gub = W[p + "mlp.experts.gate_up_proj"].float(), W[p + "mlp.experts.gate_up_proj_bias"].float()
        dn, dnb = W[p + "mlp.experts.down_proj"].float(), W[p + "mlp.experts.down_proj_bias"].float()
        y = torch.zeros_like(x)
        for t in range(T):
            for j in range(cfg.num_experts_per_tok):
                e = int(ti[t, j])
                z = h[t] @ gu[e] + gub[e]
                g, u = z[0::2].clamp(max=7.0), z[1::2].clamp(-7.0, 7.0)
                y[t] += tw[t, j] * (((u + 1) * g * torch.sigmoid(1.702 * g)) @ dn[e] + dnb[e])
        x = x + y
    x = rms(x, W["model.norm.weight"])
    logits = x @ W["lm_head.weight"].float().t()
    lp = torch.log_softmax(logits, -1)
    return [float(lp[t, ids[t + 1]]) for t in range(T - 1)]


def test_engine_matches_dense_reference():
    e = sgl.Engine(model_path="dummy:tiny-gpt-oss", device="cpu", page_size=4, chunked_prefill_size=16)
    cfg = e.model_config
    W = dict(e.runner.model.dummy_checkpoint(cfg, "cpu", torch.float32, e.server_args.random_seed))
    ids = np.random.default_rng(0).integers(3, 500, size=40).tolist()          # longer than the window (8) and the chunk (16)
    o = e.generate(input_ids=ids, sampling_params={"max_new_tokens": 4, "temperature": 0, "ignore_eos": True}, return_logprob=True,
                   logprob_start_len=0)
    got = [v for v, _, _ in o["meta_info"]["input_token_logprobs"][1:]]
    ref = _reference_logprobs(cfg, W, ids)
    assert max(abs(a - b) for a, b in zip(got, ref)) < 2e-3
    # decode continues consistently: teacher-forced prefill over prompt + outputs reproduces the decode logprobs
    full = e.generate(input_ids=ids + o["output_ids"], sampling_params={"max_new_tokens": 1, "temperature": 0}, return_logprob=True,
                      logprob_start_len=len(ids) - 1)
    tf = [v for v, _, _ in full["meta_info"]["input_token_logprobs"][1:]]
    dec = [v for v, _, _ in o["meta_info"]["output_token_logprobs"]]
    assert max(abs(a - b) for a, b in zip(tf, dec)) < 2e-3
    e.shutdown()
