"""One-shot GPU check of the late round-2 model families: tiny random Hugging Face checkpoints (bf16, head_dim 128 so the tcgen05 attention
kernels are the ones that run) served by the engine on cuda:0 and compared with the Hugging Face forward (fp32 math on the same bf16 weights).
Usage: python tools/gpu_families_check.py [device] [out.json]"""
import json
import sys
import tempfile
import time

import numpy as np
import torch
import transformers as tf

sys.path.insert(0, ".")
import sglang_b200 as sgl  # noqa: E402
from sglang_b200.multimodal import processor as P  # noqa: E402

DEV = sys.argv[1] if len(sys.argv) > 1 else "cuda"
OUT = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out/families_gpu.json"
TXT = dict(hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=2, num_key_value_heads=1, head_dim=128, vocab_size=512,
           max_position_embeddings=2048)
RNG = np.random.default_rng(0)
IMGS = [RNG.integers(0, 255, size=(60, 90, 3), dtype=np.uint8), RNG.integers(0, 255, size=(150, 56, 3), dtype=np.uint8)]
results = []


def save(model, name):
    torch.manual_seed(0)
    with torch.no_grad():
        for p in model.parameters():
            if p.dim() > 1:
                p.normal_(0, 0.05)
    model = model.to(torch.bfloat16)
    d = tempfile.mkdtemp(prefix=name)
    model.save_pretrained(d, safe_serialization=True)
    return model.float().eval(), d


def engine(d, **kw):
    return sgl.Engine(model_path=d, device=DEV, page_size=32, skip_tokenizer_init=True, log_level="error", max_total_tokens=4096,
                      disable_cuda_graph=True, max_running_requests=8, **kw)


def compare(name, hf, o, hf_ids, hf_kwargs, skip_tok=None):
    full = hf_ids + o["output_ids"]
    with torch.no_grad():
        lp = torch.log_softmax(hf(input_ids=torch.tensor([full]), **hf_kwargs).logits[0].float(), -1)
    ref = [float(lp[t, full[t + 1]]) for t in range(len(hf_ids) - 1)]
    got = [v for v, _, _ in o["meta_info"]["input_token_logprobs"][1:]]
    idx = [t for t in range(len(hf_ids) - 1) if skip_tok is None or hf_ids[t + 1] != skip_tok]
    dp = max(abs(ref[t] - got[t]) for t in idx)
    n = len(hf_ids)
    dd = max(abs(float(lp[n - 1 + j, tok]) - o["meta_info"]["output_token_logprobs"][j][0]) for j, tok in enumerate(o["output_ids"]))
    return {"family": name, "prefill_logprob_maxdiff": round(dp, 4), "decode_logprob_maxdiff": round(dd, 4), "ok": bool(dp < 0.15 and dd < 0.15)}


def run(name, fn):
    t0 = time.time()
    try:
        r = fn()
    except Exception as e:  # keep going: one family must not hide the others
        r = {"family": name, "ok": False, "error": f"{type(e).__name__}: {e}"[:300]}
    r["seconds"] = round(time.time() - t0, 1)
    results.append(r)
    print(json.dumps(r), flush=True)


SP = {"max_new_tokens": 6, "temperature": 0, "ignore_eos": True}
IDS = RNG.integers(3, 500, size=45).tolist()


def text_case(name, cfg_cls, model_cls, extra, **ekw):
    def f():
        hf, d = save(getattr(tf, model_cls)(getattr(tf, cfg_cls)(**{**TXT, **extra})), name)
        eng = engine(d, **ekw)
        try:
            arch = eng.model_config.architecture
            o = eng.generate(input_ids=IDS, sampling_params=SP, return_logprob=True, logprob_start_len=0)
        finally:
            eng.shutdown()
        r = compare(name, hf, o, IDS, {})
        r["served_by"] = arch
        return r
    run(name, f)


text_case("transformers-fallback(hunyuan-dense)", "HunYuanDenseV1Config", "HunYuanDenseV1ForCausalLM", {}, model_impl="transformers")
text_case("minimax-m2", "MiniMaxM2Config", "MiniMaxM2ForCausalLM", dict(num_local_experts=4, num_experts_per_tok=2))
text_case("ernie4.5-moe", "Ernie4_5_MoeConfig", "Ernie4_5_MoeForCausalLM", dict(moe_num_experts=4, moe_k=2, moe_intermediate_size=128, moe_num_shared_experts=1,
                                                                              moe_layer_start_index=1, moe_layer_interval=1))
text_case("qwen3.5", "Qwen3_5TextConfig", "Qwen3_5ForCausalLM", dict(num_hidden_layers=4, linear_num_key_heads=2, linear_num_value_heads=4, linear_key_head_dim=64,
                                                                      linear_value_head_dim=64, linear_conv_kernel_dim=4,
                                                                      layer_types=["linear_attention", "linear_attention", "linear_attention", "full_attention"]))


def vl_case(name, make_cfg, model_cls, ids, img_tok, hf_inputs):
    def f():
        hf, d = save(getattr(tf, model_cls)(make_cfg()), name)
        eng = engine(d, disable_radix_cache=True)
        try:
            o = eng.generate(input_ids=ids, image_data=IMGS, sampling_params=SP, return_logprob=True, logprob_start_len=0)
            exp_ids, mm = P.process_images(ids, IMGS, eng.model_config, None)
        finally:
            eng.shutdown()
        hf_ids = list(exp_ids)
        cnt = mm.get("image_token_counts") or [mm["num_image_tokens"]] * len(mm["image_offsets"])
        for off, n in zip(mm["image_offsets"], cnt):
            hf_ids[off:off + n] = [img_tok] * n
        return compare(name, hf, o, hf_ids, hf_inputs(mm, hf_ids + o["output_ids"]), skip_tok=img_tok)
    run(name, f)


vl_case("qwen3-vl",
        lambda: tf.Qwen3VLConfig(text_config=dict(**TXT, rope_parameters={"rope_type": "default", "rope_theta": 10000.0, "mrope_section": [24, 20, 20],
                                                                          "mrope_interleaved": True}, tie_word_embeddings=False),
                                 vision_config=dict(hidden_size=64, out_hidden_size=256, intermediate_size=96, depth=3, num_heads=2, patch_size=16,
                                                    spatial_merge_size=2, temporal_patch_size=2, num_position_embeddings=36, deepstack_visual_indexes=[0, 2]),
                                 image_token_id=500, video_token_id=501, vision_start_token_id=502, tie_word_embeddings=False),
        "Qwen3VLForConditionalGeneration", [5, 6, 502, 500, 7, 8, 9, 502, 500, 10, 11, 12, 13], 500,
        lambda mm, full: dict(pixel_values=torch.from_numpy(np.asarray(mm["pixel_patches"], np.float32)), image_grid_thw=torch.tensor(mm["grid_thw"]),
                              mm_token_type_ids=torch.tensor([[1 if t == 500 else 0 for t in full]])))


def m3_inputs(mm, full):
    grids, p = mm["grid_thw"], 14
    pats, s, pix = np.asarray(mm["pixel_patches"], np.float32), 0, []
    Hm, Wm = max(g[1] for g in grids) * p, max(g[2] for g in grids) * p
    for _, gh, gw in grids:
        x = pats[s:s + gh * gw].reshape(gh, gw, 3, p, p).transpose(2, 0, 3, 1, 4).reshape(3, gh * p, gw * p)
        s += gh * gw
        pad = np.zeros((3, Hm, Wm), np.float32)
        pad[:, : gh * p, : gw * p] = x
        pix.append(pad)
    return dict(pixel_values=torch.from_numpy(np.stack(pix)), image_sizes=torch.tensor([[g[1] * p, g[2] * p] for g in grids]))


vl_case("mistral3",
        lambda: tf.Mistral3Config(text_config=dict(model_type="mistral", **TXT, rope_theta=10000.0, sliding_window=None, tie_word_embeddings=False),
                                  vision_config=dict(model_type="pixtral", hidden_size=64, intermediate_size=96, num_hidden_layers=2, num_attention_heads=2,
                                                     head_dim=32, patch_size=14, image_size=280),
                                  spatial_merge_size=2, image_token_id=500, tie_word_embeddings=False),
        "Mistral3ForConditionalGeneration", [5, 6, 500, 7, 8, 9, 500, 10, 11, 12, 13], 500, m3_inputs)

vl_case("gemma3-mm",
        lambda: tf.Gemma3Config(text_config=dict(**{**TXT, "num_hidden_layers": 4}, sliding_window=8, query_pre_attn_scalar=128,
                                                 layer_types=["sliding_attention", "full_attention", "sliding_attention", "full_attention"]),
                                vision_config=dict(hidden_size=64, intermediate_size=96, num_hidden_layers=2, num_attention_heads=2, patch_size=14, image_size=112),
                                mm_tokens_per_image=16, image_token_id=500, boi_token_id=501, eoi_token_id=502),
        "Gemma3ForConditionalGeneration", [5, 6, 501, 500, 502, 7, 8, 9, 501, 500, 502, 10, 11, 12, 13], 500,
        lambda mm, full: dict(pixel_values=torch.from_numpy(np.asarray(mm["pixel_values"], np.float32)),
                              token_type_ids=torch.tensor([[1 if t == 500 else 0 for t in full]])))


def audio_case():
    from sglang_b200.multimodal.audio import process_audio_llm

    hf, d = save(tf.Qwen2AudioForConditionalGeneration(tf.Qwen2AudioConfig(
        audio_config=dict(d_model=64, encoder_layers=2, encoder_attention_heads=2, encoder_ffn_dim=96, num_mel_bins=16, max_source_positions=60),
        text_config=dict(model_type="qwen2", **{k: v for k, v in TXT.items() if k != "head_dim"}, tie_word_embeddings=False), audio_token_index=500)), "q2a")
    clips = [(RNG.standard_normal(8000).astype(np.float32) * 0.1, 16000), (RNG.standard_normal(4800).astype(np.float32) * 0.1, 16000)]
    ids = [5, 6, 500, 7, 8, 9, 500, 10, 11]
    eng = engine(d, disable_radix_cache=True)
    try:
        o = eng.generate(input_ids=ids, audio_data=clips, sampling_params=SP, return_logprob=True, logprob_start_len=0)
        exp_ids, mm = process_audio_llm(ids, clips, eng.model_config)
    finally:
        eng.shutdown()
    hf_ids = list(exp_ids)
    for off, n in zip(mm["image_offsets"], mm["image_token_counts"]):
        hf_ids[off:off + n] = [500] * n
    feats = torch.from_numpy(mm["audio_features"])
    fm = torch.zeros(feats.shape[0], feats.shape[-1], dtype=torch.long)
    for i, f in enumerate(mm["audio_frames"]):
        fm[i, :f] = 1
    return compare("qwen2-audio", hf, o, hf_ids, dict(input_features=feats, feature_attention_mask=fm), skip_tok=500)


run("qwen2-audio", audio_case)


def mixed_case():
    hf, d = save(tf.LlamaForCausalLM(tf.LlamaConfig(**TXT)), "llama")
    prompts = [RNG.integers(3, 500, size=n).tolist() for n in (5, 9, 300, 3, 500, 33, 7, 200)]
    sps = [{"max_new_tokens": m, "temperature": 0, "ignore_eos": True} for m in (20, 12, 6, 30, 5, 9, 25, 8)]
    outs, steps, mixed = {}, {}, 0
    for mc in (False, True):
        eng = engine(d, chunked_prefill_size=128, enable_mixed_chunk=mc)
        try:
            outs[mc] = [o["output_ids"] for o in eng.generate(input_ids=prompts, sampling_params=sps)]
            steps[mc] = eng.scheduler.forward_ct
            mixed = eng.scheduler.get_internal_state().get("num_mixed_decode_tokens", 0)
        finally:
            eng.shutdown()
    same = sum(a == b for a, b in zip(outs[False], outs[True]))
    return {"family": "mixed-chunk(llama)", "ok": bool(same >= 7 and mixed > 0), "identical_requests": same, "of": 8, "forward_passes": [steps[False], steps[True]],
            "mixed_decode_tokens": mixed}


run("mixed-chunk", mixed_case)
json.dump({"device": DEV, "torch": torch.__version__, "results": results}, open(OUT, "w"), indent=1)
print("ALL_OK" if all(r["ok"] for r in results) else "SOME_FAILED")
