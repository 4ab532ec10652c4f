"""Qwen2-VL against the Hugging Face implementation (tiny random model written to disk): dynamic-resolution preprocessing, vision tower,
M-RoPE prefill positions, decode delta — prompt logprobs and greedy continuation must agree."""
import json

import numpy as np
import pytest
import torch

import sglang_b200 as sgl
from sglang_b200.multimodal import processor as P

tf = pytest.importorskip("transformers")


def _hf_model(tmp_path, v25=False):
    from transformers import Qwen2VLConfig, Qwen2VLForConditionalGeneration

    torch.manual_seed(0)
    if v25 in ("glm4v", "glm4vmoe"):
        from transformers import Glm4vConfig, Glm4vForConditionalGeneration, Glm4vMoeConfig, Glm4vMoeForConditionalGeneration

        moe = dict(n_routed_experts=4, num_experts_per_tok=2, moe_intermediate_size=32, n_shared_experts=1, first_k_dense_replace=1, n_group=1, topk_group=1,
                   norm_topk_prob=True, routed_scaling_factor=1.0) if v25 == "glm4vmoe" else {}
        cfg = (Glm4vMoeConfig if moe else Glm4vConfig)(text_config=dict(**moe, hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2,
                                           vocab_size=512, max_position_embeddings=2048,
                                           rope_parameters={"rope_type": "default", "rope_theta": 10000.0, "mrope_section": [1, 1, 2],
                                                            "partial_rotary_factor": 0.5}, tie_word_embeddings=False),
                          vision_config=dict(hidden_size=32, out_hidden_size=64, intermediate_size=48, depth=2, num_heads=2, patch_size=14,
                                             image_size=112, spatial_merge_size=2, temporal_patch_size=2),
                          image_token_id=500, video_token_id=501, image_start_token_id=502, image_end_token_id=503, tie_word_embeddings=False)
        m = (Glm4vMoeForConditionalGeneration if moe else Glm4vForConditionalGeneration)(cfg).float().eval()
        for p in m.parameters():
            if p.dim() > 1:
                torch.nn.init.normal_(p, std=0.08)
        d = tmp_path / v25
        m.save_pretrained(str(d), safe_serialization=True)
        return m, str(d)
    if v25 in ("v3", "v3moe"):
        from transformers import Qwen3VLConfig, Qwen3VLForConditionalGeneration, Qwen3VLMoeConfig, Qwen3VLMoeForConditionalGeneration

        moe = dict(num_experts=4, num_experts_per_tok=2, moe_intermediate_size=32, decoder_sparse_step=1, mlp_only_layers=[]) if v25 == "v3moe" else {}
        cfg = (Qwen3VLMoeConfig if moe else Qwen3VLConfig)(text_config=dict(**moe, hidden_size=64, intermediate_size=128, num_hidden_layers=3, num_attention_heads=4, num_key_value_heads=2,
                                             head_dim=16, vocab_size=512, max_position_embeddings=2048,
                                             rope_parameters={"rope_type": "default", "rope_theta": 10000.0, "mrope_section": [2, 3, 3],
                                                              "mrope_interleaved": True}, tie_word_embeddings=False),
                            vision_config=dict(hidden_size=32, out_hidden_size=64, intermediate_size=48, depth=3, num_heads=2, patch_size=16,
                                               spatial_merge_size=2, temporal_patch_size=2, num_position_embeddings=36, deepstack_visual_indexes=[0, 2]),
                            image_token_id=500, video_token_id=501, vision_start_token_id=502, tie_word_embeddings=False)
        m = (Qwen3VLMoeForConditionalGeneration if moe else Qwen3VLForConditionalGeneration)(cfg).float().eval()
        for p in m.parameters():
            if p.dim() > 1:
                torch.nn.init.normal_(p, std=0.08)
        d = tmp_path / ("qwen3vl" + ("moe" if moe else ""))
        m.save_pretrained(str(d), safe_serialization=True)
        return m, str(d)
    if v25:
        from transformers import Qwen2_5_VLConfig, Qwen2_5_VLForConditionalGeneration

        cfg = Qwen2_5_VLConfig(text_config=dict(hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2,
                                                vocab_size=512, max_position_embeddings=2048, rope_theta=10000.0,
                                                rope_scaling={"type": "mrope", "mrope_section": [2, 3, 3]}, tie_word_embeddings=False),
                               vision_config=dict(hidden_size=32, out_hidden_size=64, intermediate_size=48, depth=3, num_heads=2, patch_size=14,
                                                  spatial_merge_size=2, temporal_patch_size=2, window_size=56, fullatt_block_indexes=[1],
                                                  tokens_per_second=1),      # (this transformers build scales an IMAGE's temporal position by it)
                               image_token_id=500, video_token_id=501, vision_start_token_id=502)
        m = Qwen2_5_VLForConditionalGeneration(cfg).float().eval()
        for p in m.parameters():
            if p.dim() > 1:
                torch.nn.init.normal_(p, std=0.08)
        d = tmp_path / "qwen25vl"
        m.save_pretrained(str(d), safe_serialization=True)
        return m, str(d)
    cfg = Qwen2VLConfig(text_config=dict(hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2,
                                         vocab_size=512, max_position_embeddings=2048, rope_theta=10000.0,
                                         rope_scaling={"type": "mrope", "mrope_section": [2, 3, 3]}, tie_word_embeddings=False),
                        vision_config=dict(embed_dim=32, depth=2, num_heads=2, hidden_size=64, mlp_ratio=2, patch_size=14, spatial_merge_size=2,
                                           temporal_patch_size=2),
                        image_token_id=500, video_token_id=501, vision_start_token_id=502)
    m = Qwen2VLForConditionalGeneration(cfg).float().eval()
    for p in m.parameters():          # default init is tiny; make every part matter
        if p.dim() > 1:
            torch.nn.init.normal_(p, std=0.08)
    d = tmp_path / "qwen2vl"
    m.save_pretrained(str(d), safe_serialization=True)
    return m, str(d)


def test_smart_resize_and_patch_layout_match_hf():
    from transformers.models.qwen2_vl.image_processing_qwen2_vl import smart_resize

    for h, w in ((40, 50), (300, 200), (28, 2000), (1000, 1000), (57, 57)):
        assert P.smart_resize(h, w, 28, 56 * 56, 28 * 28 * 100) == smart_resize(h, w, 28, 56 * 56, 28 * 28 * 100)


@pytest.mark.parametrize("v25", [False, True, "v3", "v3moe", "glm4v", "glm4vmoe"])
def test_engine_matches_hf(tmp_path, v25):
    hf, path = _hf_model(tmp_path, v25)
    rng = np.random.default_rng(0)
    # 2.5: windows of 56 px = 2x2 merge blocks; the images below give partial windows on both axes
    imgs = [rng.integers(0, 255, size=(60, 90, 3), dtype=np.uint8), rng.integers(0, 255, size=(150, 56, 3), dtype=np.uint8)]
    ids = [5, 6, 502, 500, 7, 8, 9, 502, 500, 10, 11, 12, 13]
    eng = sgl.Engine(model_path=path, device="cpu", page_size=4, skip_tokenizer_init=True, chunked_prefill_size=16, disable_radix_cache=True)
    assert type(eng.runner.model).__name__ == {False: "Qwen2VLForConditionalGeneration", True: "Qwen2_5_VLForConditionalGeneration",
                                               "v3": "Qwen3VLForConditionalGeneration", "v3moe": "Qwen3VLMoeForConditionalGeneration", "glm4v": "Glm4vForConditionalGeneration", "glm4vmoe": "Glm4vMoeForConditionalGeneration"}[v25]
    o = eng.generate(input_ids=ids, image_data=imgs, sampling_params={"max_new_tokens": 6, "temperature": 0, "ignore_eos": True},
                     return_logprob=True, logprob_start_len=0)
    # the HF side: same patches / grids, placeholder runs expanded with the image token id
    cfg = eng.model_config
    exp_ids, mm = P.process_images_qwen_vl(ids, imgs, cfg, None)
    hf_ids = list(exp_ids)
    for off, n in zip(mm["image_offsets"], mm["image_token_counts"]):
        hf_ids[off:off + n] = [500] * n
    assert len(hf_ids) == o["meta_info"]["prompt_tokens"]
    pix = torch.from_numpy(np.asarray(mm["pixel_patches"], dtype=np.float32))
    grid = torch.tensor(mm["grid_thw"], dtype=torch.long)
    full = hf_ids + o["output_ids"]
    with torch.no_grad():
        out = hf(input_ids=torch.tensor([full]), pixel_values=pix, image_grid_thw=grid,
                 mm_token_type_ids=torch.tensor([[1 if t == 500 else 0 for t in full]]))
    lp = torch.log_softmax(out.logits[0].float(), -1)
    ref_prompt = [float(lp[t, full[t + 1]]) for t in range(len(hf_ids) - 1)]
    got_prompt = [v for v, _, _ in o["meta_info"]["input_token_logprobs"][1:]]
    # image-token positions predict "image token" in HF ids but hash pads in ours: compare where the TARGET is a text token
    text_next = [t for t in range(len(hf_ids) - 1) if hf_ids[t + 1] != 500]
    assert max(abs(ref_prompt[t] - got_prompt[t]) for t in text_next) < 2e-3
    # decode: the engine's outputs are HF's greedy choices and logprobs under teacher forcing (tests the M-RoPE decode delta)
    n = len(hf_ids)
    dec = [v for v, _, _ in o["meta_info"]["output_token_logprobs"]]
    for j, tok in enumerate(o["output_ids"]):
        assert int(lp[n - 1 + j].argmax()) == tok
        assert abs(float(lp[n - 1 + j, tok]) - dec[j]) < 2e-3
    eng.shutdown()
