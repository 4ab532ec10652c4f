"""Mistral3 (Pixtral tower + patch merger + Mistral decoder) against the Hugging Face implementation: aspect-ratio-preserving resize, per-row
image-token runs with break / end tokens, 2-D rotary tower, prompt logprobs and greedy continuation."""
import numpy as np
import pytest
import torch

import sglang_b200 as sgl
from sglang_b200.multimodal import processor as P

tf = pytest.importorskip("transformers")


def test_pixtral_resize_shape_matches_hf():
    from transformers.models.pixtral.image_processing_pixtral import get_resize_output_image_size

    for h, w in ((40, 50), (300, 200), (28, 2000), (1000, 1000), (57, 57), (3000, 1000)):
        img = np.zeros((h, w, 3), np.uint8)
        assert P.pixtral_resize_shape(h, w, 1540, 28) == tuple(get_resize_output_image_size(img, size=(1540, 1540), patch_size=(28, 28)))


@pytest.mark.parametrize("kind", ["mistral3", "pixtral-llava", "lightonocr"])
def test_engine_matches_hf(tmp_path, kind):
    if not hasattr(tf, "Mistral3ForConditionalGeneration"):
        pytest.skip("transformers has no Mistral3")
    torch.manual_seed(0)
    if kind == "pixtral-llava":          # pixtral-12b layout: LLaVA container, pixtral tower, one token per patch, biased two-layer projector
        cfg = tf.LlavaConfig(text_config=dict(model_type="mistral", hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4,
                                              num_key_value_heads=2, head_dim=16, vocab_size=512, max_position_embeddings=2048, rope_theta=10000.0,
                                              sliding_window=None, tie_word_embeddings=False),
                             vision_config=dict(model_type="pixtral", hidden_size=32, intermediate_size=48, num_hidden_layers=2, num_attention_heads=2,
                                                head_dim=16, patch_size=14, image_size=280),
                             image_token_index=500, vision_feature_layer=-1, vision_feature_select_strategy="full", tie_word_embeddings=False)
        hf = tf.LlavaForConditionalGeneration(cfg).float().eval()
    elif kind == "lightonocr":
        cfg = tf.LightOnOcrConfig(text_config=dict(hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2,
                                                   head_dim=16, vocab_size=512, max_position_embeddings=2048, tie_word_embeddings=False),
                                  vision_config=dict(hidden_size=32, intermediate_size=48, num_hidden_layers=2, num_attention_heads=2, head_dim=16, patch_size=14,
                                                     image_size=280), image_token_id=500, tie_word_embeddings=False)
        hf = tf.LightOnOcrForConditionalGeneration(cfg).float().eval()
    else:
        hf = None
        cfg = tf.Mistral3Config(text_config=dict(model_type="mistral", hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4,
                                             num_key_value_heads=2, head_dim=16, vocab_size=512, max_position_embeddings=2048, rope_theta=10000.0,
                                             sliding_window=None, tie_word_embeddings=False),
                            vision_config=dict(model_type="pixtral", hidden_size=32, intermediate_size=48, num_hidden_layers=2, num_attention_heads=2,
                                               head_dim=16, patch_size=14, image_size=280),
                            spatial_merge_size=2, image_token_id=500, tie_word_embeddings=False)
    hf = hf if hf is not None else tf.Mistral3ForConditionalGeneration(cfg).float().eval()
    for p in hf.parameters():
        if p.dim() > 1:
            torch.nn.init.normal_(p, std=0.08)
    d = str(tmp_path / kind)
    hf.save_pretrained(d, safe_serialization=True)
    rng = np.random.default_rng(0)
    imgs = [rng.integers(0, 255, size=(60, 90, 3), dtype=np.uint8), rng.integers(0, 255, size=(150, 56, 3), dtype=np.uint8)]
    ids = [5, 6, 500, 7, 8, 9, 500, 10, 11, 12, 13]
    eng = sgl.Engine(model_path=d, device="cpu", page_size=4, skip_tokenizer_init=True, chunked_prefill_size=16, disable_radix_cache=True, log_level="error")
    assert type(eng.runner.model).__name__ == {"mistral3": "Mistral3ForConditionalGeneration", "pixtral-llava": "PixtralForConditionalGeneration",
                                               "lightonocr": "LightOnOCRForConditionalGeneration"}[kind]
    o = eng.generate(input_ids=ids, image_data=imgs, sampling_params={"max_new_tokens": 6, "temperature": 0, "ignore_eos": True},
                     return_logprob=True, logprob_start_len=0)
    exp_ids, mm = P.process_images(ids, imgs, eng.model_config, None)
    eng.shutdown()
    hf_ids = list(exp_ids)
    for off, n in zip(mm["image_offsets"], mm["image_token_counts"]):
        hf_ids[off:off + n] = [500] * n
    grids, p = mm["grid_thw"], 14
    if kind in ("mistral3", "lightonocr"):
        assert grids == [[1, 6, 8], [1, 12, 4]] and len(mm["image_offsets"]) == 3 + 6                 # one splice run per merged row
        assert [hf_ids[o_ + n] for o_, n in zip(mm["image_offsets"], mm["image_token_counts"])] == [12, 12, 13] + [12] * 5 + [13]
    else:
        assert len(mm["image_offsets"]) == sum(g[1] for g in grids) and mm["image_token_counts"][0] == grids[0][2]      # one run per patch row
    pats, s, pix = np.asarray(mm["pixel_patches"], np.float32), 0, []
    Hm, Wm = max(g[1] for g in grids) * p, max(g[2] for g in grids) * p
    for _, gh, gw in grids:                     # HF takes a padded image batch + the true sizes
        x = pats[s:s + gh * gw].reshape(gh, gw, 3, p, p).transpose(2, 0, 3, 1, 4).reshape(3, gh * p, gw * p)
        s += gh * gw
        pad = np.zeros((3, Hm, Wm), np.float32)
        pad[:, : gh * p, : gw * p] = x
        pix.append(pad)
    full = hf_ids + o["output_ids"]
    with torch.no_grad():
        out = hf(input_ids=torch.tensor([full]), pixel_values=torch.from_numpy(np.stack(pix)),
                 image_sizes=torch.tensor([[g[1] * p, g[2] * p] for g in grids]))
    lp = torch.log_softmax(out.logits[0].float(), -1)
    ref = [float(lp[t, full[t + 1]]) for t in range(len(hf_ids) - 1)]
    got = [v for v, _, _ in o["meta_info"]["input_token_logprobs"][1:]]
    text_next = [t for t in range(len(hf_ids) - 1) if hf_ids[t + 1] != 500]
    assert max(abs(ref[t] - got[t]) for t in text_next) < 2e-3
    n = len(hf_ids)
    for j, tok in enumerate(o["output_ids"]):
        assert int(lp[n - 1 + j].argmax()) == tok and abs(float(lp[n - 1 + j, tok]) - o["meta_info"]["output_token_logprobs"][j][0]) < 2e-3
