"""Gemma-3 multimodal against Hugging Face: SigLIP tower, pooled + normed + projected soft tokens, bidirectional attention inside an image
block (also on the sliding-window layers, whose window is shorter than one image here)."""
import numpy as np
import pytest
import torch

import sglang_b200 as sgl
from sglang_b200.multimodal import processor as P

tf = pytest.importorskip("transformers")


def test_engine_matches_hf(tmp_path):
    torch.manual_seed(0)
    cfg = tf.Gemma3Config(text_config=dict(hidden_size=64, intermediate_size=128, num_hidden_layers=4, num_attention_heads=4, num_key_value_heads=2,
                                           head_dim=16, vocab_size=512, max_position_embeddings=2048, sliding_window=8, query_pre_attn_scalar=16,
                                           layer_types=["sliding_attention", "full_attention", "sliding_attention", "full_attention"]),
                          vision_config=dict(hidden_size=32, intermediate_size=48, num_hidden_layers=2, num_attention_heads=2, patch_size=14, image_size=112),
                          mm_tokens_per_image=16, image_token_id=500, boi_token_id=501, eoi_token_id=502)
    hf = tf.Gemma3ForConditionalGeneration(cfg).float().eval()
    for p in hf.parameters():
        if p.dim() > 1:
            torch.nn.init.normal_(p, std=0.08)
    d = str(tmp_path / "gemma3mm")
    hf.save_pretrained(d, safe_serialization=True)
    rng = np.random.default_rng(0)
    imgs = [rng.integers(0, 255, size=(60, 90, 3), dtype=np.uint8), rng.integers(0, 255, size=(150, 56, 3), dtype=np.uint8)]
    ids = [5, 6, 501, 500, 502, 7, 8, 9, 501, 500, 502, 10, 11, 12, 13]
    eng = sgl.Engine(model_path=d, device="cpu", page_size=4, skip_tokenizer_init=True, chunked_prefill_size=256, disable_radix_cache=True, log_level="error")
    assert type(eng.runner.model).__name__ == "Gemma3ForConditionalGeneration"
    o = eng.generate(input_ids=ids, image_data=imgs, sampling_params={"max_new_tokens": 6, "temperature": 0, "ignore_eos": True},
                     return_logprob=True, logprob_start_len=0)
    # a text-only request on the same engine takes the ordinary causal path
    t = eng.generate(input_ids=[5, 6, 7, 8, 9], sampling_params={"max_new_tokens": 3, "temperature": 0, "ignore_eos": True}, return_logprob=True,
                     logprob_start_len=0)
    exp_ids, mm = P.process_images(ids, imgs, eng.model_config, None)
    eng.shutdown()
    hf_ids = list(exp_ids)
    for off in mm["image_offsets"]:
        hf_ids[off:off + 16] = [500] * 16
    assert len(hf_ids) == o["meta_info"]["prompt_tokens"] == len(ids) + 2 * 15
    full = hf_ids + o["output_ids"]
    with torch.no_grad():
        out = hf(input_ids=torch.tensor([full]), pixel_values=torch.from_numpy(np.asarray(mm["pixel_values"], np.float32)),
                 token_type_ids=torch.tensor([[1 if x == 500 else 0 for x in full]]))
        tl = torch.log_softmax(hf(input_ids=torch.tensor([[5, 6, 7, 8, 9] + t["output_ids"]])).logits[0].float(), -1)
    lp = torch.log_softmax(out.logits[0].float(), -1)
    ref = [float(lp[i, full[i + 1]]) for i in range(len(hf_ids) - 1)]
    got = [v for v, _, _ in o["meta_info"]["input_token_logprobs"][1:]]
    text_next = [i for i in range(len(hf_ids) - 1) if hf_ids[i + 1] != 500]
    assert max(abs(ref[i] - got[i]) for i in text_next) < 2e-3
    n = len(hf_ids)
    for j, tok in enumerate(o["output_ids"]):
        assert int(lp[n - 1 + j].argmax()) == tok and abs(float(lp[n - 1 + j, tok]) - o["meta_info"]["output_token_logprobs"][j][0]) < 2e-3
    for j, tok in enumerate(t["output_ids"]):
        assert abs(float(tl[4 + j, tok]) - t["meta_info"]["output_token_logprobs"][j][0]) < 2e-3
