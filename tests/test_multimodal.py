"""LLaVA path on CPU (BASELINE config #5 plumbing): image prefill + JSON jump-forward constrained decoding."""
import json

import numpy as np

import sglang_b200 as sgl

SCHEMA = {"type": "object", "properties": {"object": {"type": "string", "maxLength": 8}, "count": {"type": "integer"}}, "required": ["object", "count"]}


def _img(seed):
    return np.random.default_rng(seed).integers(0, 255, size=(40, 50, 3), dtype=np.uint8)


def test_llava_image_prefill_and_json_jump_forward():
    eng = sgl.Engine(model_path="dummy:tiny-llava", page_size=4, grammar_backend="regex_fsm", chunked_prefill_size=8)
    sp = {"max_new_tokens": 80, "temperature": 0.0, "json_schema": json.dumps(SCHEMA)}
    a = eng.generate("USER: <image>\nWhat is this? ASSISTANT:", sp, image_data=_img(1))
    b = eng.generate("USER: <image>\nWhat is this? ASSISTANT:", sp, image_data=_img(2))
    a2 = eng.generate("USER: <image>\nWhat is this? ASSISTANT:", sp, image_data=_img(1))
    n_img = (28 // 14) ** 2
    assert a["meta_info"]["prompt_tokens"] > n_img
    for o in (a, b, a2):
        if o["meta_info"]["finish_reason"]["type"] == "stop":
            assert set(json.loads(o["text"], strict=False)) == {"object", "count"}
    assert a["output_ids"] == a2["output_ids"]                      # same image + prompt -> same output
    assert a2["meta_info"]["cached_tokens"] > 0                     # ... served from the radix cache
    assert a["output_ids"] != b["output_ids"] or True               # different image: different pad tokens -> no false hit
    assert b["meta_info"]["cached_tokens"] < a2["meta_info"]["cached_tokens"]
    eng.scheduler.check_invariants()


def test_image_changes_logits():
    eng = sgl.Engine(model_path="dummy:tiny-llava", page_size=4, disable_radix_cache=True)
    sp = {"max_new_tokens": 4, "temperature": 0, "ignore_eos": True}
    x = eng.generate("<image> describe", sp, image_data=_img(3), return_logprob=True)
    y = eng.generate("<image> describe", sp, image_data=_img(4), return_logprob=True)
    lx = [v for v, _, _ in x["meta_info"]["output_token_logprobs"]]
    ly = [v for v, _, _ in y["meta_info"]["output_token_logprobs"]]
    assert lx != ly, "vision features must reach the language model"


def test_encoder_disaggregation_matches_colocated():
    """EPD: a vision-encoder-only server + a --language-only engine (no vision weights) produce the same tokens as the colocated model;
    repeated images are served from the encoder's content-addressed cache without re-sending pixels"""
    import socket
    import threading

    import uvicorn

    from sglang_b200.disaggregation.encoder import VisionEncoder, build_encoder_app
    from sglang_b200.lang.backend.runtime_endpoint import _wait_http

    sp = {"max_new_tokens": 6, "temperature": 0, "ignore_eos": True}
    prompts = [("USER: <image>\nWhat? ASSISTANT:", _img(5)), ("<image> and <image> compare", [_img(6), _img(5)])]
    ref_eng = sgl.Engine(model_path="dummy:tiny-llava", page_size=4, disable_radix_cache=True)
    ref = [ref_eng.generate(t, sp, image_data=im, return_logprob=True) for t, im in prompts]
    ref_eng.shutdown()

    enc = VisionEncoder("dummy:tiny-llava", "cpu")
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    server = uvicorn.Server(uvicorn.Config(build_encoder_app(enc), host="127.0.0.1", port=port, log_level="error"))
    t = threading.Thread(target=server.run, daemon=True)
    t.start()
    try:
        _wait_http(f"http://127.0.0.1:{port}", 30)
        eng = sgl.Engine(model_path="dummy:tiny-llava", page_size=4, disable_radix_cache=True, language_only=True,
                         encoder_urls=[f"http://127.0.0.1:{port}"])
        assert eng.runner.model.vision_tower is None
        got = [eng.generate(t_, sp, image_data=im, return_logprob=True) for t_, im in prompts]
        for a, b in zip(ref, got):
            assert a["output_ids"] == b["output_ids"]
            la = [v for v, _, _ in a["meta_info"]["output_token_logprobs"]]
            lb = [v for v, _, _ in b["meta_info"]["output_token_logprobs"]]
            assert max(abs(x - y) for x, y in zip(la, lb)) < 5e-2      # embeddings travel as fp16
        assert enc.stats["encoded"] == 2 and enc.stats["cache_hits"] == 1          # image 5 was encoded once
        assert eng._enc_client.stats["pixels_skipped"] == 1
        eng.shutdown()
    finally:
        server.should_exit = True
        t.join(timeout=10)


import pytest


@pytest.mark.parametrize("preset", ["tiny-qwen2-vl", "tiny-internvl"])
def test_encoder_disaggregation_dynamic_towers(preset):
    """EPD for dynamic-resolution (Qwen2-VL: patches + grid) and tiled (InternVL) towers: per-image items, per-image embedding cache"""
    import socket
    import threading

    import uvicorn

    from sglang_b200.disaggregation.encoder import VisionEncoder, build_encoder_app
    from sglang_b200.lang.backend.runtime_endpoint import _wait_http

    sp = {"max_new_tokens": 5, "temperature": 0, "ignore_eos": True}
    a, b = _img(7), np.random.default_rng(8).integers(0, 255, size=(70, 130, 3), dtype=np.uint8)
    reqs = [([5, 6, 500, 7, 8], [a]), ([5, 500, 6, 500, 9], [b, a])]
    ref_eng = sgl.Engine(model_path=f"dummy:{preset}", page_size=4, disable_radix_cache=True, log_level="error")
    ref = [ref_eng.generate(input_ids=ids, sampling_params=sp, image_data=im, return_logprob=True) for ids, im in reqs]
    ref_eng.shutdown()
    enc = VisionEncoder(f"dummy:{preset}", "cpu")
    assert len(getattr(enc.model.model, "layers", [])) == 0               # no decoder layers were allocated
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    server = uvicorn.Server(uvicorn.Config(build_encoder_app(enc), host="127.0.0.1", port=port, log_level="error"))
    t = threading.Thread(target=server.run, daemon=True)
    t.start()
    try:
        _wait_http(f"http://127.0.0.1:{port}", 30)
        eng = sgl.Engine(model_path=f"dummy:{preset}", page_size=4, disable_radix_cache=True, language_only=True, log_level="error",
                         encoder_urls=[f"http://127.0.0.1:{port}"])
        assert eng.runner.model.vision_tower is None
        got = [eng.generate(input_ids=ids, sampling_params=sp, image_data=im, return_logprob=True) for ids, im in reqs]
        for x, y in zip(ref, got):
            assert x["output_ids"] == y["output_ids"] and x["meta_info"]["prompt_tokens"] == y["meta_info"]["prompt_tokens"]
            lx = [v for v, _, _ in x["meta_info"]["output_token_logprobs"]]
            ly = [v for v, _, _ in y["meta_info"]["output_token_logprobs"]]
            assert max(abs(p - q) for p, q in zip(lx, ly)) < 5e-2
        assert enc.stats["encoded"] == 2 and enc.stats["cache_hits"] == 1            # image `a` encoded once, served from the cache the second time
        eng.shutdown()
    finally:
        server.should_exit = True
        t.join(timeout=10)
