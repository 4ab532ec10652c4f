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
