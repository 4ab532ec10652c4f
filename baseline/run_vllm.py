#!/usr/bin/env python
"""External sanity arm: vLLM 0.22 (as installed in the image, LIBRARY code) on the same synthetic trace, through its
public offline API ``vllm.LLM.generate``.  Not the headline baseline — a second same-box point next to the reference arm."""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "LOCAL_WORLD_SIZE", "GROUP_RANK", "ROLE_RANK", "ROLE_WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT",
          "TORCHELASTIC_RUN_ID", "TORCHELASTIC_RESTART_COUNT", "TORCHELASTIC_MAX_RESTARTS", "TORCHELASTIC_USE_AGENT_STORE",
          "TORCHELASTIC_ERROR_FILE", "OMP_NUM_THREADS"):
    os.environ.pop(k, None)
os.environ.setdefault("FLASHINFER_WORKSPACE_BASE", os.path.join(HERE, "_fi_ws"))
os.environ.setdefault("HF_HUB_OFFLINE", "1")
os.environ.setdefault("TRANSFORMERS_OFFLINE", "1")
os.environ.setdefault("VLLM_NO_USAGE_STATS", "1")
os.environ.setdefault("TOKENIZERS_PARALLELISM", "false")
sys.path.append(ROOT)


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--model", default="llama-3-8b")
    ap.add_argument("--num-prompts", type=int, default=1024)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--max-running", type=int, default=1024)
    a = ap.parse_args()
    import numpy as np

    from bench import ClockSampler, sharegpt_shaped_trace
    from vllm import LLM, SamplingParams

    model_dir = os.path.join(HERE, "models", a.model)
    V = int(json.load(open(os.path.join(model_dir, "config.json")))["vocab_size"])
    pin, pout = sharegpt_shaped_trace(a.num_prompts, a.seed)
    total_out, total_in = int(sum(pout)), int(sum(pin))
    sps = [SamplingParams(max_tokens=m, temperature=0.0, ignore_eos=True, detokenize=False) for m in pout]

    def fresh(step: int):
        r = np.random.default_rng(a.seed + 100 + step)
        return [{"prompt_token_ids": r.integers(300, V - 1000, size=n).tolist()} for n in pin]

    llm = LLM(model=model_dir, load_format="dummy", skip_tokenizer_init=True, tensor_parallel_size=a.gpus, max_num_seqs=a.max_running,
              seed=a.seed, dtype="bfloat16", disable_log_stats=True)
    for w in range(a.warmup):
        llm.generate(fresh(-1 - w), sps, use_tqdm=False)
    sampler = ClockSampler(0)
    sampler.start()
    t0 = time.perf_counter()
    n_out = 0
    for k in range(a.steps):
        outs = llm.generate(fresh(k), sps, use_tqdm=False)
        n_out += sum(len(o.outputs[0].token_ids) for o in outs)
    wall = time.perf_counter() - t0
    sampler.stop_evt.set()
    sampler.join(timeout=2)
    value = n_out / wall
    import vllm

    print(json.dumps({
        "impl": "vllm", "metric": "output_tokens_per_s", "value": value, "unit": "tok/s", "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": 1000.0 * wall / a.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic",
        "config": {"model": a.model, "global_batch": a.num_prompts,
                   "seq_len": f"sharegpt-shaped lognormal: mean in {total_in // a.num_prompts} / out {total_out // a.num_prompts}",
                   "parallelism": f"tp{a.gpus}", "step": "one full trace (request-rate inf, continuous batching)",
                   "vllm_version": vllm.__version__, "max_running": a.max_running, "output_tokens_per_step": total_out,
                   "timing": "wall clock around vllm.LLM.generate; value == e2e.value"},
        "e2e": {"value": value, "unit": "tok/s", "h2d_bytes_per_step": None, "d2h_bytes_per_step": None},
        "gpu_launches": None, "clocks": sampler.summary()}), flush=True)


if __name__ == "__main__":
    main()
