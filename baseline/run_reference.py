#!/usr/bin/env python
"""Reference arm of bench.py: the UNMODIFIED reference (pip --no-deps install of /root/reference/python under baseline/_ref) driven
through its own public API (``sglang.Engine``) on the same synthetic ShareGPT-shaped trace as our arm.

What is *not* the reference here, and why (everything is on a reference-arm-only sys.path, none of it is product code):
  * ``baseline/ref_shims/orjson.py``   json-backed stand-in: the ``orjson`` wheel is not in this offline image;
  * ``baseline/ref_shims/sgl_kernel``  import surface of the ``sgl_kernel`` wheel (no wheel; its CMake build needs network).  The reference's
    hot elementwise ops come from its OWN in-tree JIT kernels (``sglang/kernels/jit``, compiled with nvcc through tvm-ffi — they build and
    load fine here, cache in ``baseline/_ref_jit_cache``); the shim only has to exist so that ``from sgl_kernel import ...`` succeeds, and
    backs a few fall-back ops with library kernels (vLLM ``_C``, FlashInfer) in case a JIT module is unavailable;
  * ``baseline/ref_shims/IPython``     notebook helper imported at module import.
Version pins the image cannot satisfy (torch 2.13 / flashinfer 0.6.17 / tvm-ffi 0.1.11 vs 2.11 / 0.6.11 / 0.1.9) are waived with the
reference's own switch ``SGLANG_SKIP_SGL_KERNEL_VERSION_CHECK=1``.

Modes (tried in order by bench.py until one produces a line; the one that ran is recorded in ``config.reference_mode``):
  stock      the reference's own defaults on sm_100: trtllm_mha attention (FlashInfer trtllm-gen cubins), FlashInfer sampling, CUDA graphs
             (+ --bf16-gemm-backend torch in every mode, see MODES)
  fi         --attention-backend flashinfer
  degraded   --attention-backend triton --sampling-backend pytorch          
"""
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
          "TORCH_NCCL_ASYNC_ERROR_HANDLING", "TORCHELASTIC_ERROR_FILE", "OMP_NUM_THREADS"):
    os.environ.pop(k, None)      # the reference spawns and wires its own TP workers; torchrun's rendezvous must not leak into them
os.environ.setdefault("SGLANG_SKIP_SGL_KERNEL_VERSION_CHECK", "1")
os.environ.setdefault("SGLANG_JIT_CACHE_DIR", os.path.join(HERE, "_ref_jit_cache"))
os.environ.setdefault("FLASHINFER_WORKSPACE_BASE", os.path.join(HERE, "_fi_ws"))
os.environ.setdefault("HF_HUB_OFFLINE", "1")
os.environ.setdefault("TRANSFORMERS_OFFLINE", "1")
os.environ.setdefault("TOKENIZERS_PARALLELISM", "false")
sys.path[:0] = [os.path.join(HERE, "ref_shims"), os.path.join(HERE, "_ref")]
sys.path.append(ROOT)

# every mode carries the reference's own --bf16-gemm-backend torch: its sm_100 default ("cutedsl") imports cutlass.cute.experimental, which the
# image's nvidia-cutlass-dsl refuses below CUDA toolkit 13.1 (this image: 12.9) — F.linear / cuBLAS is the reference's documented alternative
# --mem-fraction-static 0.8: with its own default (0.854 on this part) the reference runs out of memory while capturing its prefill CUDA graphs
# on this image (run D: "Tried to allocate 420 MiB ... 178.2 GiB in use"); lowering the fraction is the reference's documented remedy.
# --cuda-graph-max-bs-prefill 2048: at its default (16384 tokens x 74 shapes) every captured prefill shape took 10-15 GB on this torch build and the
# capture ran the 178 GB card out of memory at ANY mem fraction (runs D / E); 2048 is the value the reference's own OOM message recommends.  The
# `*_nopg` modes (--disable-prefill-cuda-graph) are the second fallback.
_COMMON = {"bf16_gemm_backend": "torch", "mem_fraction_static": 0.8, "cuda_graph_max_bs_prefill": 2048}
_NOPG = {"bf16_gemm_backend": "torch", "mem_fraction_static": 0.8, "disable_prefill_cuda_graph": True}
MODES = {
    "stock": dict(_COMMON),
    "fi": dict(_COMMON, attention_backend="flashinfer"),
    "fi_nopg": dict(_NOPG, attention_backend="flashinfer"),
    "degraded": dict(_COMMON, attention_backend="triton", sampling_backend="pytorch"),
    "degraded_nopg": dict(_NOPG, attention_backend="triton", sampling_backend="pytorch"),
}


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--mode", default="stock", choices=sorted(MODES))
    ap.add_argument("--model", default="llama-3-8b")
    ap.add_argument("--num-prompts", type=int, default=1024)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--max-running", type=int, default=1024)
    a = ap.parse_args()

    import numpy as np
    import torch  # noqa: F401

    from bench import ClockSampler, sharegpt_shaped_trace

    import sglang as sgl

    model_dir = os.path.join(HERE, "models", a.model)
    cfg = json.load(open(os.path.join(model_dir, "config.json")))
    V = int(cfg["vocab_size"])
    pin, pout = sharegpt_shaped_trace(a.num_prompts, a.seed)
    total_out, total_in = int(sum(pout)), int(sum(pin))
    sps = [{"max_new_tokens": m, "temperature": 0, "ignore_eos": True} for m in pout]

    def fresh_ids(step: int):
        r = np.random.default_rng(a.seed + 100 + step)      # identical stream to bench.py's own arm
        return [r.integers(300, V - 1000, size=n).tolist() for n in pin]

    kw = dict(model_path=model_dir, load_format="dummy", skip_tokenizer_init=True, tp_size=a.gpus, max_running_requests=a.max_running,
              random_seed=a.seed, log_level="error")
    kw.update(MODES[a.mode])
    if a.gpus > 1:
        # the reference's custom all-reduce lives in the sgl_kernel wheel (absent offline, see the header): its own switch falls back to NCCL
        kw["disable_custom_all_reduce"] = True
    eng = sgl.Engine(**kw)
    sa = eng.server_args if hasattr(eng, "server_args") else None
    try:
        for w in range(a.warmup):
            eng.generate(input_ids=fresh_ids(-1 - w), sampling_params=sps)
        sampler = ClockSampler(0)
        sampler.start()
        t0 = time.perf_counter()
        n_out = 0
        for k in range(a.steps):
            outs = eng.generate(input_ids=fresh_ids(k), sampling_params=sps)
            n_out += sum(int(o["meta_info"]["completion_tokens"]) for o in outs)
        wall = time.perf_counter() - t0
        sampler.stop_evt.set()
        sampler.join(timeout=2)
        assert n_out == total_out * a.steps, (n_out, total_out * a.steps)
        value = n_out / wall
        line = {
            "impl": "reference", "metric": "output_tokens_per_s", "value": value, "unit": "tok/s", "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": 1000.0 * wall / a.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"model": a.model, "global_batch": a.num_prompts,
                       "seq_len": f"sharegpt-shaped lognormal: mean in {total_in // a.num_prompts} / out {total_out // a.num_prompts}",
                       "parallelism": f"tp{a.gpus}", "step": "one full trace (request-rate inf, continuous batching)",
                       "l2": "inputs_larger_than_L2 (16 GB weights + KV streamed every decode step)",
                       "reference_mode": a.mode, "reference_flags": dict(MODES[a.mode], **({"disable_custom_all_reduce": True} if a.gpus > 1 else {})),
                       "attention_backend": getattr(sa, "attention_backend", None), "sampling_backend": getattr(sa, "sampling_backend", None),
                       "page_size": getattr(sa, "page_size", None), "cuda_graph": not getattr(sa, "disable_cuda_graph", False),
                       "max_running": a.max_running, "input_tokens_per_step": total_in, "output_tokens_per_step": total_out,
                       "timing": "wall clock around sglang.Engine.generate (the reference's scheduler and GPU work live in child "
                                 "processes, so there is no stream to put CUDA events on); value == e2e.value",
                       "shims": "orjson (json), sgl_kernel import surface, IPython stub; version pins waived"},
            "e2e": {"value": value, "unit": "tok/s", "h2d_bytes_per_step": None, "d2h_bytes_per_step": None,
                    "note": "token ids -> sglang.Engine.generate -> token ids; the reference copies its batch metadata H2D and the sampled ids "
                            "D2H every forward step inside this region (bytes not instrumented: stock code path)"},
            "gpu_launches": None, "clocks": sampler.summary()}
        print(json.dumps(line), flush=True)
    finally:
        try:
            eng.shutdown()
        except Exception:
            pass


if __name__ == "__main__":
    main()
