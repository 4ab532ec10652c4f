"""Static-batch latency benchmark without a server (reference ``python/sglang/bench_one_batch.py``): prefill latency for
(batch, input_len), then per-token decode latency, device-timed.

    python -m sglang_b200.bench_one_batch --model-path dummy:llama-3-8b --batch-size 1 16 64 --input-len 1024 --output-len 64
"""
from __future__ import annotations

import argparse
import json
import time

import numpy as np


def main(argv=None):
    ap = argparse.ArgumentParser("bench_one_batch")
    ap.add_argument("--model-path", default="dummy:llama-3-8b")
    ap.add_argument("--batch-size", type=int, nargs="+", default=[1, 16, 64])
    ap.add_argument("--input-len", type=int, default=1024)
    ap.add_argument("--output-len", type=int, default=32)
    ap.add_argument("--tp-size", type=int, default=1)
    ap.add_argument("--result-file", default=None)
    a = ap.parse_args(argv)
    import torch

    from .entrypoints.engine import Engine

    eng = Engine(model_path=a.model_path, tp_size=a.tp_size, log_level="warning", max_running_requests=max(a.batch_size),
                 cuda_graph_max_bs=max(a.batch_size), disable_radix_cache=True)
    V = eng.model_config.vocab_size
    rng = np.random.default_rng(0)
    rows = []
    for bs in a.batch_size:
        prompts = [rng.integers(10, V - 10, size=a.input_len).tolist() for _ in range(bs)]
        for _ in range(2):
            eng.generate(input_ids=prompts, sampling_params={"max_new_tokens": 1, "temperature": 0, "ignore_eos": True})
        sync = torch.cuda.synchronize if torch.cuda.is_available() else (lambda: None)
        sync(); t0 = time.perf_counter()
        eng.generate(input_ids=prompts, sampling_params={"max_new_tokens": 1, "temperature": 0, "ignore_eos": True})
        sync(); t_prefill = time.perf_counter() - t0
        sync(); t0 = time.perf_counter()
        eng.generate(input_ids=prompts, sampling_params={"max_new_tokens": a.output_len, "temperature": 0, "ignore_eos": True})
        sync(); t_total = time.perf_counter() - t0
        t_dec = max(1e-9, t_total - t_prefill) / max(1, a.output_len - 1)
        row = {"batch_size": bs, "input_len": a.input_len, "output_len": a.output_len, "prefill_latency_s": t_prefill,
               "prefill_throughput_tok_s": bs * a.input_len / t_prefill, "median_decode_latency_s": t_dec,
               "decode_throughput_tok_s": bs / t_dec}
        rows.append(row)
        print(json.dumps(row))
    if a.result_file:
        with open(a.result_file, "a") as f:
            for r in rows:
                f.write(json.dumps(r) + "\n")
    return rows


if __name__ == "__main__":
    main()
