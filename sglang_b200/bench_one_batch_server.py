"""Static-batch latency benchmark THROUGH a running server (reference ``python/sglang/bench_one_batch_server.py``): for every (batch size,
input length, output length) one streaming ``/generate`` call carrying the whole batch; reports time to first token of the batch,
inter-token latency and input / output throughput, optionally after flushing the prefix cache so that every point starts cold.

    python -m sglang_b200.bench_one_batch_server --base-url http://127.0.0.1:30000 --batch-size 1 16 64 --input-len 1024 --output-len 16 64
"""
from __future__ import annotations

import argparse
import itertools
import json
import time
from typing import Any, Dict, List

import numpy as np
import requests


def run_one(base_url: str, bs: int, in_len: int, out_len: int, vocab: int, rng, flush: bool = True, timeout: float = 600.0) -> Dict[str, Any]:
    if flush:
        try:
            requests.post(base_url + "/flush_cache", timeout=30)
        except requests.RequestException:
            pass
    prompts = [rng.integers(10, max(11, vocab - 10), size=in_len).tolist() for _ in range(bs)]
    body = {"input_ids": prompts, "sampling_params": {"max_new_tokens": out_len, "temperature": 0, "ignore_eos": True}, "stream": True}
    t0 = time.perf_counter()
    first: Dict[Any, float] = {}
    last = t0
    done_tokens = 0
    counts: Dict[Any, int] = {}
    with requests.post(base_url + "/generate", json=body, stream=True, timeout=timeout) as r:
        r.raise_for_status()
        for line in r.iter_lines():
            if not line.startswith(b"data:") or b"[DONE]" in line:
                continue
            now = time.perf_counter()
            o = json.loads(line[5:])
            idx = o.get("index", o["meta_info"].get("id"))          # one stream carries the whole batch: chunks are told apart by request
            n = int(o["meta_info"]["completion_tokens"])
            if n > 0 and idx not in first:
                first[idx] = now - t0
            counts[idx] = n
            last = now
    total = last - t0
    done_tokens = sum(counts.values())
    ttft = max(first.values()) if first else float("nan")           # the batch's first step is complete when its slowest request has a token
    decode_s = max(1e-9, total - ttft)
    return {"batch_size": bs, "input_len": in_len, "output_len": out_len, "latency_s": round(total, 4), "ttft_s": round(ttft, 4),
            "itl_ms": round(1e3 * decode_s / max(1, out_len - 1), 3), "input_throughput_tok_s": round(bs * in_len / max(ttft, 1e-9), 1),
            "output_throughput_tok_s": round((done_tokens - len(first)) / decode_s, 1) if out_len > 1 else 0.0,
            "overall_throughput_tok_s": round((bs * in_len + done_tokens) / total, 1), "completed_tokens": done_tokens}


def main(argv=None) -> List[Dict[str, Any]]:
    ap = argparse.ArgumentParser("bench_one_batch_server")
    ap.add_argument("--base-url", default="http://127.0.0.1:30000")
    ap.add_argument("--batch-size", type=int, nargs="+", default=[1, 16])
    ap.add_argument("--input-len", type=int, nargs="+", default=[1024])
    ap.add_argument("--output-len", type=int, nargs="+", default=[16])
    ap.add_argument("--no-flush-cache", action="store_true")
    ap.add_argument("--skip-warmup", action="store_true")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--result-file", default=None)
    a = ap.parse_args(argv)
    info = requests.get(a.base_url + "/get_server_info", timeout=30).json()
    vocab = int((info.get("model_config") or {}).get("vocab_size") or info.get("vocab_size") or 32000)
    rng = np.random.default_rng(a.seed)
    if not a.skip_warmup:
        run_one(a.base_url, 1, 16, 4, vocab, rng)
    rows = []
    for bs, il, ol in itertools.product(a.batch_size, a.input_len, a.output_len):
        row = run_one(a.base_url, bs, il, ol, vocab, rng, flush=not a.no_flush_cache)
        rows.append(row)
        print(json.dumps(row), flush=True)
        if a.result_file:
            with open(a.result_file, "a") as f:
                f.write(json.dumps(row) + "\n")
    return rows


if __name__ == "__main__":
    main()
