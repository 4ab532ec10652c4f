"""Offline throughput through ``Engine.generate`` (reference ``python/sglang/bench_offline_throughput.py``)."""
from __future__ import annotations

import argparse
import json
import time


def main(argv=None):
    ap = argparse.ArgumentParser("bench_offline_throughput")
    ap.add_argument("--model-path", default="dummy:llama-3-8b")
    ap.add_argument("--num-prompts", type=int, default=512)
    ap.add_argument("--dataset-name", default="sharegpt-shaped", choices=["sharegpt-shaped", "random"])
    ap.add_argument("--random-input-len", type=int, default=1024)
    ap.add_argument("--random-output-len", type=int, default=256)
    ap.add_argument("--tp-size", type=int, default=1)
    ap.add_argument("--max-running-requests", type=int, default=512)
    a = ap.parse_args(argv)
    from .bench_serving import sample_random, sample_sharegpt_shaped
    from .entrypoints.engine import Engine

    eng = Engine(model_path=a.model_path, tp_size=a.tp_size, log_level="warning", max_running_requests=a.max_running_requests,
                 cuda_graph_max_bs=a.max_running_requests)
    V = eng.model_config.vocab_size
    reqs = sample_sharegpt_shaped(a.num_prompts, V, 0) if a.dataset_name == "sharegpt-shaped" else \
        sample_random(a.num_prompts, a.random_input_len, a.random_output_len, 0.0, V, 0)
    ids = [r[0] for r in reqs]
    sps = [{"max_new_tokens": r[1], "temperature": 0, "ignore_eos": True} for r in reqs]
    eng.generate(input_ids=ids[:16], sampling_params=sps[:16])
    t0 = time.perf_counter()
    outs = eng.generate(input_ids=ids, sampling_params=sps)
    dt = time.perf_counter() - t0
    n_in, n_out = sum(len(x) for x in ids), sum(o["meta_info"]["completion_tokens"] for o in outs)
    res = {"duration_s": dt, "num_prompts": len(ids), "input_tokens": n_in, "output_tokens": n_out, "request_throughput": len(ids) / dt,
           "input_throughput": n_in / dt, "output_throughput": n_out / dt, "total_throughput": (n_in + n_out) / dt}
    print(json.dumps(res, indent=1))
    return res


if __name__ == "__main__":
    main()
