"""Online serving benchmark client.

Parity: reference ``python/sglang/benchmark/serving.py`` (``python -m sglang.bench_serving``: datasets sharegpt / random /
generated-shared-prefix, ``--request-rate`` (Poisson arrivals, inf = all at once), ``--max-concurrency``; metrics from
``calculate_metrics:1095-1230``: request / input / output / total throughput, mean / median / p99 TTFT, TPOT, ITL, e2e).

    python -m sglang_b200.bench_serving --base-url http://127.0.0.1:30000 --dataset-name random --num-prompts 500 \
        --random-input-len 1024 --random-output-len 256 --request-rate inf
"""
from __future__ import annotations

import argparse
import asyncio
import json
import random
import time
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Tuple

import numpy as np


@dataclass
class RequestOutput:
    ok: bool = False
    ttft: float = 0.0
    latency: float = 0.0
    itl: List[float] = field(default_factory=list)
    prompt_len: int = 0
    output_len: int = 0
    error: str = ""


def sample_random(n: int, in_len: int, out_len: int, range_ratio: float, vocab: int, seed: int) -> List[Tuple[List[int], int]]:
    rng = np.random.default_rng(seed)
    ins = rng.integers(max(1, int(in_len * range_ratio)), in_len + 1, size=n)
    outs = rng.integers(max(1, int(out_len * range_ratio)), out_len + 1, size=n)
    return [(rng.integers(10, vocab - 10, size=int(i)).tolist(), int(o)) for i, o in zip(ins, outs)]


def sample_sharegpt_shaped(n: int, vocab: int, seed: int, max_in: int = 2048, max_out: int = 1024) -> List[Tuple[List[int], int]]:
    """ShareGPT-like log-normal lengths (the dataset is not available offline)"""
    rng = np.random.default_rng(seed)
    ins = np.clip(rng.lognormal(5.2, 1.0, n), 4, max_in).astype(int)
    outs = np.clip(rng.lognormal(4.95, 0.9, n), 4, max_out).astype(int)
    return [(rng.integers(10, vocab - 10, size=int(i)).tolist(), int(o)) for i, o in zip(ins, outs)]


def sample_shared_prefix(groups: int, per_group: int, sys_len: int, q_len: int, out_len: int, vocab: int, seed: int):
    rng = np.random.default_rng(seed)
    reqs = []
    for _ in range(groups):
        sysp = rng.integers(10, vocab - 10, size=sys_len).tolist()
        for _ in range(per_group):
            reqs.append((sysp + rng.integers(10, vocab - 10, size=q_len).tolist(), out_len))
    order = rng.permutation(len(reqs))
    return [reqs[i] for i in order]


def sample_random_image(n: int, in_len: int, out_len: int, vocab: int, seed: int, image_token_id: int, images: int, width: int, height: int):
    """vision-language load (reference ``--dataset-name image``): random text around ``images`` placeholder tokens, each paired with a random-pixel
    PNG of the given resolution (data URL) — exercises preprocessing, the vision tower and the embed splice under load"""
    import base64
    import io

    from PIL import Image

    rng = np.random.default_rng(seed)
    reqs = []
    for _ in range(n):
        text = rng.integers(10, vocab - 10, size=in_len).tolist()
        text = [t if t != image_token_id else t + 1 for t in text]
        cut = sorted(rng.integers(0, in_len + 1, size=images).tolist())
        ids, prev = [], 0
        for c in cut:
            ids += text[prev:c] + [image_token_id]
            prev = c
        ids += text[prev:]
        pics = []
        for _ in range(images):
            buf = io.BytesIO()
            Image.fromarray(rng.integers(0, 255, size=(height, width, 3), dtype=np.uint8)).save(buf, format="PNG")
            pics.append("data:image/png;base64," + base64.b64encode(buf.getvalue()).decode())
        reqs.append((ids, out_len, pics))
    return reqs


async def _one(session: Any, url: str, ids: List[int], out_len: int, extra: Dict[str, Any], headers: Dict[str, str], images: Any = None) -> RequestOutput:
    res = RequestOutput(prompt_len=len(ids))
    payload = {"input_ids": ids, "stream": True,
               "sampling_params": {"max_new_tokens": out_len, "temperature": 0.0, "ignore_eos": True, **extra}}
    if images:
        payload["image_data"] = images
    t0 = time.perf_counter()
    last = t0
    n_prev = 0
    try:
        async with session.post(url + "/generate", json=payload, headers=headers) as resp:
            if resp.status != 200:
                res.error = f"HTTP {resp.status}: {(await resp.text())[:200]}"
                return res
            async for raw in resp.content:
                line = raw.decode().strip()
                if not line.startswith("data:") or "[DONE]" in line:
                    continue
                obj = json.loads(line[5:])
                n = obj["meta_info"]["completion_tokens"]
                now = time.perf_counter()
                if n > n_prev:
                    if n_prev == 0:
                        res.ttft = now - t0
                    else:
                        res.itl.extend([(now - last) / (n - n_prev)] * (n - n_prev))
                    last, n_prev = now, n
        res.latency = time.perf_counter() - t0
        res.output_len = n_prev
        res.ok = n_prev > 0
    except Exception as e:  # connection errors are data, not crashes
        res.error = f"{type(e).__name__}: {e}"
    return res


async def run(args: argparse.Namespace) -> Dict[str, Any]:
    import aiohttp

    vocab = args.vocab_size
    if args.dataset_name == "random":
        reqs = sample_random(args.num_prompts, args.random_input_len, args.random_output_len, args.random_range_ratio, vocab, args.seed)
    elif args.dataset_name in ("sharegpt", "sharegpt-shaped"):
        reqs = sample_sharegpt_shaped(args.num_prompts, vocab, args.seed)
    elif args.dataset_name == "generated-shared-prefix":
        reqs = sample_shared_prefix(args.gsp_num_groups, args.gsp_prompts_per_group, args.gsp_system_prompt_len, args.gsp_question_len,
                                    args.gsp_output_len, vocab, args.seed)
    elif args.dataset_name == "random-image":
        w, h = (int(x) for x in args.image_resolution.lower().split("x"))
        reqs = sample_random_image(args.num_prompts, args.random_input_len, args.random_output_len, vocab, args.seed, args.image_token_id,
                                   args.image_count, w, h)
    else:
        raise ValueError(args.dataset_name)
    headers = {"Authorization": f"Bearer {args.api_key}"} if args.api_key else {}
    sem = asyncio.Semaphore(args.max_concurrency) if args.max_concurrency else None
    rng = random.Random(args.seed)
    timeout = aiohttp.ClientTimeout(total=6 * 3600)
    async with aiohttp.ClientSession(timeout=timeout, connector=aiohttp.TCPConnector(limit=0)) as session:
        if args.flush_cache:
            await session.post(args.base_url + "/flush_cache")

        async def guarded(ids, o, imgs=None):
            if sem is None:
                return await _one(session, args.base_url, ids, o, {}, headers, imgs)
            async with sem:
                return await _one(session, args.base_url, ids, o, {}, headers, imgs)

        t0 = time.perf_counter()
        tasks = []
        for r_ in reqs:
            tasks.append(asyncio.create_task(guarded(*r_)))
            if args.request_rate != float("inf"):
                await asyncio.sleep(rng.expovariate(args.request_rate))
        outs = await asyncio.gather(*tasks)
        dur = time.perf_counter() - t0
    ok = [o for o in outs if o.ok]
    pct = lambda xs, q: float(np.percentile(xs, q)) if len(xs) else float("nan")
    ttft = [o.ttft for o in ok]
    tpot = [(o.latency - o.ttft) / (o.output_len - 1) for o in ok if o.output_len > 1]
    itl = [x for o in ok for x in o.itl]
    e2e = [o.latency for o in ok]
    m = {"backend": "sglang_b200", "dataset": args.dataset_name, "request_rate": args.request_rate, "duration_s": dur,
         "completed": len(ok), "failed": len(outs) - len(ok), "total_input_tokens": sum(o.prompt_len for o in ok),
         "total_output_tokens": sum(o.output_len for o in ok), "request_throughput": len(ok) / dur,
         "input_throughput": sum(o.prompt_len for o in ok) / dur, "output_throughput": sum(o.output_len for o in ok) / dur,
         "total_throughput": sum(o.prompt_len + o.output_len for o in ok) / dur,
         "mean_ttft_ms": 1e3 * float(np.mean(ttft)) if ttft else None, "median_ttft_ms": 1e3 * pct(ttft, 50), "p99_ttft_ms": 1e3 * pct(ttft, 99),
         "mean_tpot_ms": 1e3 * float(np.mean(tpot)) if tpot else None, "median_tpot_ms": 1e3 * pct(tpot, 50), "p99_tpot_ms": 1e3 * pct(tpot, 99),
         "mean_itl_ms": 1e3 * float(np.mean(itl)) if itl else None, "median_itl_ms": 1e3 * pct(itl, 50), "p99_itl_ms": 1e3 * pct(itl, 99),
         "mean_e2e_ms": 1e3 * float(np.mean(e2e)) if e2e else None, "median_e2e_ms": 1e3 * pct(e2e, 50),
         "errors": sorted({o.error for o in outs if o.error})[:3]}
    return m


def main(argv: Optional[List[str]] = None) -> Dict[str, Any]:
    ap = argparse.ArgumentParser("sglang_b200.bench_serving")
    ap.add_argument("--base-url", default="http://127.0.0.1:30000")
    ap.add_argument("--dataset-name", default="sharegpt-shaped", choices=["random", "sharegpt", "sharegpt-shaped", "generated-shared-prefix", "random-image"])
    ap.add_argument("--image-token-id", type=int, default=None, help="random-image: the model's image placeholder token id")
    ap.add_argument("--image-count", type=int, default=1)
    ap.add_argument("--image-resolution", default="448x448", help="random-image: WIDTHxHEIGHT of the generated pictures")
    ap.add_argument("--num-prompts", type=int, default=500)
    ap.add_argument("--request-rate", type=float, default=float("inf"))
    ap.add_argument("--max-concurrency", type=int, default=None)
    ap.add_argument("--random-input-len", type=int, default=1024)
    ap.add_argument("--random-output-len", type=int, default=256)
    ap.add_argument("--random-range-ratio", type=float, default=0.0)
    ap.add_argument("--gsp-num-groups", type=int, default=8)
    ap.add_argument("--gsp-prompts-per-group", type=int, default=16)
    ap.add_argument("--gsp-system-prompt-len", type=int, default=1024)
    ap.add_argument("--gsp-question-len", type=int, default=64)
    ap.add_argument("--gsp-output-len", type=int, default=64)
    ap.add_argument("--vocab-size", type=int, default=32000)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--api-key", default=None)
    ap.add_argument("--flush-cache", action="store_true")
    ap.add_argument("--output-file", default=None)
    args = ap.parse_args(argv)
    m = asyncio.run(run(args))
    print(json.dumps(m, indent=1))
    if args.output_file:
        with open(args.output_file, "a") as f:
            f.write(json.dumps(m) + "\n")
    return m


if __name__ == "__main__":
    main()
