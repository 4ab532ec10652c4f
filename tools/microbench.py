"""CUDA-event microbenchmarks of the hot ops at Llama-3-8B shapes (prefill chunk and decode batch).
Roofline denominators from MEASURED_PEAKS.json.  Output -> gpurun_out/microbench.json"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch

from sglang_b200 import ops
from gpu_cases import _attn_setup

ops.load_native()
dev = "cuda"
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650, "bf16_tflops": 1590}
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        s.record(); fn(); e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]


res = {}
# ---------------- GEMMs
for name, (N, K) in {"qkv": (6144, 4096), "o": (4096, 4096), "gate_up": (28672, 4096), "down": (4096, 14336)}.items():
    w = torch.randn(N, K, device=dev, dtype=torch.bfloat16)
    for M in (16, 64, 256, 1024, 8192):
        a = torch.randn(M, K, device=dev, dtype=torch.bfloat16)
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        t_lib = timeit(lambda: torch.nn.functional.linear(a, w))
        row = {"lib_ms": t_lib}
        for bn in (32, 64, 128, 256):
            if N % bn:
                continue
            try:
                row[f"tc_bn{bn}_ms"] = timeit(lambda: torch.ops.sglb.gemm(a, w, out, None, 0, bn))
            except Exception as e:
                row[f"tc_bn{bn}_ms"] = str(e)[:80]
        flops = 2.0 * M * N * K
        bytes_ = 2.0 * (M * K + N * K + M * N)
        best = min(v for k, v in row.items() if k.startswith("tc_") and isinstance(v, float))
        row["tc_best_tflops"] = flops / best / 1e9
        row["lib_tflops"] = flops / t_lib / 1e9
        row["tc_best_gbs"] = bytes_ / best / 1e6
        row["roofline_ms"] = max(flops / (peaks["bf16_tflops"] * 1e9), bytes_ / (peaks["hbm_gbs"] * 1e6))
        res[f"gemm_{name}_M{M}"] = row
        print(name, M, {k: (round(v, 4) if isinstance(v, float) else v) for k, v in row.items()}, flush=True)

# ---------------- attention
def attn_case(tag, B, q_lens, seq_lens, decode, splits=1):
    s = _attn_setup(B, q_lens, seq_lens, 32, 8, 128, 32)
    out = torch.empty_like(s["q"])
    ws = None
    if splits > 1:
        ws = (torch.empty(splits, s["T"], 32, 128, dtype=torch.float32, device=dev), torch.empty(splits, s["T"], 32, dtype=torch.float32, device=dev))
    row = {}
    for impl, nm in ((0, "tc"), (1, "simt")):
        if impl == 1 and sum(q_lens) * max(seq_lens) > 3e7:
            continue
        f = lambda: ops.paged_attention(s["q"], s["kc"], s["vc"], s["pt"], s["slots"], s["seq"], None if decode else s["cu"],
                                        None if decode else s["ql"], max(q_lens), 32, 128 ** -0.5, num_splits=splits if impl == 0 else 1,
                                        workspace=ws, impl=impl, out=out)
        row[nm + "_ms"] = timeit(f, iters=5)
    kv_bytes = sum(seq_lens) * 8 * 128 * 2 * 2
    flops = sum(4.0 * 32 * 128 * ql * (sl - ql / 2) for ql, sl in zip(q_lens, seq_lens))
    row["kv_gbs_tc"] = kv_bytes / row["tc_ms"] / 1e6
    row["tflops_tc"] = flops / row["tc_ms"] / 1e9
    row["roofline_ms"] = max(kv_bytes / (peaks["hbm_gbs"] * 1e6), flops / (peaks["bf16_tflops"] * 1e9))
    res["attn_" + tag] = row
    print(tag, {k: round(v, 4) for k, v in row.items()}, flush=True)


import numpy as np
rng = np.random.default_rng(0)
attn_case("decode_bs256_ctx512", 256, [1] * 256, [512] * 256, True)
attn_case("decode_bs256_ragged", 256, [1] * 256, rng.integers(32, 1500, 256).tolist(), True)
attn_case("decode_bs64_ctx2048", 64, [1] * 64, [2048] * 64, True)
attn_case("decode_bs8_ctx4096_split8", 8, [1] * 8, [4096] * 8, True, splits=8)
attn_case("decode_bs8_ctx4096_split1", 8, [1] * 8, [4096] * 8, True, splits=1)
ql = rng.integers(50, 600, 28).tolist()
attn_case("prefill_28seq_8k", 28, ql, ql, False)
attn_case("prefill_1x8192", 1, [8192], [8192], False)
attn_case("prefill_4x2048", 4, [2048] * 4, [2048] * 4, False)

# ---------------- elementwise / sampling
T, H = 8192, 4096
x = torch.randn(T, H, device=dev, dtype=torch.bfloat16); r = torch.randn_like(x); w = torch.ones(H, device=dev, dtype=torch.bfloat16)
t = timeit(lambda: ops.fused_add_rmsnorm(x, r, w, 1e-5)); res["fused_add_rmsnorm_8192"] = {"ms": t, "gbs": 4 * T * H * 2 / t / 1e6}
x2 = x[:256].clone(); r2 = r[:256].clone()
t = timeit(lambda: ops.fused_add_rmsnorm(x2, r2, w, 1e-5)); res["fused_add_rmsnorm_256"] = {"ms": t}
g = torch.randn(T, 28672, device=dev, dtype=torch.bfloat16)
t = timeit(lambda: ops.silu_and_mul(g)); res["silu_mul_8192"] = {"ms": t, "gbs": 3 * T * 14336 * 2 / t / 1e6}
lg = torch.randn(256, 128256, device=dev, dtype=torch.bfloat16)
one = torch.ones(256, device=dev); k1 = torch.ones(256, dtype=torch.int32, device=dev)
t = timeit(lambda: ops.sample(lg, one, k1, one, torch.zeros(256, device=dev))); res["sample_greedy_256"] = {"ms": t}
kk = torch.full((256,), 50, dtype=torch.int32, device=dev)
t = timeit(lambda: ops.sample(lg, one, kk, one * 0.9, torch.zeros(256, device=dev))); res["sample_topk_topp_256"] = {"ms": t}
for k, v in res.items():
    if not k.startswith(("gemm", "attn")):
        print(k, v, flush=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "microbench.json"), "w"), indent=1)
