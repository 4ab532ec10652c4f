"""Microbenchmarks for the two non-bf16-dense configs BASELINE.json names: block-scaled fp8 linears (Llama-3-70B TP=8 shard shapes and
Llama-3-8B shapes) and the Mixtral-8x7B expert block.  CUDA events, L2 flushed (256 MB write) between iterations, median of 7.
fp8: our tcgen05 kind::f8f6f4 block-scaled GEMM (+ activation quant kernel) vs cuBLAS bf16 on the same logical GEMM and vs
torch._scaled_mm (per-tensor fp8, library) where available.  MoE: native grouped path (align + gather + 2 grouped tcgen05 GEMMs +
combine) vs the sorted per-expert cuBLAS loop (`_fused_moe_sorted`)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from sglang_b200 import ops
from sglang_b200.ops import moe_ops, quant_ops, torch_ref
ops.load_native()
dev = "cuda"
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
hbm = float(peaks.get("hbm_gbs", 6570)); tf = float(peaks.get("bf16_tflops", 1689))
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def t(fn, it=7):
    for _ in range(3): fn()
    ts = []
    for _ in range(it):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); s.record(); fn(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
    return sorted(ts)[len(ts) // 2]
res = {"fp8": {}, "moe": {}}
shapes = {"8b_gate_up": (28672, 4096), "70b_tp8_qkv": (1280, 8192), "70b_tp8_gate_up": (7168, 8192), "70b_tp8_down": (8192, 3584)}
for name, (N, K) in shapes.items():
    w = torch.randn(N, K, device=dev) * 0.02
    wq, ws = torch_ref.block_quant_weight_fp8(w, 128, False)
    wb = w.to(torch.bfloat16)
    wq_t = wq.t()   # column-major view for _scaled_mm
    one = torch.ones(1, device=dev)
    for M in (16, 256, 4096):
        x = (torch.randn(M, K, device=dev) * 0.5).to(torch.bfloat16)
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        xq, xs = quant_ops.per_token_group_quant_fp8(x, 128, False)
        ms_q = t(lambda: quant_ops.per_token_group_quant_fp8(x, 128, False))
        ms_g = t(lambda: torch.ops.sglb.gemm_fp8_block(xq, xs, wq, ws, out))
        ms_b = t(lambda: torch.mm(x, wb.t(), out=out))
        row = {"quant_ms": round(ms_q, 4), "fp8_gemm_ms": round(ms_g, 4), "cublas_bf16_ms": round(ms_b, 4),
               "fp8_tflops": round(2.0 * M * N * K / ms_g / 1e9, 1), "bf16_tflops": round(2.0 * M * N * K / ms_b / 1e9, 1),
               "fp8_weight_gbs": round(N * K / ms_g / 1e6, 1)}
        try:
            ms_s = t(lambda: torch._scaled_mm(xq, wq_t, scale_a=one, scale_b=one, out_dtype=torch.bfloat16))
            row["scaled_mm_pertensor_ms"] = round(ms_s, 4)
        except Exception as e:
            row["scaled_mm_pertensor_ms"] = f"unavailable: {type(e).__name__}"
        res["fp8"][f"{name}_M{M}"] = row
        print("fp8", name, M, row, flush=True)
E, K, H, I = 8, 2, 4096, 14336
w13 = torch.randn(E, 2 * I, H, device=dev, dtype=torch.bfloat16) * 0.02
w2 = torch.randn(E, H, I, device=dev, dtype=torch.bfloat16) * 0.02
for T in (16, 256, 2048, 8192):
    x = torch.randn(T, H, device=dev, dtype=torch.bfloat16) * 0.5
    tw, ids = moe_ops.topk_softmax(torch.randn(T, E, device=dev), K, True)
    ms_n = t(lambda: moe_ops.fused_moe(x, w13, w2, tw, ids, 0))
    ms_l = t(lambda: moe_ops._fused_moe_sorted(x, w13, w2, tw, ids, 0), it=5)
    ms_r = t(lambda: moe_ops.topk_softmax(torch.randn(T, E, device=dev), K, True))
    flops = 2.0 * T * K * 3 * H * I
    wbytes = (w13.numel() + w2.numel()) * 2 * min(1.0, T * K / E / 1.0 if T * K < E else 1.0)
    roof = max(flops / (tf * 1e9), wbytes / (hbm * 1e6))
    row = {"native_ms": round(ms_n, 4), "sorted_cublas_loop_ms": round(ms_l, 4), "router_ms": round(ms_r, 4),
           "native_tflops": round(flops / ms_n / 1e9, 1), "roofline_ms": round(roof, 4), "frac_of_roofline": round(roof / ms_n, 3)}
    res["moe"][f"T{T}"] = row
    print("moe", T, row, flush=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "fp8_moe_bench.json"), "w"), indent=1)
