"""Tiny single-kernel drivers for ncu captures: python tools/ncu_target.py <attn_decode|attn_prefill|gemm|gemm_small> [n_iters]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch

from sglang_b200 import ops
from gpu_cases import _attn_setup

ops.load_native()
what = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4
dev = "cuda"
if what.startswith("attn"):
    if what == "attn_decode":
        B, ql, sl, dec = 64, [1] * 64, [512] * 64, True
    else:
        B, ql, sl, dec = 1, [2048], [2048], False
    s = _attn_setup(B, ql, sl, 32, 8, 128, 32)
    out = torch.empty_like(s["q"])
    for _ in range(n):
        ops.paged_attention(s["q"], s["kc"], s["vc"], s["pt"], s["slots"], s["seq"], None if dec else s["cu"], None if dec else s["ql"],
                            max(ql), 32, 128 ** -0.5, impl=0, out=out)
elif what.startswith("gemm"):
    M = 8192 if what == "gemm" else 64
    a = torch.randn(M, 4096, device=dev, dtype=torch.bfloat16)
    w = torch.randn(6144, 4096, device=dev, dtype=torch.bfloat16)
    o = torch.empty(M, 6144, device=dev, dtype=torch.bfloat16)
    for _ in range(n):
        torch.ops.sglb.gemm(a, w, o, None, 0, 0)
elif what == "attn_decode_persist":
    B = 512
    s = _attn_setup(B, [1] * B, [512] * B, 32, 8, 128, 32)
    out = torch.empty_like(s["q"])
    for _ in range(n):
        ops.paged_attention(s["q"], s["kc"], s["vc"], s["pt"], s["slots"], s["seq"], None, None, 1, 32, 128 ** -0.5, impl=0, out=out)
elif what == "skinny":
    a = torch.randn(64, 14336, device=dev, dtype=torch.bfloat16)
    w = torch.randn(4096, 14336, device=dev, dtype=torch.bfloat16) * 0.02
    for _ in range(n):
        ops.gemm_skinny_slabs(a, w)
elif what == "skinny_silu":
    a = torch.randn(64, 4096, device=dev, dtype=torch.bfloat16)
    w = torch.randn(2 * 14336, 4096, device=dev, dtype=torch.bfloat16) * 0.02
    for _ in range(n):
        ops.gemm_skinny(a, w, silu_split=True)
elif what == "fp8":
    from sglang_b200.ops import quant_ops, torch_ref
    x = torch.randn(2048, 4096, device=dev, dtype=torch.bfloat16)
    wq, ws = torch_ref.block_quant_weight_fp8(torch.randn(14336, 4096, device=dev) * 0.02, 128, False)
    xq, xs = quant_ops.per_token_group_quant_fp8(x, 128, False)
    out = torch.empty(2048, 14336, device=dev, dtype=torch.bfloat16)
    for _ in range(n):
        torch.ops.sglb.gemm_fp8_block(xq, xs, wq, ws, out)
elif what == "moe":
    from sglang_b200.ops import moe_ops
    T, E, K, H, I = 2048, 8, 2, 4096, 14336
    x = torch.randn(T, H, device=dev, dtype=torch.bfloat16) * 0.5
    w13 = torch.randn(E, 2 * I, H, device=dev, dtype=torch.bfloat16) * 0.02
    w2 = torch.randn(E, H, I, device=dev, dtype=torch.bfloat16) * 0.02
    tw, ids = moe_ops.topk_softmax(torch.randn(T, E, device=dev), K, True)
    for _ in range(n):
        moe_ops.fused_moe(x, w13, w2, tw, ids, 0)
torch.cuda.synchronize()
print("done", what)
