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
torch.cuda.synchronize()
print("done", what)
