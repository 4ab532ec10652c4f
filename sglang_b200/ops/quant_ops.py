"""fp8 block-scaled linear: activation per-token-group quant + block-scaled GEMM.

GPU path: own quant kernel + own tcgen05 ``kind::f8f6f4`` GEMM with per-128-K-block scale application
(``csrc/kernels/gemm_fp8.cu``) when built; otherwise (and on CPU) the fp32 reference in ``torch_ref``.
Reference: ``python/sglang/srt/layers/quantization/fp8_utils.py:552-808`` (dispatch), ``fp8_kernel.py``.
"""
from __future__ import annotations

import torch

from . import torch_ref


def _has(op: str) -> bool:
    return hasattr(torch.ops, "sglb") and hasattr(torch.ops.sglb, op)


def per_token_group_quant_fp8(x: torch.Tensor, group: int = 128, ue8m0: bool = False):
    if x.is_cuda and _has("per_token_group_quant_fp8") and x.dtype == torch.bfloat16:
        q = torch.empty(x.shape, dtype=torch.float8_e4m3fn, device=x.device)
        s = torch.empty(x.shape[0], x.shape[1] // group, dtype=torch.float32, device=x.device)
        torch.ops.sglb.per_token_group_quant_fp8(x, q, s, group, ue8m0)
        return q, s
    return torch_ref.per_token_group_quant_fp8(x, group, ue8m0)


def fp8_block_linear(x: torch.Tensor, wq: torch.Tensor, ws: torch.Tensor, block: int = 128, ue8m0: bool = False,
                     out_dtype=torch.bfloat16) -> torch.Tensor:
    xq, xs = per_token_group_quant_fp8(x, block, ue8m0)
    if x.is_cuda and _has("gemm_fp8_block"):
        out = torch.empty(x.shape[0], wq.shape[0], dtype=out_dtype, device=x.device)
        torch.ops.sglb.gemm_fp8_block(xq, xs, wq, ws, out)
        return out
    return torch_ref.fp8_block_gemm(xq, xs, wq, ws, block, out_dtype)
