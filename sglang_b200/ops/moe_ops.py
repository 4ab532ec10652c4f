"""MoE ops: routing (top-k softmax), token alignment by expert, grouped expert MLP.

GPU path: ``topk_softmax`` / ``moe_align`` / ``grouped_gemm`` from ``csrc/kernels/moe.cu`` when built (own kernels:
warp-per-token routing, counting-sort alignment to 128-row tiles, tcgen05 grouped GEMM that walks a tile->expert
map); otherwise a sort + per-expert GEMM loop.  CPU: ``torch_ref.fused_moe``.
Reference: ``moe_topk_softmax_kernels.cu:628-715``, ``moe_align_kernel.cu:350-386``, ``fused_moe_triton_kernels.py:324``.
"""
from __future__ import annotations

from typing import Tuple

import torch

from . import silu_and_mul, linear, torch_ref


def _has(op: str) -> bool:
    return hasattr(torch.ops, "sglb") and hasattr(torch.ops.sglb, op)


def topk_softmax(router_logits: torch.Tensor, top_k: int, renormalize: bool = True) -> Tuple[torch.Tensor, torch.Tensor]:
    if router_logits.is_cuda and _has("moe_topk_softmax"):
        T = router_logits.shape[0]
        w = torch.empty(T, top_k, dtype=torch.float32, device=router_logits.device)
        ids = torch.empty(T, top_k, dtype=torch.int32, device=router_logits.device)
        torch.ops.sglb.moe_topk_softmax(router_logits, w, ids, renormalize)
        return w, ids
    return torch_ref.moe_topk_softmax(router_logits, top_k, renormalize)


def fused_moe(x: torch.Tensor, w13: torch.Tensor, w2: torch.Tensor, topk_w: torch.Tensor, topk_ids: torch.Tensor,
              expert_offset: int = 0) -> torch.Tensor:
    if not x.is_cuda:
        return torch_ref.fused_moe(x, w13, w2, topk_w, topk_ids, expert_offset)
    if _has("moe_grouped_mlp"):
        out = torch.empty_like(x)
        torch.ops.sglb.moe_grouped_mlp(x, w13, w2, topk_w, topk_ids, out, expert_offset)
        return out
    return _fused_moe_sorted(x, w13, w2, topk_w, topk_ids, expert_offset)


def _fused_moe_sorted(x, w13, w2, topk_w, topk_ids, expert_offset):
    """sort (token, k) pairs by expert, run one GEMM pair per local expert on its contiguous slice, scatter-add."""
    T, H = x.shape
    E = w13.shape[0]
    K = topk_ids.shape[1]
    flat = topk_ids.reshape(-1).long() - expert_offset
    local = (flat >= 0) & (flat < E)
    key = torch.where(local, flat, torch.full_like(flat, E))
    order = torch.argsort(key, stable=True)
    counts = torch.bincount(key, minlength=E + 1)[:E].tolist()   # host sync: acceptable off the graph path
    tok = (order // K)
    xs = x.index_select(0, tok[: sum(counts)])
    ws = topk_w.reshape(-1)[order[: sum(counts)]].to(torch.float32)
    out = torch.zeros(T, H, dtype=torch.float32, device=x.device)
    off = 0
    for e, c in enumerate(counts):
        if c == 0:
            continue
        h = linear(xs[off:off + c], w13[e])
        h = silu_and_mul(h)
        y = linear(h, w2[e]).float() * ws[off:off + c, None]
        out.index_add_(0, tok[off:off + c], y)
        off += c
    return out.to(x.dtype)
