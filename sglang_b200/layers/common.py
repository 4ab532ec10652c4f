"""Norm / rotary / attention layer modules used by every model.

Parity: reference ``python/sglang/srt/layers/layernorm.py:423`` (RMSNorm with fused residual add),
``rotary_embedding/*`` (cos/sin cache, neox & gptj, llama3 scaling), ``radix_attention.py:91`` (model-facing
attention op that delegates to the backend).  Here RoPE, the paged KV-cache write and the attention call are one
layer: ``RadixAttention.forward(q, k, v, fb)`` -> ``ops.rope_kv_write`` (one kernel) -> ``ops.paged_attention``.
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import nn

from .. import ops
from ..model_executor.forward_batch import ForwardBatch
from ..ops import torch_ref


def _rs_residual_active() -> bool:
    from ..parallel.parallel_state import get_state

    st = get_state()
    nvl = getattr(st.tp, "nvl", None) if st is not None else None
    return bool(nvl is not None and nvl.rs_active)


class RMSNorm(nn.Module):
    def __init__(self, hidden_size: int, eps: float = 1e-6, dtype=torch.bfloat16, gemma: bool = False):
        super().__init__()
        self.weight = nn.Parameter(torch.zeros(hidden_size, dtype=dtype) if gemma else torch.ones(hidden_size, dtype=dtype),
                                   requires_grad=False)
        self.eps = eps
        self.gemma = gemma

    def forward(self, x: torch.Tensor, residual: Optional[torch.Tensor] = None):
        """residual given: (in place) residual += x; x = norm(residual); returns (x, residual)"""
        if hasattr(x, "consume_norm"):   # PendingAllReduce / PendingSlabs: (cross-rank | split-K) sum + residual add + norm in ONE kernel
            return x.consume_norm(residual, self.weight, self.eps, self.gemma), residual
        if residual is not None and _rs_residual_active():
            raise RuntimeError("row-sharded residual stream is active (NvlComm.rs_active) but a residual update bypassed the fused "
                               "consume_norm path: this model family must not enable it")
        if residual is None:
            return ops.rmsnorm(x, self.weight, self.eps, gemma=self.gemma)
        ops.fused_add_rmsnorm(x, residual, self.weight, self.eps, gemma=self.gemma)
        return x, residual


class RotaryEmbedding:
    """holds the fp32 cos|sin cache [max_pos, rotary_dim]"""
    _cache = {}

    def __init__(self, head_dim: int, max_position: int, base: float, rope_scaling=None, is_neox: bool = True,
                 partial_rotary_factor: float = 1.0, device="cpu"):
        self.is_neox = is_neox
        rd = int(head_dim * partial_rotary_factor)
        key = (head_dim, rd, max_position, base, str(rope_scaling), str(device))
        if key not in RotaryEmbedding._cache:
            RotaryEmbedding._cache[key] = torch_ref.build_cos_sin_cache(head_dim, max_position, base, rope_scaling, rd).to(device)
        self.cos_sin = RotaryEmbedding._cache[key]


class RadixAttention(nn.Module):
    def __init__(self, num_heads: int, head_dim: int, num_kv_heads: int, layer_id: int, scale: Optional[float] = None,
                 rope: Optional[RotaryEmbedding] = None, sliding_window: int = -1, logit_cap: float = 0.0):
        super().__init__()
        self.num_heads, self.head_dim, self.num_kv_heads, self.layer_id = num_heads, head_dim, num_kv_heads, layer_id
        self.scale = scale if scale is not None else head_dim ** -0.5
        self.rope = rope
        self.window = sliding_window
        self.logit_cap = logit_cap
        self.sinks = None            # [num_heads] fp32 learned attention sinks (GPT-OSS), see forward
        self.chunk = 0               # > 0: attend only inside fixed position chunks (Llama-4 local layers)

    def forward(self, q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, fb: ForwardBatch, nslabs: int = 1,
                slab_stride: int = 0) -> torch.Tensor:
        """q [T, Hq*D], k/v [T, Hkv*D] (views into the fused QKV output; with ``nslabs > 1`` into slab 0 of a split-K QKV GEMM
        whose partials the RoPE kernel sums)"""
        T = q.shape[0]
        q3 = q.view(T, self.num_heads, self.head_dim)
        k3 = k.view(T, self.num_kv_heads, self.head_dim)
        v3 = v.view(T, self.num_kv_heads, self.head_dim)
        pool = fb.kv_pool
        kc, vc = pool.k(self.layer_id), pool.v(self.layer_id)
        dcp = getattr(fb, "dcp", None)
        out_loc = fb.out_cache_loc if dcp is None else dcp.local_loc
        page_table = fb.page_table
        if self.window > 0 and fb.out_cache_loc_swa is not None:      # sliding layer of a hybrid model: its own pool and page-table half
            out_loc = fb.out_cache_loc_swa
            page_table = fb.page_table[:, fb.swa_pt_offset:]
        if self.rope is not None:
            ov = fb.rope_override            # M-RoPE: a per-token cos/sin table indexed by row number instead of the position table
            ops.rope_kv_write(q3, k3, v3, fb.positions if ov is None else ov[1], self.rope.cos_sin if ov is None else ov[0], kc, vc, out_loc,
                              self.rope.is_neox, pool.k_scale, pool.v_scale, nslabs, slab_stride)
        else:
            assert nslabs == 1
            loc = out_loc.long()
            kc[loc] = k3.to(kc.dtype)
            vc[loc] = v3.to(vc.dtype)
        if dcp is not None:
            return self._forward_dcp(q3, kc, vc, fb, dcp, pool)
        cp = getattr(fb, "cp", None)
        if cp is not None:       # context-parallel prefill: the other groups' K/V rows of this layer (layers/cp.py)
            from ..parallel.parallel_state import get_cp_group
            from .cp import gather_write_kv

            gather_write_kv(get_cp_group(), cp, out_loc, kc, vc)
        sinks = self.sinks
        lse = torch.empty(T, self.num_heads, dtype=torch.float32, device=q3.device) if sinks is not None else None
        mask = fb.custom_mask if not self.chunk else chunk_attention_mask(fb, self.chunk)
        if self.window > 0 and fb.custom_mask_sliding is not None:
            mask = fb.custom_mask_sliding
        o = ops.paged_attention(q3, kc, vc, page_table, fb.req_slots, fb.seq_lens, fb.cu_q, fb.q_lens, fb.max_q_len,
                                fb.page_size, self.scale, causal=True, window=self.window, logit_cap=self.logit_cap,
                                k_scale=pool.k_scale, v_scale=pool.v_scale, num_splits=fb.num_splits,
                                workspace=fb.attn_workspace, custom_mask=mask, lse=lse, tree_anc=fb.tree_anc if not self.chunk else None)
        if sinks is not None:       # attention sinks (GPT-OSS): exp(sink_h) joins the soft-max denominator only -> o * sigmoid(lse - sink_h)
            o = (o.float() * torch.sigmoid(lse - sinks[None, :]).unsqueeze(-1)).to(o.dtype)
        return o.view(T, self.num_heads * self.head_dim)


def chunk_attention_mask(fb: ForwardBatch, chunk: int) -> torch.Tensor:
    """flattened per-request [q_len, kv_len] masks "same chunk and causal" (Llama-4 chunked attention), cached on the batch"""
    cache = fb.__dict__.setdefault("_chunk_masks", {})
    if chunk in cache:
        return cache[chunk]
    import numpy as np

    sl = np.asarray(fb.seq_lens_cpu if fb.seq_lens_cpu is not None else fb.seq_lens.cpu().numpy())
    ql = np.ones_like(sl) if fb.cu_q is None else np.asarray(fb.q_lens_cpu)
    parts = []
    for L, n in zip(sl.tolist(), ql.tolist()):
        qp = torch.arange(L - n, L)[:, None]
        kp = torch.arange(L)[None, :]
        parts.append(((kp <= qp) & (kp // chunk == qp // chunk)).reshape(-1))
    m = torch.cat(parts).to(fb.seq_lens.device) if parts else torch.zeros(0, dtype=torch.bool, device=fb.seq_lens.device)
    cache[chunk] = m
    return m


def _radix_forward_dcp(self, q3, kc, vc, fb, dcp, pool):
    """decode context parallel attention (layers/dcp.py): gather the group's query heads, attend to this rank's token shard, merge"""
    from ..parallel.parallel_state import get_dcp_group
    from .dcp import merge_partials

    assert self.window <= 0, "DCP: sliding-window attention is not supported"
    g = get_dcp_group()
    T = q3.shape[0]
    qg = g.all_gather(q3.contiguous(), dim=1)                              # [T, Hg, D]
    Hg = qg.shape[1]
    lse = torch.full((T, Hg), float("-inf"), dtype=torch.float32, device=q3.device)
    kw = dict(window=-1, logit_cap=self.logit_cap, k_scale=pool.k_scale, v_scale=pool.v_scale)
    if fb.mode.is_decode() and fb.cu_q is None:
        o = torch.empty_like(qg)
        ops.paged_attention(qg, kc, vc, fb.page_table, fb.req_slots, dcp.local_lens, None, None, 1, dcp.local_page_size, self.scale,
                            causal=True, out=o, lse=lse, **kw)
        lse = torch.where(dcp.empty.unsqueeze(-1), torch.full_like(lse, float("-inf")), lse)
    else:
        o = torch.zeros_like(qg)
        if dcp.v_rows.numel():
            qv = qg[dcp.v_rows].contiguous()
            ov = torch.empty_like(qv)
            lv = torch.empty(qv.shape[0], Hg, dtype=torch.float32, device=q3.device)
            ops.paged_attention(qv, kc, vc, fb.page_table, dcp.v_slots, dcp.v_seq_lens, dcp.v_cu_q, dcp.v_q_lens, dcp.v_max_q,
                                dcp.local_page_size, self.scale, causal=True, out=ov, lse=lv, **kw)
            o[dcp.v_rows] = ov
            lse[dcp.v_rows] = lv
    out = merge_partials(g, o, lse)
    return out.to(q3.dtype).reshape(T, self.num_heads * self.head_dim)


RadixAttention._forward_dcp = _radix_forward_dcp


def get_act(name: str):
    """plain (non-gated) activations"""
    import torch.nn.functional as F

    if name == "relu2":
        return lambda x: torch.relu(x).square()
    if name == "xielu":          # learnable per layer: the owning block installs its own module (models/families2.py XIELU)
        return None
    if name in ("gelu_pytorch_tanh", "gelu_tanh", "gelu_new", "gelu_fast"):
        return lambda x: F.gelu(x, approximate="tanh")
    return {"gelu": F.gelu, "relu": torch.relu, "silu": F.silu}[name]


def get_act_and_mul(name: str):
    if name == "silu":
        return ops.silu_and_mul
    if name in ("gelu_pytorch_tanh", "gelu_tanh", "gelu_new"):
        return lambda x: ops.gelu_and_mul(x, "tanh")
    if name == "gelu":
        return lambda x: ops.gelu_and_mul(x, "none")
    raise ValueError(f"unsupported activation {name}")
