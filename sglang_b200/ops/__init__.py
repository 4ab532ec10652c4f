"""Op front door.  CUDA tensors -> our sm_100a kernels (``torch.ops.sglb.*`` from ``_C.so``); CPU tensors ->
the fp32 torch reference (``torch_ref``) so the scheduler/model stack is testable without a GPU.

There is no multi-backend dispatch on the GPU: if the extension is missing on a CUDA box the ops raise
(``SGLB_REQUIRE_NATIVE``), they never fall back silently (the driver records which .so files were loaded).
"""
from __future__ import annotations

import os
from pathlib import Path
from typing import Optional, Tuple

import torch

from ..environ import envs
from . import torch_ref

_SO = Path(__file__).resolve().parent.parent / "_C.so"
_loaded = False
_load_error: Optional[str] = None


def load_native(build_if_missing: bool = False) -> bool:
    global _loaded, _load_error
    if _loaded:
        return True
    if not _SO.exists() and build_if_missing:
        from .. import build as _b

        _b.build_all()
    try:
        torch.ops.load_library(str(_SO))
        _loaded = True
    except Exception as e:  # pragma: no cover - depends on the box
        _load_error = f"{type(e).__name__}: {e}"
    return _loaded


def native_available() -> bool:
    return load_native()


def _use_native(t: torch.Tensor) -> bool:
    if not t.is_cuda or envs.SGLB_FORCE_TORCH_OPS.get():
        return False
    if not load_native():
        if envs.SGLB_REQUIRE_NATIVE.get():
            raise RuntimeError(f"sglang_b200 CUDA extension not loaded ({_load_error}); run `python -m sglang_b200.build`")
        return False
    return True


def launch_count() -> int:
    return int(torch.ops.sglb.launch_count()) if _loaded else 0


def reset_launch_count() -> None:
    if _loaded:
        torch.ops.sglb.reset_launch_count()


# ------------------------------------------------------------------ norms / activations
def rmsnorm(x: torch.Tensor, weight: torch.Tensor, eps: float, out: Optional[torch.Tensor] = None, gemma: bool = False) -> torch.Tensor:
    if _use_native(x) and x.dtype == torch.bfloat16:
        shp = x.shape
        x2 = x.reshape(-1, shp[-1])
        out = torch.empty_like(x2) if out is None else out.view(-1, shp[-1])
        torch.ops.sglb.rmsnorm(x2, None, weight, out, eps, gemma)
        return out.view(shp)
    r = torch_ref.gemma_rmsnorm(x, weight, eps) if gemma else torch_ref.rmsnorm(x, weight, eps)
    if out is not None:
        out.copy_(r)
        return out
    return r


def fused_add_rmsnorm(x: torch.Tensor, residual: torch.Tensor, weight: torch.Tensor, eps: float, gemma: bool = False) -> None:
    """residual += x ; x = rmsnorm(residual) * w   (in place)"""
    if _use_native(x) and x.dtype == torch.bfloat16:
        torch.ops.sglb.rmsnorm(x, residual, weight, x, eps, gemma)
        return
    if gemma:
        s = x.float() + residual.float()
        residual.copy_(s.to(residual.dtype))
        x.copy_(torch_ref.gemma_rmsnorm(residual, weight, eps))
    else:
        torch_ref.fused_add_rmsnorm(x, residual, weight, eps)


def silu_and_mul(x: torch.Tensor) -> torch.Tensor:
    if _use_native(x) and x.dtype == torch.bfloat16:
        out = torch.empty(x.shape[0], x.shape[1] // 2, dtype=x.dtype, device=x.device)
        torch.ops.sglb.act_and_mul(x, out, 0)
        return out
    return torch_ref.silu_and_mul(x)


def gelu_and_mul(x: torch.Tensor, approximate: str = "none") -> torch.Tensor:
    if _use_native(x) and x.dtype == torch.bfloat16:
        out = torch.empty(x.shape[0], x.shape[1] // 2, dtype=x.dtype, device=x.device)
        torch.ops.sglb.act_and_mul(x, out, 2 if approximate == "tanh" else 1)
        return out
    return torch_ref.gelu_and_mul(x, approximate)


# ------------------------------------------------------------------ rope + KV write
def rope_kv_write(q, k, v, positions, cos_sin, k_cache, v_cache, out_cache_loc, is_neox: bool = True,
                  k_scale: float = 1.0, v_scale: float = 1.0, nslabs: int = 1, slab_stride: int = 0) -> None:
    """``nslabs > 1``: q/k/v are slab 0 of a split-K QKV GEMM; the kernel sums the slabs while it rotates (q is written back
    into slab 0, k/v go to the cache)."""
    if _use_native(q) and q.dtype == torch.bfloat16:
        torch.ops.sglb.rope_kv_write(q, k, v, positions, cos_sin, k_cache, v_cache, out_cache_loc, is_neox, k_scale, v_scale,
                                     nslabs, slab_stride)
        return
    assert nslabs == 1
    torch_ref.rope_kv_write(q, k, v, positions, cos_sin, k_cache, v_cache, out_cache_loc, is_neox, k_scale, v_scale)


# ------------------------------------------------------------------ attention
def attn_impl() -> int:
    v = envs.SGLB_ATTN_IMPL.get()
    return 1 if v == "simt" else 0


def paged_attention(q, k_cache, v_cache, page_table, req_slots, seq_lens, cu_q, q_lens, max_q_len: int, page_size: int,
                    scale: float, causal: bool = True, window: int = -1, logit_cap: float = 0.0, k_scale: float = 1.0,
                    v_scale: float = 1.0, num_splits: int = 1, workspace=None, tok2req=None, out=None, lse=None,
                    impl: Optional[int] = None, custom_mask=None, tree_anc=None):
    """q [T,Hq,D]; caches [slots,Hkv,D]; cu_q None => decode (q_len 1).  ``tree_anc`` int32 [T]: speculative token-tree verification on the
    tcgen05 kernel (bit k = draft node k is an ancestor-or-self of this query node); ``custom_mask`` = the general (reference-path) form."""
    tree_ok = (tree_anc is not None and _use_native(q) and q.dtype == torch.bfloat16 and attn_impl() == 0 and q.shape[-1] == 128
               and k_cache.dtype == torch.bfloat16 and 128 % (q.shape[1] // k_cache.shape[1]) == 0 and int(max_q_len) <= 32 and window <= 0)
    if _use_native(q) and q.dtype == torch.bfloat16 and (tree_ok or (custom_mask is None and tree_anc is None)):
        if out is None:
            out = torch.empty(q.shape, dtype=q.dtype, device=q.device)
        impl = attn_impl() if impl is None else impl
        D = q.shape[-1]
        Hq, Hkv = q.shape[1], k_cache.shape[1]
        G = Hq // Hkv
        if impl == 0 and (D != 128 or k_cache.dtype != torch.bfloat16 or 128 % G != 0 or page_size < 8 or page_size > 128):
            impl = 1
        part_o = part_lse = None
        if impl in (0, 2) and num_splits > 1:
            if workspace is None:
                T = q.shape[0]
                part_o = torch.empty(num_splits, T, Hq, D, dtype=torch.float32, device=q.device)
                part_lse = torch.empty(num_splits, T, Hq, dtype=torch.float32, device=q.device)
            else:
                part_o, part_lse = workspace
        else:
            num_splits = 1
        if impl == 1 and cu_q is not None and tok2req is None:
            ql = cu_q[1:] - cu_q[:-1]
            tok2req = torch.repeat_interleave(torch.arange(ql.numel(), device=q.device, dtype=torch.int32), ql.long()).to(torch.int32)
        if tree_ok:
            impl, num_splits, part_o, part_lse = 0, 1, None, None
        if impl == 0 and envs.SGLB_ATTN_PERSIST.get():
            impl = 2        # same kernel family, persistent CTAs with an atomic work queue (falls back to the classic launch when small)
        torch.ops.sglb.attn_fwd(q, k_cache, v_cache, out, lse, part_o, part_lse, page_table, req_slots, seq_lens, cu_q, tok2req,
                                int(max_q_len), int(page_size), int(num_splits), float(scale), bool(causal), int(window),
                                float(logit_cap), float(k_scale), float(v_scale), int(impl), tree_anc if tree_ok else None)
        return out
    if custom_mask is None and tree_anc is not None:
        custom_mask = torch_ref.tree_anc_to_mask(tree_anc, q_lens, seq_lens)
    r = torch_ref.paged_attention(q, k_cache, v_cache, page_table, req_slots, seq_lens, q_lens, page_size, scale, causal,
                                  window, logit_cap, k_scale, v_scale, custom_mask=custom_mask, return_lse=lse is not None)
    if lse is not None:
        o, l = r
        lse.copy_(l)
        r = o
    if out is not None:
        out.copy_(r)
        return out
    return r


# ------------------------------------------------------------------ GEMM
EPI_NONE, EPI_BIAS, EPI_SILU_MUL, EPI_PUSH, EPI_SILU_SPLIT, EPI_ROWSCALE = 0, 1, 2, 3, 4, 5


def gemm_impl() -> str:
    return envs.SGLB_GEMM_IMPL.get()


def linear(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None, impl: Optional[str] = None,
           out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x [T,K] @ w[N,K]^T.  impl: 'tc' = our tcgen05 GEMM, 'lib' = cuBLASLt (plain library GEMM)."""
    impl = impl or gemm_impl()
    if impl != "lib" and streamk_ok(x, w) and (out is None or (out.dim() == 2 and out.stride(1) == 1 and out.stride(0) % 8 == 0
                                                                 and out.data_ptr() % 16 == 0)):
        # decode shapes: persistent stream-K weight streaming (gemm_streamk.cu) — also writes straight into an NVL arena slot (``out``)
        return gemm_streamk(x, w, bias, out=out)
    if impl in ("tc", "auto") and _use_native(x) and x.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 \
            and x.dim() == 2 and x.shape[1] % 8 == 0 and w.shape[0] % 8 == 0 and x.stride(0) % 8 == 0 and x.stride(1) == 1 \
            and impl == "tc":
        if out is None:
            out = torch.empty(x.shape[0], w.shape[0], dtype=x.dtype, device=x.device)
        torch.ops.sglb.gemm(x, w, out, bias, EPI_BIAS if bias is not None else EPI_NONE, 0)
        return out
    if out is not None and bias is None and x.dim() == 2 and out.is_contiguous():
        return torch.mm(x, w.t(), out=out)       # cuBLAS writes straight into the caller's buffer (e.g. an NVL arena slot)
    r = torch.nn.functional.linear(x, w, bias)
    if out is not None:
        out.copy_(r)
        return out
    return r


def gemm_tc(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None, epilogue: int = EPI_NONE, bn: int = 0,
            out: Optional[torch.Tensor] = None) -> torch.Tensor:
    n_out = w.shape[0] // 2 if epilogue in (EPI_SILU_MUL, EPI_SILU_SPLIT) else w.shape[0]
    if out is None:
        out = torch.empty(x.shape[0], n_out, dtype=x.dtype, device=x.device)
    torch.ops.sglb.gemm(x, w, out, bias, epilogue, bn)
    return out


# ------------------------------------------------------------------ decode-shape ("skinny") GEMM: swap-AB + split-K slabs
SKINNY_MAX_M = 256


def skinny_ok(x: torch.Tensor, w: torch.Tensor) -> bool:
    """decode-shape policy: tokens <= 256, bf16, 16-byte aligned rows (``SGLB_SKINNY=0`` disables, ``1`` forces the check only)"""
    return (x.is_cuda and x.dim() == 2 and 0 < x.shape[0] <= SKINNY_MAX_M and x.dtype == torch.bfloat16
            and w.dtype == torch.bfloat16 and w.dim() == 2 and x.stride(1) == 1 and x.stride(0) % 8 == 0 and x.shape[1] % 8 == 0
            and w.shape[0] % 8 == 0 and _use_native(x))


def gemm_skinny(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None, silu_split: bool = False,
                out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x[T<=256,K] @ w[N,K]^T in one pass (no K split).  ``silu_split``: w = [gate; up] -> silu(x@gate^T) * (x@up^T)"""
    n_out = w.shape[0] // 2 if silu_split else w.shape[0]
    if out is None:
        out = torch.empty(x.shape[0], n_out, dtype=x.dtype, device=x.device)
    torch.ops.sglb.gemm_skinny(x, w, out, bias, EPI_SILU_SPLIT if silu_split else (EPI_BIAS if bias is not None else EPI_NONE), 1)
    return out


def streamk_ok(x: torch.Tensor, w: torch.Tensor) -> bool:
    """shape / layout policy of the stream-K decode GEMM (``SGLB_STREAMK_MAX_M`` tokens at most, bf16, TMA-compatible strides)"""
    mx = envs.SGLB_STREAMK_MAX_M.get()
    return (mx > 0 and x.is_cuda and x.dim() == 2 and 0 < x.shape[0] <= min(mx, 256) and x.dtype == torch.bfloat16 and w.dtype == torch.bfloat16
            and w.dim() == 2 and x.stride(1) == 1 and w.stride(1) == 1 and x.stride(0) % 8 == 0 and w.stride(0) % 8 == 0 and x.shape[1] % 8 == 0
            and w.shape[0] % 8 == 0 and x.data_ptr() % 16 == 0 and w.data_ptr() % 16 == 0 and _use_native(x))


def gemm_streamk(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None, silu_split: bool = False,
                 out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x[T<=256,K] @ w[N,K]^T on the persistent stream-K tcgen05 kernel (csrc/kernels/gemm_streamk.cu): every SM streams an equal share of
    the weights, split tiles are reduced in-kernel (deterministic), TMA-store epilogue.  ``silu_split``: w = [gate; up] -> silu(g) * u."""
    n_out = w.shape[0] // 2 if silu_split else w.shape[0]
    if out is None:
        out = torch.empty(x.shape[0], n_out, dtype=x.dtype, device=x.device)
    torch.ops.sglb.gemm_streamk(x, w, out, bias, EPI_SILU_SPLIT if silu_split else (EPI_BIAS if bias is not None else EPI_NONE))
    return out


def gemm_skinny_slabs(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None, max_splits: int = 8) -> torch.Tensor:
    """split-K partials [splits, T, N] (bf16); the consumer sums the slabs (``slab_reduce_add_rmsnorm``, ``rope_kv_write``)"""
    T, K = x.shape
    N = w.shape[0]
    s = int(torch.ops.sglb.gemm_skinny_splits(T, N, K, max_splits))
    slabs = torch.empty(s, T, N, dtype=x.dtype, device=x.device)
    if s == 1:
        torch.ops.sglb.gemm_skinny(x, w, slabs[0], bias, EPI_BIAS if bias is not None else EPI_NONE, 1)
    else:
        torch.ops.sglb.gemm_skinny(x, w, slabs, bias, EPI_BIAS if bias is not None else EPI_NONE, s)
    return slabs


def slab_reduce_add_rmsnorm(slabs: torch.Tensor, residual: Optional[torch.Tensor], weight: torch.Tensor, eps: float,
                            gemma: bool = False) -> torch.Tensor:
    """out = rmsnorm(residual += sum_s slabs[s]) * w    (residual updated in place when given)"""
    out = torch.empty(slabs.shape[1], slabs.shape[2], dtype=slabs.dtype, device=slabs.device)
    torch.ops.sglb.slab_reduce_add_rmsnorm(slabs, residual, weight, out, eps, gemma)
    return out


# ------------------------------------------------------------------ misc index ops
def apply_pt_delta(page_table: torch.Tensor, rows: torch.Tensor, cols: torch.Tensor, vals: torch.Tensor, n: Optional[torch.Tensor] = None,
                   n_max: Optional[int] = None) -> None:
    n_max = rows.numel() if n_max is None else n_max
    if n_max == 0:
        return
    if _use_native(page_table):
        torch.ops.sglb.apply_pt_delta(page_table, rows, cols, vals, n, n_max)
        return
    page_table[rows.long(), cols.long()] = vals


def embedding(weight: torch.Tensor, ids: torch.Tensor, gather_idx: Optional[torch.Tensor] = None, vocab_start: int = 0) -> torch.Tensor:
    T = ids.shape[0] if gather_idx is None else gather_idx.shape[0]
    if _use_native(weight) and weight.dtype == torch.bfloat16:
        out = torch.empty(T, weight.shape[1], dtype=weight.dtype, device=weight.device)
        torch.ops.sglb.embedding(weight, ids, gather_idx, out, vocab_start)
        return out
    if gather_idx is not None:
        ids = ids[gather_idx.long()]
    idl = ids.long() - vocab_start
    ok = (idl >= 0) & (idl < weight.shape[0])
    out = torch.nn.functional.embedding(idl.clamp(0, weight.shape[0] - 1), weight)
    return out * ok[:, None].to(out.dtype)


def gather_rows(x: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    if _use_native(x) and x.dtype == torch.bfloat16:
        out = torch.empty(idx.shape[0], x.shape[1], dtype=x.dtype, device=x.device)
        torch.ops.sglb.gather_rows(x, idx, out)
        return out
    return x[idx.long()]


# ------------------------------------------------------------------ sampling
def sample(logits, temperature, top_k, top_p, min_p, seeds=None, out_slots=None, next_token_buf=None, out_logprobs=None,
           global_seed: int = 0, step_ptr=None, step: int = 0, generator=None) -> torch.Tensor:
    B = logits.shape[0]
    if _use_native(logits):
        out = torch.empty(B, dtype=torch.int32, device=logits.device)
        torch.ops.sglb.sample(logits, temperature, top_k, top_p, min_p, seeds, out_slots, next_token_buf, out, out_logprobs,
                              int(global_seed) & 0x7FFFFFFFFFFFFFFF, step_ptr, int(step))
        return out
    g = generator
    if g is None:
        g = torch.Generator(device="cpu")
        g.manual_seed((int(global_seed) * 1000003 + int(step)) & 0x7FFFFFFF)
    u = torch.rand(B, generator=g)
    if seeds is not None:      # seeded rows: one uniform from (seed, own output position) — same contract as the kernel
        for j, (sd, pos) in enumerate(seeds.tolist()):
            if pos >= 0:
                gj = torch.Generator(device="cpu")
                gj.manual_seed(((int(sd) * 0x9E3779B1) ^ (int(pos) * 0x85EBCA6B + 0x5bd1e995)) & 0x7FFFFFFFFFFFFFFF)
                u[j] = torch.rand(1, generator=gj)[0]
    u = u.to(logits.device)
    tok = torch_ref.sample(logits, temperature, top_k, top_p, min_p, u).to(torch.int32)
    if out_logprobs is not None:
        lp = torch.log_softmax(logits.float() / temperature.clamp(min=1e-6)[:, None], dim=-1)
        out_logprobs.copy_(lp.gather(1, tok.long()[:, None]).squeeze(1))
    if next_token_buf is not None and out_slots is not None:
        next_token_buf[out_slots.long()] = tok
    return tok


def causal_conv1d(x: torch.Tensor, w: torch.Tensor, bias, conv_state: torch.Tensor, cu_q, req_slots: torch.Tensor, seq_lens: torch.Tensor,
                  max_q_len: int = 1, silu: bool = True) -> torch.Tensor:
    """Mamba causal depthwise conv1d over a ragged batch, rolling state updated in place (csrc/kernels/mamba.cu); fp32 [T, C] out"""
    if _use_native(x) and x.dtype in (torch.bfloat16, torch.float32) and x.stride(1) == 1 and 2 <= w.shape[1] <= 8:
        u = torch.empty(x.shape[0], x.shape[1], dtype=torch.float32, device=x.device)
        torch.ops.sglb.causal_conv1d(x, w.contiguous(), bias, conv_state, cu_q, req_slots, seq_lens, int(max_q_len), bool(silu), u)
        return u
    return torch_ref.causal_conv1d(x, w, bias, conv_state, cu_q, req_slots, seq_lens, silu)


def selective_state_update(state: torch.Tensor, req_slots: torch.Tensor, x: torch.Tensor, dt: torch.Tensor, A: torch.Tensor, Bm: torch.Tensor,
                           Cm: torch.Tensor, D=None) -> torch.Tensor:
    """Mamba-2 single-token state update + read-out, ONE pass over the fp32 state (csrc/kernels/mamba.cu); y fp32 [B, H, P]"""
    if (_use_native(state) and state.dtype == torch.float32 and state.shape[3] % 4 == 0 and all(t.dtype == torch.float32 for t in (x, dt, A, Bm, Cm))
            and Bm.stride(0) % 4 == 0 and Bm.stride(0) == Cm.stride(0) and Bm.data_ptr() % 16 == 0 and Cm.data_ptr() % 16 == 0 and x.stride(1) == 1):
        B, (H, P) = req_slots.numel(), state.shape[1:3]
        y = torch.empty(B, H, P, dtype=torch.float32, device=state.device)
        torch.ops.sglb.selective_state_update(state, req_slots, x, dt.contiguous(), A.contiguous(), Bm, Cm, D, y)
        return y
    return torch_ref.selective_state_update(state, req_slots, x, dt, A, Bm, Cm, D)


def gated_delta_update(state: torch.Tensor, req_slots: torch.Tensor, u: torch.Tensor, g: torch.Tensor, beta: torch.Tensor, Hk: int) -> Optional[torch.Tensor]:
    """gated delta rule, one token per request: u fp32 [B, 2*Hk*Dk + Hv*Dv] = conv output rows [q | k | v] (q, k un-normalised), state fp32
    [slots, Hv, Dk, Dv] updated in place -> o fp32 [B, Hv, Dv]; None when the native kernel does not cover the shape (caller falls back)"""
    if (_use_native(state) and state.dtype == torch.float32 and u.dtype == torch.float32 and u.stride(1) == 1 and state.shape[2] in (64, 128)
            and state.shape[3] <= (256 if state.shape[2] == 64 else 128)):
        B, Hv, Dv = req_slots.numel(), state.shape[1], state.shape[3]
        out = torch.empty(B, Hv, Dv, dtype=torch.float32, device=state.device)
        torch.ops.sglb.gated_delta_update(state, req_slots, u, g.float().contiguous(), beta.float().contiguous(), int(Hk), out)
        return out
    return None


def greedy_shard(logits: torch.Tensor, v_valid: int, v_offset: int, rank: int, tp: int, buf: torch.Tensor) -> None:
    """argmax of this rank's vocabulary shard -> buf[rank, row] = (max logit, idx byte 0, 1, 2) as bf16; the other ranks' planes are zeroed so
    that an all-reduce (sum) of ``buf`` is an exact all-gather (sampling.cu greedy_shard_kernel)"""
    B = logits.shape[0]
    if _use_native(logits):
        torch.ops.sglb.greedy_shard(logits, int(v_valid), int(v_offset), int(rank), int(tp), buf)
        return
    buf[:, :B].zero_()
    if v_valid > 0:
        val, idx = logits[:, :v_valid].float().max(dim=1)
        g = idx + int(v_offset)
    else:
        val = torch.full((B,), -3.3e38, dtype=torch.float32, device=logits.device)
        g = torch.zeros(B, dtype=torch.long, device=logits.device)
    buf[rank, :B, 0] = val.clamp(min=-3.3e38).to(buf.dtype)
    buf[rank, :B, 1] = (g & 255).to(buf.dtype)
    buf[rank, :B, 2] = ((g >> 8) & 255).to(buf.dtype)
    buf[rank, :B, 3] = ((g >> 16) & 255).to(buf.dtype)


def greedy_merge(buf: torch.Tensor, B: int, out_slots=None, next_token_buf=None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """winner over the gathered per-rank (max, index) pairs; equal maxima keep the lowest global index"""
    if out is None:
        out = torch.empty(B, dtype=torch.int32, device=buf.device)
    if _use_native(buf):
        torch.ops.sglb.greedy_merge(buf, int(B), out_slots, next_token_buf, out)
        return out
    v = buf[:, :B, 0].float()                              # [tp, B]
    r = v.argmax(dim=0)                                    # first maximum = lowest rank
    w = buf[:, :B, 1:].float().long()                      # [tp, B, 3]
    sel = w.gather(0, r.view(1, B, 1).expand(1, B, 3)).squeeze(0)
    tok = (sel[:, 0] | (sel[:, 1] << 8) | (sel[:, 2] << 16)).to(torch.int32)
    out[:B] = tok
    if next_token_buf is not None and out_slots is not None:
        next_token_buf[out_slots[:B].long()] = tok
    return out


def apply_token_bitmask(logits: torch.Tensor, bitmask: torch.Tensor, row_idx: Optional[torch.Tensor] = None) -> None:
    if _use_native(logits):
        torch.ops.sglb.apply_token_bitmask(logits, bitmask, row_idx)
        return
    if row_idx is None:
        torch_ref.apply_token_bitmask(logits, bitmask)
    else:
        sub = logits[row_idx.long()]
        torch_ref.apply_token_bitmask(sub, bitmask)
        logits[row_idx.long()] = sub


def apply_penalties(logits, counts, prompt_mask, freq, pres, rep) -> None:
    if _use_native(logits) and logits.dtype == torch.float32:
        torch.ops.sglb.apply_penalties(logits, counts, prompt_mask, freq, pres, rep)
        return
    pm = prompt_mask.bool() if prompt_mask is not None else torch.zeros_like(counts, dtype=torch.bool)
    torch_ref.apply_penalties(logits, counts, pm, freq, pres, rep)


def merge_attn_states(o1, lse1, o2, lse2):
    return torch_ref.merge_attn_states(o1, lse1, o2, lse2)
