"""Tensor-parallel linear layers + vocab-parallel embedding / LM head.

Parity: reference ``python/sglang/srt/layers/linear.py`` (``ReplicatedLinear:204``, ``ColumnParallelLinear:302``,
``MergedColumnParallelLinear:507``, ``QKVParallelLinear:936``, ``RowParallelLinear:1407`` forward ``:1596`` ->
all-reduce ``:1643-1662``) and ``vocab_parallel_embedding.py:188,591``.  Weight shards are ``[out_local, in]``
(K-major) which is exactly the B-operand layout of the tcgen05 GEMM.

Quantisation plugs in through ``quant_method`` (``layers/quantization.py``): ``create_weights`` /
``process_weights_after_loading`` / ``apply``.
"""
from __future__ import annotations

from typing import List, Optional

import torch
from torch import nn

from .sp import get_sp_ctx
from .. import ops
from ..parallel.parallel_state import get_tp_group
from .quantization import QuantMethod, UnquantizedMethod


class PendingAllReduce:
    """Handle for partial sums that sit in this call's NVL arena slot on every rank (not yet reduced)."""
    __slots__ = ("nvl", "off", "T", "H", "mode", "splits")

    def __init__(self, nvl, off: int, T: int, H: int, mode: str = "pull", splits: int = 1):
        self.nvl, self.off, self.T, self.H, self.mode, self.splits = nvl, off, T, H, mode, splits

    def consume_norm(self, residual, weight: torch.Tensor, eps: float, gemma: bool) -> torch.Tensor:
        """cross-rank sum + residual add + RMSNorm in one kernel; returns the normalised activations"""
        if self.mode == "rs":       # large T: reduce-scatter + norm + all-gather, row-sharded residual; the result lives in the arena
            return self.nvl.rs_add_rmsnorm_ag(self.off, self.T, self.H, residual, weight, eps, gemma)
        out = torch.empty(self.T, self.H, dtype=weight.dtype, device=weight.device)
        if self.mode == "push":     # partials of all ranks were pushed into OUR arena by the peers' GEMM epilogues
            self.nvl.push_reduce_add_rmsnorm(self.off, self.T, self.H, residual, weight, out, eps, gemma, self.splits)
        else:                       # partials sit in each producer's arena: pull them over NVLink
            self.nvl.fused_allreduce_add_rmsnorm(self.off, residual, weight, out, eps, gemma)
        return out

    def resolve(self) -> torch.Tensor:
        """plain all-reduce of the parked partials (used when the consumer is not an RMSNorm)"""
        out = torch.empty(self.T, self.H, dtype=torch.bfloat16, device=self.nvl.buf.device)
        assert self.mode == "pull", "push / reduce-scatter partials are always consumed by an RMSNorm"
        self.nvl.allreduce_from_slot(self.off, out)
        return out


class PendingLibraryAllReduce:
    """Partial sums of a LARGE row-parallel output whose all-reduce goes to NCCL (NVLS algorithm inside the library): measured on 4 x B200
    (profiles/mgpu_tp4_r2.md) the library beats both own kernels from ~2 MB up (8 MB: NCCL + local norm 57 us, own one-shot fused 75 us, own
    reduce-scatter + norm + all-gather 77 us), while the own fused kernels win below ~1 MB.  Size-based dispatch, as the reference does
    (custom all-reduce for small messages, NCCL for large ones: distributed/parallel_state.py:648-757)."""
    __slots__ = ("y", "group", "T", "H")

    def __init__(self, y: torch.Tensor, group):
        self.y, self.group, self.T, self.H = y, group, y.shape[0], y.shape[1]

    def consume_norm(self, residual, weight: torch.Tensor, eps: float, gemma: bool) -> torch.Tensor:
        y = self.resolve()
        if residual is None:
            return ops.rmsnorm(y, weight, eps, gemma=gemma)
        ops.fused_add_rmsnorm(y, residual, weight, eps, gemma=gemma)
        return y

    def resolve(self) -> torch.Tensor:
        import torch.distributed as dist

        self.group.comm_calls += 1
        self.group.comm_bytes += self.y.numel() * self.y.element_size()
        dist.all_reduce(self.y, group=self.group.device_group)
        return self.y


class PendingSlabs:
    """split-K partial products [splits, T, H] of a LOCAL decode-shape GEMM (ops.gemm_skinny_slabs); the next RMSNorm sums
    the slabs, adds the residual and normalises in one kernel."""
    __slots__ = ("slabs", "T", "H")

    def __init__(self, slabs: torch.Tensor):
        self.slabs, self.T, self.H = slabs, slabs.shape[1], slabs.shape[2]

    def consume_norm(self, residual, weight: torch.Tensor, eps: float, gemma: bool) -> torch.Tensor:
        out = torch.empty(self.T, self.H, dtype=weight.dtype, device=weight.device)
        torch.ops.sglb.slab_reduce_add_rmsnorm(self.slabs, residual, weight, out, eps, gemma)
        return out

    def resolve(self) -> torch.Tensor:
        return self.slabs[0] if self.slabs.shape[0] == 1 else self.slabs.float().sum(0).to(self.slabs.dtype)


def _skinny_allowed(layer, x: torch.Tensor, push: bool = False) -> bool:
    """decode-shape tcgen05 GEMM policy.  ``SGLB_SKINNY=tp`` (default): only where it buys a fusion cuBLAS cannot give — the
    tensor-parallel GEMM + all-reduce producer (``push``); measured in-graph on B200, cuBLAS + our elementwise kernels is on par
    or faster for the purely local GEMMs (profiles/gemm_decode_r1.md).  ``on``: every decode GEMM; ``off``: never."""
    from ..environ import envs

    mode = envs.SGLB_SKINNY.get()
    if mode == "off" or (mode == "tp" and not push):
        return False
    return (getattr(layer, "fast_path_ok", True) and isinstance(layer.quant, UnquantizedMethod) and ops.gemm_impl() != "lib"
            and ops.skinny_ok(x, layer.weight))


def _divide(a: int, b: int) -> int:
    assert a % b == 0, f"{a} is not divisible by {b}"
    return a // b


class LinearBase(nn.Module):
    def __init__(self, in_features: int, out_features: int, bias: bool, dtype: torch.dtype,
                 quant: Optional[QuantMethod] = None, prefix: str = ""):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        self.quant = quant or UnquantizedMethod()
        self.prefix = prefix
        self.quant.create_weights(self, in_features, out_features, dtype)
        self.bias = nn.Parameter(torch.zeros(out_features, dtype=dtype), requires_grad=False) if bias else None

    def apply(self, x: torch.Tensor, bias: Optional[torch.Tensor] = None) -> torch.Tensor:
        shp = x.shape
        y = self.quant.apply(self, x.reshape(-1, shp[-1]), bias)
        return y.view(*shp[:-1], y.shape[-1])


class ReplicatedLinear(LinearBase):
    def __init__(self, in_features, out_features, bias=False, dtype=torch.bfloat16, quant=None, prefix=""):
        super().__init__(in_features, out_features, bias, dtype, quant, prefix)

    def forward(self, x):
        return self.apply(x, self.bias)

    def weight_loader(self, param: torch.Tensor, loaded: torch.Tensor, shard_id=None) -> None:
        param.data.copy_(loaded)


class ColumnParallelLinear(LinearBase):
    def __init__(self, in_features, out_features, bias=False, gather_output=False, dtype=torch.bfloat16, quant=None, prefix="",
                 tp_group=None):
        self.tp = tp_group or get_tp_group()
        self.gather_output = gather_output
        self.out_total = out_features
        super().__init__(in_features, _divide(out_features, self.tp.world_size), bias, dtype, quant, prefix)

    def forward(self, x):
        y = self.apply(x, self.bias)
        if self.gather_output and self.tp.world_size > 1:
            y = self.tp.all_gather(y, dim=-1)
        return y

    def weight_loader(self, param, loaded, shard_id=None):
        n = param.shape[0]
        param.data.copy_(loaded.narrow(0, self.tp.rank * n, n))


class MergedColumnParallelLinear(LinearBase):
    """several column-parallel matrices sharing the input (gate_up_proj); each is sharded separately"""

    def __init__(self, in_features, out_sizes: List[int], bias=False, dtype=torch.bfloat16, quant=None, prefix="", tp_group=None):
        self.tp = tp_group or get_tp_group()
        self.out_sizes = out_sizes
        self.local_sizes = [_divide(s, self.tp.world_size) for s in out_sizes]
        super().__init__(in_features, sum(self.local_sizes), bias, dtype, quant, prefix)

    def forward(self, x):
        return self.apply(x, self.bias)

    def weight_loader(self, param, loaded, shard_id: Optional[int] = None):
        if shard_id is None:  # already merged checkpoint tensor
            off = 0
            chunks = []
            for s, ls in zip(self.out_sizes, self.local_sizes):
                chunks.append(loaded.narrow(0, off + self.tp.rank * ls, ls))
                off += s
            param.data.copy_(torch.cat(chunks, 0))
            return
        ls = self.local_sizes[shard_id]
        start = sum(self.local_sizes[:shard_id])
        param.data.narrow(0, start, ls).copy_(loaded.narrow(0, self.tp.rank * ls, ls))


class QKVParallelLinear(LinearBase):
    def __init__(self, hidden_size, head_dim, num_heads, num_kv_heads, bias=False, dtype=torch.bfloat16, quant=None, prefix="",
                 tp_group=None):
        self.tp = tp_group or get_tp_group()
        tp = self.tp.world_size
        self.head_dim = head_dim
        self.num_heads = _divide(num_heads, tp)
        if num_kv_heads >= tp:
            self.num_kv_heads, self.kv_replicas = _divide(num_kv_heads, tp), 1
        else:  # replicate KV heads when tp > num_kv_heads (reference linear.py:936)
            self.num_kv_heads, self.kv_replicas = 1, _divide(tp, num_kv_heads)
        self.q_size = self.num_heads * head_dim
        self.kv_size = self.num_kv_heads * head_dim
        super().__init__(hidden_size, self.q_size + 2 * self.kv_size, bias, dtype, quant, prefix)

    def forward(self, x):
        return self.apply(x, self.bias)

    def weight_loader(self, param, loaded, shard_id: Optional[str] = None):
        r = self.tp.rank
        if shard_id == "q":
            param.data.narrow(0, 0, self.q_size).copy_(loaded.narrow(0, r * self.q_size, self.q_size))
        elif shard_id in ("k", "v"):
            off = self.q_size + (0 if shard_id == "k" else self.kv_size)
            src = (r // self.kv_replicas) * self.kv_size
            param.data.narrow(0, off, self.kv_size).copy_(loaded.narrow(0, src, self.kv_size))
        else:
            raise ValueError(shard_id)


class RowParallelLinear(LinearBase):
    """input features sharded; partial sums are all-reduced (or left partial when the caller fuses the reduction
    into the next op: ``reduce_results=False``)."""

    def __init__(self, in_features, out_features, bias=False, reduce_results=True, dtype=torch.bfloat16, quant=None, prefix="",
                 tp_group=None):
        self.tp = tp_group or get_tp_group()
        self.reduce_results = reduce_results
        self.in_total = in_features
        super().__init__(_divide(in_features, self.tp.world_size), out_features, bias, dtype, quant, prefix)

    # messages up to this size take the fused path: the GEMM writes its partial product straight into the NVL arena and the
    # *next* RMSNorm kernel performs the cross-rank sum (+ residual + norm) while pulling the peers' rows over NVLink.
    FUSED_AR_MAX_BYTES = 16 << 20      # [T, H] bf16 partials up to e.g. 2048 x 4096: every decode batch and most prefill chunks

    def forward(self, x, defer_reduce: bool = False):
        """``defer_reduce``: return a ``PendingAllReduce`` (partial sums parked in the symmetric arena) instead of
        reducing; the consumer must be ``RMSNorm.forward_pending``."""
        bias = self.bias if (self.bias is not None and self.tp.rank == 0) else None   # added once: the reduction stays a pure sum
        tp = self.tp
        if tp.world_size > 1 and self.reduce_results:
            sp = get_sp_ctx()
            if sp is not None and sp.group is tp:       # sequence parallel: the reduction ends scattered over the token rows (layers/sp.py)
                return sp.reduce_scatter(self.apply(x, bias))
        if tp.world_size == 1:
            if defer_reduce and bias is None and _skinny_allowed(self, x):
                return PendingSlabs(ops.gemm_skinny_slabs(x, self.weight))      # reduced inside the next RMSNorm kernel
            return self.apply(x, bias)
        nvl = tp.nvl
        T = x.shape[0] if x.dim() == 2 else x.numel() // x.shape[-1]
        nbytes = T * self.out_features * 2
        if (defer_reduce and self.reduce_results and nvl is not None and x.dim() == 2 and nbytes <= self.FUSED_AR_MAX_BYTES
                and isinstance(self.quant, UnquantizedMethod) and bias is None and x.dtype == torch.bfloat16
                and getattr(self, "fast_path_ok", True)):
            from ..environ import envs
            from ..parallel.nvl_comm import CH_FUSED_NORM
            from ..parallel.parallel_state import get_pp_group

            if envs.SGLB_FUSED_AR.get() == "push" and get_pp_group().world_size == 1 and x.stride(0) % 8 == 0:
                # ONE kernel: tcgen05 GEMM whose epilogue pushes each tile into every peer's arena (+ flag publication)
                if _skinny_allowed(self, x, push=True):      # decode shape: swap-AB + split-K, slabs [world*splits] summed by the consumer
                    off, splits = nvl.gemm_push_skinny(x, self.weight)
                    return PendingAllReduce(nvl, off, T, self.out_features, "push", splits)
                off = nvl.gemm_push(x, self.weight)
                return PendingAllReduce(nvl, off, T, self.out_features, "push")
            if nbytes >= nvl.library_ar_min_bytes and not nvl.wants_rs(nbytes):
                return PendingLibraryAllReduce(ops.linear(x, self.weight), tp)
            if nvl.wants_rs(nbytes):      # large T: the consumer reduce-scatters (row-sharded residual), see nvl_comm.rs_add_rmsnorm_ag
                from ..parallel.nvl_comm import CH_RS_AG

                off, view = nvl.stage(CH_RS_AG, (T, self.out_features))
                ops.linear(x, self.weight, None, out=view)
                return PendingAllReduce(nvl, off, T, self.out_features, "rs")
            off, view = nvl.stage(CH_FUSED_NORM, (T, self.out_features))
            ops.linear(x, self.weight, None, out=view)
            return PendingAllReduce(nvl, off, T, self.out_features, "pull")
        y = self.apply(x, bias)
        if self.reduce_results:
            y = tp.all_reduce(y)
        return y

    def weight_loader(self, param, loaded, shard_id=None):
        if param.dim() == 1:
            param.data.copy_(loaded)
            return
        k = param.shape[1]
        param.data.copy_(loaded.narrow(1, self.tp.rank * k, k))


class VocabParallelEmbedding(nn.Module):
    def __init__(self, vocab_size: int, hidden_size: int, dtype=torch.bfloat16, tp_group=None, pad_to: int = 64):
        super().__init__()
        self.tp = tp_group or get_tp_group()
        tp = self.tp.world_size
        self.vocab_size = vocab_size
        self.padded_vocab = (vocab_size + pad_to * tp - 1) // (pad_to * tp) * (pad_to * tp)
        self.local_vocab = self.padded_vocab // tp
        self.vocab_start = self.tp.rank * self.local_vocab
        self.weight = nn.Parameter(torch.zeros(self.local_vocab, hidden_size, dtype=dtype), requires_grad=False)

    def forward(self, ids: torch.Tensor, gather_idx: Optional[torch.Tensor] = None) -> torch.Tensor:
        y = ops.embedding(self.weight, ids, gather_idx, self.vocab_start)
        if self.tp.world_size > 1:
            y = self.tp.all_reduce(y)
        return y

    def weight_loader(self, param, loaded, shard_id=None):
        n = min(self.local_vocab, max(0, loaded.shape[0] - self.vocab_start))
        if n > 0:
            param.data[:n].copy_(loaded.narrow(0, self.vocab_start, n))


class ParallelLMHead(VocabParallelEmbedding):
    """vocab-sharded LM head; logits are all-gathered (reference logits_processor.py:338)"""

    def logits(self, hidden: torch.Tensor) -> torch.Tensor:
        y = ops.linear(hidden, self.weight)
        if self.tp.world_size > 1:
            y = self.tp.all_gather(y, dim=-1)
        return y[:, : self.vocab_size]


def vocab_parallel_greedy(head: VocabParallelEmbedding, hidden: torch.Tensor, buf: torch.Tensor, out_slots, next_token_buf,
                          out_tokens: torch.Tensor) -> torch.Tensor:
    """Greedy next tokens WITHOUT gathering [B, V] logits: per-rank shard argmax -> 8 bytes per (rank, row) exchanged through the arena
    all-reduce (sum with zeros = exact all-gather) -> merge.  ``buf`` is a zero-initialised bf16 [tp, Bpad, 4] scratch owned by the caller
    (static inside a captured graph).  Replaces the NCCL all-gather of logits_processor.py:877-905 on the TP decode path."""
    tp = head.tp
    B = hidden.shape[0]
    y = ops.linear(hidden, head.weight)                                        # [B, V / tp]
    v_valid = max(0, min(head.local_vocab, head.vocab_size - head.vocab_start))
    ops.greedy_shard(y, v_valid, head.vocab_start, tp.rank_in_group, tp.world_size, buf)
    if tp.world_size == 1:
        g = buf
    elif buf.is_cuda:
        g = tp.all_reduce(buf)
    else:
        g = tp.all_reduce(buf.float()).to(buf.dtype)       # Gloo (CPU tier): no bf16 reduction
    return ops.greedy_merge(g, B, out_slots, next_token_buf, out_tokens)
