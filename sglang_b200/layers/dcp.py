"""Decode context parallelism (DCP): the ranks of a TP group that would otherwise hold *copies* of the same KV head (``tp > num_kv_heads``,
MQA / few-head GQA) split the tokens of every sequence between them instead, so the KV capacity and the per-step KV traffic of each rank
drop by ``dcp_size``.

Parity: reference ``python/sglang/srt/layers/dcp/{layout,comm}.py`` (token interleave ``layout.py``, all-gather Q -> local attention ->
LSE all-gather + output reduce-scatter ``comm.py:83,112``) and ``--dcp-size``.

Layout (why it needs no allocator change): the scheduler, radix cache and page tables work on *virtual* pages of ``page_size`` tokens; token
``p`` of a sequence belongs to DCP rank ``p % D`` and is stored at local offset ``(p % page_size) // D`` of the *same* page id on that rank,
whose physical pages hold ``page_size / D`` tokens.  Every rank therefore uses the identical page table, and

* local KV length of a sequence of length ``L`` on rank ``r``: ``floor((L - 1 - r) / D) + 1``;
* a query at position ``qp`` sees local entries ``i <= floor((qp - r) / D)``.  Queries with the same ``qp % D`` have consecutive limits, so a
  prefill is expressed as ``D`` ordinary causal *virtual requests* per sequence (rows gathered by residue class) — the unmodified paged
  attention kernel (native tcgen05 or the fp32 reference) runs on them and returns ``(o, lse)``;
* partial results are merged across ranks with the log-sum-exp weights and reduce-scattered back to head-sharded layout.

Decode needs only device arithmetic on ``seq_lens`` (graph-capturable); prefill metadata is built on the host from the scheduler's mirrors.
"""
from __future__ import annotations

import dataclasses
from typing import Any, List, Optional

import numpy as np
import torch


@dataclasses.dataclass
class DcpMeta:
    rank: int
    size: int
    local_page_size: int
    local_loc: torch.Tensor                    # i32 [T] local KV slot of every new token (slot 0 = dummy page for tokens owned elsewhere)
    # decode
    local_lens: Optional[torch.Tensor] = None        # i32 [B] clamped to >= 1 (kernel-safe) ...
    empty: Optional[torch.Tensor] = None             # bool [B] ... and which of them are really empty on this rank
    # extend: virtual requests
    v_rows: Optional[torch.Tensor] = None            # i64 [Tv] rows of the [T, ...] query tensor, grouped by virtual request
    v_slots: Optional[torch.Tensor] = None           # i32 [Bv]
    v_seq_lens: Optional[torch.Tensor] = None        # i32 [Bv]
    v_q_lens: Optional[torch.Tensor] = None          # i32 [Bv]
    v_cu_q: Optional[torch.Tensor] = None            # i32 [Bv + 1]
    v_max_q: int = 0


def local_out_loc(out_cache_loc: torch.Tensor, page_size: int, rank: int, size: int) -> torch.Tensor:
    """virtual slot (page * page_size + offset) -> this rank's slot, or 0 (dummy page) when another rank owns the token"""
    loc = out_cache_loc.long()
    page, off = loc // page_size, loc % page_size
    local = page * (page_size // size) + off // size
    return torch.where(off % size == rank, local, torch.zeros_like(local)).to(torch.int32)


def build_meta(group: Any, page_size: int, out_cache_loc: torch.Tensor, seq_lens: torch.Tensor, decode: bool,
               seq_lens_cpu: Optional[np.ndarray] = None, q_lens_cpu: Optional[np.ndarray] = None,
               req_slots_cpu: Optional[np.ndarray] = None) -> DcpMeta:
    r, D = group.rank, group.world_size
    assert page_size % D == 0, "DCP: page_size must be a multiple of dcp_size"
    meta = DcpMeta(r, D, page_size // D, local_out_loc(out_cache_loc, page_size, r, D))
    dev = out_cache_loc.device
    if decode:
        n = torch.div(seq_lens.to(torch.int32) - 1 - r, D, rounding_mode="floor") + 1
        meta.empty = n <= 0
        meta.local_lens = torch.clamp(n, min=1).to(torch.int32)
        return meta
    rows: List[np.ndarray] = []
    slots, lens, qls = [], [], []
    cu = 0
    for b in range(len(seq_lens_cpu)):
        L, ql = int(seq_lens_cpu[b]), int(q_lens_cpu[b])
        P = L - ql
        for c in range(D):
            first = P + ((c - P) % D)
            if c < r and first < D:          # limit floor((first - r) / D) < 0: that query sees nothing on this rank
                first += D
            if first >= L:
                continue
            qp = np.arange(first, L, D, dtype=np.int64)
            rows.append(cu + (qp - P))
            slots.append(int(req_slots_cpu[b]))
            lens.append((int(qp[-1]) - r) // D + 1)
            qls.append(qp.size)
        cu += ql
    t32 = lambda a: torch.tensor(np.asarray(a, dtype=np.int32), dtype=torch.int32, device=dev)
    meta.v_rows = torch.from_numpy(np.concatenate(rows) if rows else np.zeros(0, np.int64)).to(dev)
    meta.v_slots, meta.v_seq_lens, meta.v_q_lens = t32(slots), t32(lens), t32(qls)
    meta.v_cu_q = t32(np.concatenate([[0], np.cumsum(qls)]))
    meta.v_max_q = max(qls, default=0)
    return meta


def merge_partials(group: Any, o: torch.Tensor, lse: torch.Tensor) -> torch.Tensor:
    """o [T, Hg, D] (this rank's partial over its tokens, any dtype), lse [T, Hg] fp32 natural log (-inf: nothing seen here)
    -> [T, Hg / size, D] fp32: the exact softmax-weighted combination, for this rank's own heads"""
    T, Hg, Dh = o.shape
    W = group.world_size
    all_lse = group.all_gather(lse.unsqueeze(0).contiguous(), dim=0)            # [W, T, Hg]
    m = all_lse.max(dim=0).values
    m = torch.where(torch.isfinite(m), m, torch.zeros_like(m))
    tot = m + torch.log(torch.exp(all_lse - m).sum(dim=0))
    w = torch.exp(lse - tot)
    ow = torch.where((w > 0).unsqueeze(-1), o.float() * w.unsqueeze(-1), torch.zeros((), dtype=torch.float32, device=o.device))
    ow = ow.view(T, W, Hg // W, Dh).permute(1, 0, 2, 3).contiguous().view(W * T, Hg // W, Dh)
    return group.reduce_scatter(ow)
