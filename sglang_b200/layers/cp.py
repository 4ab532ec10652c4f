"""Prefill context parallelism (CP): the query rows of a prefill are split between ``cp_size`` attention groups, K/V are all-gathered.

Parity: reference ``python/sglang/srt/layers/cp/zigzag.py:105`` (zig-zag Q sharding over ``attn_cp_group``; KV all-gathered per layer
``:326-337,441-469``), ``dp_attention.py:1003`` (``attn_cp_all_gather_into_tensor``), ``--enable-prefill-cp`` / ``--cp-size``.

Shape of the thing here: the TP group is cut into ``cp`` attention groups of ``tp / cp`` ranks (attention weights are sharded inside a group,
replicated across groups — the same layout as DP-attention, and the MLP / MoE / embedding / LM head stay parallel over the whole TP group via
the same gather / slice context).  Every rank sees every request (SPMD scheduler), so

* a *prefill* with enough rows runs context-parallel: request ``b``'s new positions ``[P, L)`` are cut into ``2 cp`` blocks and group ``g`` takes
  blocks ``g`` and ``2 cp - 1 - g`` (zig-zag: equal causal work).  A block ``[s, e)`` is an ordinary causal *virtual request* with
  ``seq_len = e, q_len = e - s`` on the same page-table row, so the unmodified paged attention kernel serves it.  Per layer each group ropes and
  writes its own K/V rows, all-gathers K/V over the ``cp`` group and writes the other groups' rows too — afterwards every group holds the full
  KV of its heads;
* *decode* (and tiny prefills, multimodal splices, LoRA, speculative verification) simply run the plain tensor-parallel path: all groups compute
  the same attention on their replicated KV, which is exact.

Per-layer attention FLOPs and Q/O traffic of a long prefill drop by ``cp``; the cost is one K/V all-gather per layer (NVLink: bandwidth-cheap
relative to the quadratic attention it removes at long sequence lengths).
"""
from __future__ import annotations

import dataclasses
from typing import Any, List

import numpy as np
import torch


@dataclasses.dataclass
class CpMeta:
    rank: int
    size: int
    rows: torch.Tensor               # i64 [Tl] rows of the full batch this group computes (block order)
    inv_perm: torch.Tensor           # i64 [T]  position of every original row inside the group-major gathered layout
    gather_loc: torch.Tensor         # i64 [cp * mx] KV slot of every row of the padded all-gathered K/V (0 = dummy page)
    max_rows: int
    v_slots: torch.Tensor
    v_seq_lens: torch.Tensor
    v_q_lens: torch.Tensor
    v_cu_q: torch.Tensor
    v_max_q: int


def _blocks(P: int, L: int, n: int) -> List[tuple]:
    """cut positions [P, L) into n near-equal consecutive blocks"""
    ql = L - P
    base, rem = divmod(ql, n)
    out, s = [], P
    for i in range(n):
        e = s + base + (1 if i < rem else 0)
        out.append((s, e))
        s = e
    return out


def build_meta(rank: int, size: int, seq_lens: np.ndarray, q_lens: np.ndarray, req_slots: np.ndarray, out_cache_loc: torch.Tensor) -> CpMeta:
    per_rank_rows: List[List[np.ndarray]] = [[] for _ in range(size)]
    slots, lens, qls = [], [], []
    cu = 0
    for b in range(len(seq_lens)):
        L, ql = int(seq_lens[b]), int(q_lens[b])
        P = L - ql
        bl = _blocks(P, L, 2 * size)
        for g in range(size):
            for (s, e) in (bl[g], bl[2 * size - 1 - g]):
                if e <= s:
                    continue
                per_rank_rows[g].append(cu + np.arange(s - P, e - P, dtype=np.int64))
                if g == rank:
                    slots.append(int(req_slots[b]))
                    lens.append(e)
                    qls.append(e - s)
        cu += ql
    T = cu
    cat = [np.concatenate(r) if r else np.zeros(0, np.int64) for r in per_rank_rows]
    mx = max(1, max(c.size for c in cat))
    dev = out_cache_loc.device
    loc = out_cache_loc.long()
    gl = torch.zeros(size * mx, dtype=torch.int64, device=dev)
    inv = np.zeros(T, dtype=np.int64)
    off = 0
    for g, c in enumerate(cat):
        if c.size:
            gl[g * mx: g * mx + c.size] = loc[torch.from_numpy(c).to(dev)]
        inv[c] = off + np.arange(c.size)
        off += c.size
    t32 = lambda a: torch.tensor(np.asarray(a, dtype=np.int32), dtype=torch.int32, device=dev)
    return CpMeta(rank, size, torch.from_numpy(cat[rank]).to(dev), torch.from_numpy(inv).to(dev), gl, mx, t32(slots), t32(lens), t32(qls),
                  t32(np.concatenate([[0], np.cumsum(qls)])), max(qls, default=0))


def gather_write_kv(group: Any, meta: CpMeta, local_loc: torch.Tensor, kc: torch.Tensor, vc: torch.Tensor) -> None:
    """this layer's K/V rows of the local block were just rotated and written into the cache (fused RoPE + KV-write kernel): read them back
    in cache format, all-gather over the CP group and write every other group's rows into the local cache"""
    loc = local_loc.long()
    n = loc.numel()
    kv = torch.zeros((2, meta.max_rows) + tuple(kc.shape[1:]), dtype=kc.dtype if kc.dtype != torch.float8_e4m3fn else torch.uint8,
                     device=kc.device)
    raw = (lambda t: t.view(torch.uint8)) if kc.dtype == torch.float8_e4m3fn else (lambda t: t)
    kv[0, :n] = raw(kc[loc])
    kv[1, :n] = raw(vc[loc])
    g = group.all_gather(kv.unsqueeze(0), dim=0)                         # [cp, 2, mx, Hkv, D]
    raw(kc)[meta.gather_loc] = g[:, 0].reshape((-1,) + tuple(kc.shape[1:]))
    raw(vc)[meta.gather_loc] = g[:, 1].reshape((-1,) + tuple(kc.shape[1:]))
