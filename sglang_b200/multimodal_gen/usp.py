"""Unified sequence parallelism for bidirectional (diffusion) attention.

Reference behaviour: ``multimodal_gen/runtime/layers/usp.py:674`` (``ring_attn``) and ``layers/attention/layer.py:322``
(``UlyssesAttention``), sized by ``ulysses_degree`` x ``ring_degree`` (``runtime/server_args/server_args.py:262-263``).

Design here: the sequence-parallel world of size ``U*R`` is a ``R x U`` grid — Ulysses groups are contiguous (they do the
bandwidth-heavy all-to-all, so on an NVSwitch box either axis is fine, but contiguous ranks keep it on one NVLink domain
when the ring axis crosses nodes).  Every rank owns ``S/(U*R)`` rows of the sequence.

* Ulysses: all-to-all turns [S/U, H] into [S, H/U]: full sequence, a slice of heads.  Attention is then exact and local.
* Ring:    K/V blocks rotate around the ring; each step attends the local Q to the visiting block and the partial
  results are merged by log-sum-exp.  The send of step i+1 is posted before the math of step i (overlap on a side stream
  when on CUDA; with gloo the isend/irecv pair is simply asynchronous).
* USP:     Ulysses inside, ring outside.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional, Tuple

import torch
import torch.distributed as dist


@dataclass
class UspGroup:
    ulysses: Optional[dist.ProcessGroup]
    ring: Optional[dist.ProcessGroup]
    ulysses_size: int = 1
    ring_size: int = 1
    ring_rank: int = 0
    ring_ranks: Tuple[int, ...] = (0,)
    sp_rank: int = 0                       # row-block index of this rank: ring_rank * U + ulysses_rank

    @property
    def sp_size(self) -> int:
        return self.ulysses_size * self.ring_size

    @staticmethod
    def build(ulysses_degree: int, ring_degree: int, ranks=None, backend=None) -> "UspGroup":
        """Every rank of ``ranks`` must call this (new_group is collective)."""
        world = list(ranks) if ranks is not None else list(range(dist.get_world_size()))
        assert len(world) == ulysses_degree * ring_degree, "ulysses_degree * ring_degree must equal the SP world"
        me = dist.get_rank()
        ug = rg = None
        rring: Tuple[int, ...] = (me,)
        for r in range(ring_degree):                                   # contiguous Ulysses groups
            members = world[r * ulysses_degree:(r + 1) * ulysses_degree]
            g = dist.new_group(members, backend=backend) if ulysses_degree > 1 else None
            if me in members:
                ug = g
        for u in range(ulysses_degree):                                # strided ring groups
            members = world[u::ulysses_degree]
            g = dist.new_group(members, backend=backend) if ring_degree > 1 else None
            if me in members:
                rg, rring = g, tuple(members)
        idx = world.index(me)
        return UspGroup(ug, rg, ulysses_degree, ring_degree, idx // ulysses_degree, rring, idx)


def _attend(q, k, v, scale) -> Tuple[torch.Tensor, torch.Tensor]:
    """q [Sq, H, D], k/v [Sk, H, D] -> (out [Sq, H, D] fp32, lse [Sq, H] fp32).  Bidirectional."""
    s = torch.einsum("qhd,khd->hqk", q.float(), k.float()) * scale
    lse = torch.logsumexp(s, dim=-1)
    o = torch.einsum("hqk,khd->qhd", torch.exp(s - lse.unsqueeze(-1)), v.float())
    return o, lse.transpose(0, 1).contiguous()


def _merge(o1, l1, o2, l2):
    l = torch.logaddexp(l1, l2)
    return o1 * torch.exp(l1 - l).unsqueeze(-1) + o2 * torch.exp(l2 - l).unsqueeze(-1), l


def ring_attention(q, k, v, grp: UspGroup, scale: Optional[float] = None) -> torch.Tensor:
    """q/k/v: this rank's [S_local, H, D] block.  Exact bidirectional attention over the concatenation of all ring blocks."""
    scale = scale if scale is not None else 1.0 / math.sqrt(q.shape[-1])
    R = grp.ring_size
    if R == 1:
        return _attend(q, k, v, scale)[0].to(q.dtype)
    nxt, prv = grp.ring_ranks[(grp.ring_rank + 1) % R], grp.ring_ranks[(grp.ring_rank - 1) % R]
    kv = torch.stack([k, v]).contiguous()
    out = lse = None
    for step in range(R):
        reqs, incoming = [], None
        if step < R - 1:                                               # post the rotation before the math
            incoming = torch.empty_like(kv)
            ops = [dist.P2POp(dist.isend, kv, nxt, grp.ring), dist.P2POp(dist.irecv, incoming, prv, grp.ring)]
            reqs = dist.batch_isend_irecv(ops)
        o, l = _attend(q, kv[0], kv[1], scale)
        out, lse = (o, l) if out is None else _merge(out, lse, o, l)
        for r in reqs:
            r.wait()
        if incoming is not None:
            kv = incoming
    return out.to(q.dtype)


def _a2a_seq_to_head(x: torch.Tensor, group, U: int) -> torch.Tensor:
    """[S/U, H, D] -> [S, H/U, D]"""
    s, h, d = x.shape
    send = x.reshape(s, U, h // U, d).permute(1, 0, 2, 3).contiguous()          # [U, s, h/U, d]: chunk u -> rank u
    recv = torch.empty_like(send)
    dist.all_to_all_single(recv, send, group=group)
    return recv.reshape(U * s, h // U, d)                                      # rank-major == sequence-major


def _a2a_head_to_seq(x: torch.Tensor, group, U: int) -> torch.Tensor:
    """[S, H/U, D] -> [S/U, H, D]"""
    S, hu, d = x.shape
    send = x.reshape(U, S // U, hu, d).contiguous()
    recv = torch.empty_like(send)
    dist.all_to_all_single(recv, send, group=group)
    return recv.permute(1, 0, 2, 3).reshape(S // U, U * hu, d)


def ulysses_attention(q, k, v, grp: UspGroup, scale: Optional[float] = None, inner=None) -> torch.Tensor:
    U = grp.ulysses_size
    inner = inner or (lambda a, b, c: _attend(a, b, c, scale if scale is not None else 1.0 / math.sqrt(q.shape[-1]))[0].to(q.dtype))
    if U == 1:
        return inner(q, k, v)
    assert q.shape[1] % U == 0, "heads must divide the Ulysses degree"
    if dist.get_backend(grp.ulysses) == "gloo":                                 # gloo has no all_to_all_single for CPU tensors in every build
        a2a_sh, a2a_hs = _gather_seq_to_head, _gather_head_to_seq
    else:
        a2a_sh, a2a_hs = _a2a_seq_to_head, _a2a_head_to_seq
    qf, kf, vf = (a2a_sh(t, grp.ulysses, U) for t in (q, k, v))
    return a2a_hs(inner(qf, kf, vf), grp.ulysses, U)


def _gather_seq_to_head(x, group, U):
    parts = [torch.empty_like(x) for _ in range(U)]
    dist.all_gather(parts, x.contiguous(), group=group)
    r = dist.get_rank(group)
    hu = x.shape[1] // U
    return torch.cat(parts, 0)[:, r * hu:(r + 1) * hu].contiguous()


def _gather_head_to_seq(x, group, U):
    parts = [torch.empty_like(x) for _ in range(U)]
    dist.all_gather(parts, x.contiguous(), group=group)
    r = dist.get_rank(group)
    s = x.shape[0] // U
    return torch.cat([p[r * s:(r + 1) * s] for p in parts], 1).contiguous()


def usp_attention(q, k, v, grp: Optional[UspGroup], scale: Optional[float] = None) -> torch.Tensor:
    """Ulysses inside, ring outside.  q/k/v are this rank's [S/(U*R), H, D] rows."""
    if grp is None or grp.sp_size == 1:
        scale = scale if scale is not None else 1.0 / math.sqrt(q.shape[-1])
        return _attend(q, k, v, scale)[0].to(q.dtype)
    return ulysses_attention(q, k, v, grp, scale, inner=lambda a, b, c: ring_attention(a, b, c, grp, scale))
