"""DP-attention: attention runs per data-parallel group on that group's own requests; MLP / MoE / vocab layers run tensor- (or
expert-) parallel over the whole TP group on the UNION of all groups' tokens.

Parity: reference ``python/sglang/srt/layers/dp_attention.py:788-866`` (gather hidden states across DP ranks before the MLP, scatter
after; padding modes ``:81``), ``layers/communicator.py`` (ScatterMode state machine), ``scheduler_components/dp_attn.py`` (idle
batches keep the ranks in lock-step).

Mechanics here: every forward starts with ``ctx.begin()``; ``gather(x, key)`` all-gathers the per-group row counts once per key and
forward (tokens / sampled rows), pads the local rows to the maximum, all-gathers over the TP group and compacts the valid rows in DP
order; ``slice(y, key)`` returns this group's rows.  Ranks inside one attention-TP group hold identical rows, so one representative per
group is kept.  A group without work runs the same layer sequence on zero rows (``ModelRunner.forward_idle``) so that the collectives
of the other groups complete.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from ..parallel.parallel_state import get_state


class DpAttnContext:
    def __init__(self) -> None:
        st = get_state()
        self.tp = st.tp
        self.dp_size, self.dp_rank = st.attn_dp_size, st.attn_dp_rank
        self.attn_tp_size = st.tp.world_size // st.attn_dp_size
        self._counts: Dict[str, List[int]] = {}
        self.mb = 0                 # two-batch overlap: which micro-batch the current layer call belongs to (separate row counts)
        self.cp_active = False      # prefill-CP reuses this context, but only inside a context-parallel prefill (layers/cp.py)

    def begin(self) -> None:
        self._counts = {}
        from .dwdp import get_dwdp_manager

        m = get_dwdp_manager()
        if m is not None:
            m.reset()

    def counts(self, key: str, n_local: int, device: torch.device) -> List[int]:
        key = f"{key}#{self.mb}"
        c = self._counts.get(key)
        if c is None:
            t = torch.tensor([n_local], dtype=torch.int64, device=device)
            allc = self.tp.all_gather(t, dim=0).tolist()
            c = [int(allc[d * self.attn_tp_size]) for d in range(self.dp_size)]
            self._counts[key] = c
        return c

    def gather(self, x: torch.Tensor, key: str) -> torch.Tensor:
        cnt = self.counts(key, x.shape[0], x.device)
        mx = max(max(cnt), 1)
        pad = x.new_zeros((mx,) + tuple(x.shape[1:]))
        pad[: x.shape[0]] = x
        g = self.tp.all_gather(pad.unsqueeze(0), dim=0)                    # [tp, mx, ...]
        parts = [g[d * self.attn_tp_size, : cnt[d]] for d in range(self.dp_size) if cnt[d] > 0]
        return torch.cat(parts, 0) if parts else x.new_zeros((0,) + tuple(x.shape[1:]))

    def slice(self, y: torch.Tensor, key: str) -> torch.Tensor:
        cnt = self._counts[f"{key}#{self.mb}"]
        off = sum(cnt[: self.dp_rank])
        return y[off: off + cnt[self.dp_rank]]


_CTX: Optional[DpAttnContext] = None


def get_dp_ctx() -> Optional[DpAttnContext]:
    """None unless DP-attention is active"""
    global _CTX
    st = get_state()
    if st.attn_dp_size <= 1:
        _CTX = None
        return None
    if _CTX is None or _CTX.tp is not st.tp:
        _CTX = DpAttnContext()
    if st.cp_mode and not _CTX.cp_active:
        return None         # outside a CP prefill every group holds every token: the plain tensor-parallel path is exact
    return _CTX


def get_cp_ctx() -> Optional[DpAttnContext]:
    """the (possibly inactive) shared context of prefill-CP mode"""
    global _CTX
    st = get_state()
    if not st.cp_mode:
        return None
    if _CTX is None or _CTX.tp is not st.tp:
        _CTX = DpAttnContext()
    return _CTX
