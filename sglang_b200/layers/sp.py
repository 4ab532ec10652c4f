"""Sequence parallelism for the token-wise half of a tensor-parallel decoder (``--enable-sp``).

Parity: reference ``layers/communicator.py`` (``ScatterMode``: reduce-scatter instead of all-reduce after the row-parallel GEMMs, all-gather before
the column-parallel ones, ``:740,791``) and ``layers/k3_sp_collective.py``.

Between the tensor-parallel GEMM blocks the residual stream is kept *scattered*: rank ``r`` holds rows ``[r n, (r+1) n)`` of the (padded) token
dimension.  ``o_proj`` / ``down_proj`` end in a reduce-scatter (same bytes on the wire as the all-reduce's first half), RMSNorm and the residual add
run on ``1 / tp`` of the rows, and the normalised rows are all-gathered right before the next column-parallel GEMM (the all-reduce's second half,
moved past the norm).  Wire volume is unchanged; the elementwise work and the activation footprint of a long prefill drop by ``tp``.
Decode batches keep the fused all-reduce + norm kernels (few rows: latency-bound, nothing to save).
"""
from __future__ import annotations

from typing import Any, Optional

import torch


class SpContext:
    def __init__(self) -> None:
        self.active = False
        self.group: Any = None
        self.T = 0              # real rows
        self.Tp = 0             # padded to a multiple of the group size
        self.forwards = 0

    def begin(self, group: Any, T: int) -> None:
        W = group.world_size
        self.group, self.T, self.Tp, self.active = group, T, (T + W - 1) // W * W, True
        self.forwards += 1

    def end(self) -> None:
        self.active = False

    @property
    def n(self) -> int:
        return self.Tp // self.group.world_size

    def scatter(self, x: torch.Tensor) -> torch.Tensor:
        """full rows [T, H] (identical on every rank) -> this rank's rows of the padded layout"""
        r = self.group.rank
        out = x.new_zeros((self.n,) + tuple(x.shape[1:]))
        lo, hi = r * self.n, min(self.T, (r + 1) * self.n)
        if hi > lo:
            out[: hi - lo] = x[lo:hi]
        return out

    def gather(self, x: torch.Tensor) -> torch.Tensor:
        """local rows [n, H] -> all padded rows [Tp, H]"""
        return self.group.all_gather(x.contiguous(), dim=0)

    def reduce_scatter(self, partial: torch.Tensor) -> torch.Tensor:
        """partial sums for [T or Tp, H] rows -> this rank's reduced rows [n, H]"""
        if partial.shape[0] != self.Tp:
            pad = partial.new_zeros((self.Tp,) + tuple(partial.shape[1:]))
            pad[: partial.shape[0]] = partial
            partial = pad
        return self.group.reduce_scatter(partial)


_CTX = SpContext()


def get_sp_ctx() -> Optional[SpContext]:
    return _CTX if _CTX.active else None


def sp_context() -> SpContext:
    return _CTX
