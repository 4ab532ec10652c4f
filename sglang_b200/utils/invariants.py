"""Tensor invariants: ``expect(t, NotNaN(), InRange(0, n), name="out_cache_loc")`` — cheap declarative checks at subsystem boundaries.

Parity: reference ``python/sglang/srt/utils/invariants.py:46-229`` (``expect`` with log / crash levels).  Level comes from ``SGLB_INVARIANTS``:
``off`` (default: zero cost, no device sync), ``log`` (warn once per site), ``crash`` (raise ``InvariantError``).
"""
from __future__ import annotations

import logging
import os
from typing import Any, Optional, Set

import torch

logger = logging.getLogger(__name__)
_warned: Set[str] = set()


class InvariantError(AssertionError):
    pass


class Check:
    def describe(self) -> str:
        return type(self).__name__

    def bad(self, t: torch.Tensor) -> Optional[str]:
        raise NotImplementedError


class NotNaN(Check):
    def bad(self, t):
        return "contains NaN" if t.is_floating_point() and bool(torch.isnan(t).any()) else None


class Finite(Check):
    def bad(self, t):
        return "contains inf/NaN" if t.is_floating_point() and not bool(torch.isfinite(t).all()) else None


class InRange(Check):
    def __init__(self, lo: float, hi: float):       # lo <= x < hi
        self.lo, self.hi = lo, hi

    def bad(self, t):
        if t.numel() == 0:
            return None
        mn, mx = t.min().item(), t.max().item()
        return f"range [{mn}, {mx}] outside [{self.lo}, {self.hi})" if (mn < self.lo or mx >= self.hi) else None


class Unique(Check):
    def bad(self, t):
        return "duplicate entries" if t.numel() and torch.unique(t).numel() != t.numel() else None


class PageAligned(Check):
    def __init__(self, page_size: int):
        self.ps = page_size

    def bad(self, t):
        return f"not a multiple of the page size {self.ps}" if t.numel() and bool((t % self.ps != 0).any()) else None


class Shape(Check):
    def __init__(self, *shape: Optional[int]):
        self.shape = shape

    def bad(self, t):
        ok = t.dim() == len(self.shape) and all(s is None or s == d for s, d in zip(self.shape, t.shape))
        return None if ok else f"shape {tuple(t.shape)} != {self.shape}"


def level() -> str:
    return os.environ.get("SGLB_INVARIANTS", "off").lower()


def expect(t: Any, *checks: Check, name: str = "tensor") -> Any:
    lv = level()
    if lv == "off" or not isinstance(t, torch.Tensor):
        return t
    for c in checks:
        msg = c.bad(t)
        if msg:
            full = f"invariant violated: {name} {msg} ({c.describe()})"
            if lv == "crash":
                raise InvariantError(full)
            if full not in _warned:
                _warned.add(full)
                logger.warning(full)
    return t
