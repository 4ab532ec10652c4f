"""Per-category device timing with CUDA events (wall clock on CPU): ``with timer("attention"): ...`` — accumulates without synchronising;
``summary()`` synchronises once.  Parity: reference ``python/sglang/srt/utils/device_timer.py``."""
from __future__ import annotations

import contextlib
import time
from typing import Dict, List, Tuple

import torch


class DeviceTimer:
    def __init__(self, device: str = "cuda"):
        self.cuda = device.startswith("cuda") and torch.cuda.is_available()
        self._pending: List[Tuple[str, object, object]] = []
        self._acc: Dict[str, float] = {}
        self._n: Dict[str, int] = {}

    @contextlib.contextmanager
    def __call__(self, category: str):
        if self.cuda:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            yield
            b.record()
            self._pending.append((category, a, b))
        else:
            t0 = time.perf_counter()
            yield
            self._acc[category] = self._acc.get(category, 0.0) + (time.perf_counter() - t0) * 1e3
            self._n[category] = self._n.get(category, 0) + 1

    def summary(self) -> Dict[str, Dict[str, float]]:
        if self.cuda and self._pending:
            torch.cuda.synchronize()
            for cat, a, b in self._pending:
                self._acc[cat] = self._acc.get(cat, 0.0) + a.elapsed_time(b)
                self._n[cat] = self._n.get(cat, 0) + 1
            self._pending.clear()
        return {k: {"ms": v, "calls": self._n[k], "ms_per_call": v / max(1, self._n[k])} for k, v in self._acc.items()}
