"""Numerical debugging: dump named tensors (or module outputs via forward hooks) per step and rank, then diff two dump directories.

Parity: reference ``python/sglang/srt/debug_utils/{dumper,dump_comparator}.py``.

    SGLB_DUMP_DIR=/tmp/a python ...      # run A:   dumper.dump("attn_out", t)   /  dumper.hook(model, r"layers\\.\\d+\\.mlp$")
    SGLB_DUMP_DIR=/tmp/b python ...      # run B (other build / flag / TP layout)
    python -m sglang_b200.debug_utils.dumper /tmp/a /tmp/b
"""
from __future__ import annotations

import os
import re
import sys
from typing import Any, Dict, List, Optional

import torch


class Dumper:
    def __init__(self, directory: Optional[str] = None, rank: int = 0):
        self.dir = directory or os.environ.get("SGLB_DUMP_DIR")
        self.rank = rank
        self.step = 0
        self._count: Dict[str, int] = {}
        self._handles: List[Any] = []
        if self.dir:
            os.makedirs(self.dir, exist_ok=True)

    @property
    def enabled(self) -> bool:
        return bool(self.dir)

    def next_step(self) -> None:
        self.step += 1
        self._count.clear()

    def dump(self, name: str, t: Any) -> None:
        if not self.enabled:
            return
        if hasattr(t, "resolve"):
            t = t.resolve()
        if isinstance(t, (tuple, list)):
            t = t[0]
        if not isinstance(t, torch.Tensor):
            return
        k = self._count.get(name, 0)
        self._count[name] = k + 1
        safe = re.sub(r"[^A-Za-z0-9_.-]", "_", name)
        torch.save(t.detach().float().cpu(), os.path.join(self.dir, f"step{self.step:05d}.rank{self.rank}.{safe}.{k}.pt"))

    def hook(self, model: torch.nn.Module, pattern: str) -> int:
        rx = re.compile(pattern)
        n = 0
        for name, mod in model.named_modules():
            if rx.search(name):
                self._handles.append(mod.register_forward_hook(lambda m, i, o, name=name: self.dump(name, o)))
                n += 1
        return n

    def unhook(self) -> None:
        for h in self._handles:
            h.remove()
        self._handles.clear()


def compare_dumps(a: str, b: str, atol: float = 1e-3, rtol: float = 1e-2) -> List[Dict[str, Any]]:
    """rows sorted by first divergence: {name, max_abs, max_rel, shape, ok}"""
    rows = []
    for f in sorted(os.listdir(a)):
        if not f.endswith(".pt"):
            continue
        pb = os.path.join(b, f)
        if not os.path.exists(pb):
            rows.append({"name": f, "missing": True, "ok": False})
            continue
        x, y = torch.load(os.path.join(a, f)), torch.load(pb)
        if x.shape != y.shape:
            rows.append({"name": f, "shape": (tuple(x.shape), tuple(y.shape)), "ok": False})
            continue
        d = (x - y).abs()
        mx = float(d.max()) if d.numel() else 0.0
        rel = float((d / (y.abs() + 1e-6)).max()) if d.numel() else 0.0
        rows.append({"name": f, "max_abs": mx, "max_rel": rel, "shape": tuple(x.shape), "ok": bool(torch.allclose(x, y, atol=atol, rtol=rtol))})
    return rows


def main(argv: List[str]) -> int:
    rows = compare_dumps(argv[0], argv[1])
    bad = [r for r in rows if not r["ok"]]
    for r in rows:
        print(("OK   " if r["ok"] else "DIFF ") + r["name"], {k: v for k, v in r.items() if k not in ("name", "ok")})
    print(f"{len(rows) - len(bad)}/{len(rows)} tensors match; first divergence: {bad[0]['name'] if bad else None}")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
