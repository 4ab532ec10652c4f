"""Plugin loading: out-of-tree code registers models, grammar backends, speculative algorithms, tool-call formats or request hooks.

Parity: reference ``python/sglang/srt/plugins/hook_registry.py`` + ``load_plugins`` (entry points) and the registries it feeds.

A plugin is a module with ``register(registry)``.  Sources: the ``sglang_b200.plugins`` entry-point group of installed distributions and the
comma-separated module list in ``SGLB_PLUGINS``.  ``registry`` exposes the in-tree registration functions plus named hook lists
(``on_request_received``, ``on_request_finished``, ``on_server_start``) that the engine calls.
"""
from __future__ import annotations

import importlib
import logging
import os
from typing import Any, Callable, Dict, List

logger = logging.getLogger(__name__)


class HookRegistry:
    def __init__(self) -> None:
        self.hooks: Dict[str, List[Callable[..., Any]]] = {}
        self.loaded: List[str] = []

    def add_hook(self, name: str, fn: Callable[..., Any]) -> None:
        self.hooks.setdefault(name, []).append(fn)

    def run(self, name: str, *a: Any, **kw: Any) -> None:
        for fn in self.hooks.get(name, ()):
            try:
                fn(*a, **kw)
            except Exception:
                logger.exception("plugin hook %s failed", name)

    # ---- in-tree registries a plugin may extend
    @staticmethod
    def register_model(arch: str, cls: Any) -> None:
        from .models.registry import register_model

        register_model(arch, cls)

    @staticmethod
    def register_tool_call_format(name: str, detector_cls: Any) -> None:
        from .function_call import function_call_parser as fcp

        fcp.DETECTORS[name] = detector_cls


REGISTRY = HookRegistry()
_loaded = False


def load_plugins(force: bool = False) -> HookRegistry:
    global _loaded
    if _loaded and not force:
        return REGISTRY
    _loaded = True
    mods: List[str] = [m.strip() for m in os.environ.get("SGLB_PLUGINS", "").split(",") if m.strip()]
    try:
        from importlib.metadata import entry_points

        mods += [ep.value.split(":")[0] for ep in entry_points(group="sglang_b200.plugins")]
    except Exception:
        pass
    for m in mods:
        if m in REGISTRY.loaded:
            continue
        try:
            mod = importlib.import_module(m)
            if hasattr(mod, "register"):
                mod.register(REGISTRY)
            REGISTRY.loaded.append(m)
            logger.info("loaded plugin %s", m)
        except Exception:
            logger.exception("failed to load plugin %s", m)
    return REGISTRY
