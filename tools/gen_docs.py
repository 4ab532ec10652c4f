"""Regenerate the reference pages under docs/ from the code itself (registry, ServerArgs, environment table, HTTP routes):
``python tools/gen_docs.py``.  tests/test_docs.py fails when the committed pages are stale."""
from __future__ import annotations

import dataclasses
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def supported_models() -> str:
    from sglang_b200.models import registry

    registry._ensure()
    reg = registry._REGISTRY
    from sglang_b200.configs.model_config import _HF_ARCH_MAP

    rows = ["# Supported model architectures", "",
            "Generated from `sglang_b200/models/registry.py` (`python tools/gen_docs.py`).  `architectures[0]` of a Hugging Face `config.json` selects the "
            "class; every family marked *HF* is checked against the Hugging Face implementation in `tests/test_hf_parity.py` / `tests/test_qwen2_vl.py` / "
            "`tests/test_whisper.py`.", "", "| `architectures` entry | implementation | module |", "|---|---|---|"]
    for arch in sorted(reg):
        cls = reg[arch]
        rows.append(f"| `{arch}` | `{cls.__name__}` | `{cls.__module__.replace('sglang_b200.', '')}` |")
    rows += ["", "Aliases (loaded through another class):", ""]
    for a, b in sorted(_HF_ARCH_MAP.items()):
        rows.append(f"* `{a}` → `{b}`")
    return "\n".join(rows) + "\n"


def server_arguments() -> str:
    from sglang_b200.server_args import ServerArgs

    rows = ["# Server arguments", "", "Generated from `sglang_b200/server_args.py`.  Every field is a CLI flag (`--field-name`), a keyword of "
            "`sglang_b200.Engine(...)` and a key of a `--config file.yaml`.", "", "| flag | type | default |", "|---|---|---|"]
    for f in dataclasses.fields(ServerArgs):
        d = f.default if f.default is not dataclasses.MISSING else (f.default_factory() if f.default_factory is not dataclasses.MISSING else "")
        t = getattr(f.type, "__name__", str(f.type)).replace("typing.", "")
        rows.append(f"| `--{f.name.replace('_', '-')}` | `{t}` | `{d!r}` |")
    return "\n".join(rows) + "\n"


def environment() -> str:
    from sglang_b200.environ import envs

    rows = ["# Environment variables", "", "Generated from `sglang_b200/environ.py`.", "", "| variable | default | meaning |", "|---|---|---|"]
    for name in sorted(n for n in dir(envs) if n.startswith("SGLB_")):
        v = getattr(envs, name)
        rows.append(f"| `{name}` | `{getattr(v, 'default', '')!r}` | {getattr(v, 'doc', '') or ''} |")
    extra = {"SGLB_INVARIANTS": "off / log / crash: tensor invariants (utils/invariants.py)", "SGLB_RANK_CONSENSUS": "off / log / crash: SPMD batch-decision "
             "consensus check (utils/rank_consensus.py)", "SGLB_DUMP_DIR": "directory for the tensor dumper (debug_utils/dumper.py)",
             "SGLB_PLUGINS": "comma-separated plugin modules (plugins.py)"}
    for k, v in extra.items():
        rows.append(f"| `{k}` | | {v} |")
    return "\n".join(rows) + "\n"


def http_api() -> str:
    from sglang_b200.entrypoints.engine import Engine
    from sglang_b200.entrypoints.http_server import create_app

    eng = Engine(model_path="dummy:tiny-llama", device="cpu", log_level="error")
    app = create_app(eng)
    rows = ["# HTTP API", "", "Generated from the FastAPI app (`entrypoints/http_server.py`, `entrypoints/adapters.py`, `entrypoints/openai/*`).", "",
            "| path | methods |", "|---|---|"]
    seen = {}
    for r in app.routes:
        p, m = getattr(r, "path", None), getattr(r, "methods", None)
        if p and m is None and r.__class__.__name__ == "APIWebSocketRoute":
            m = {"WEBSOCKET"}
        if p and m and not p.startswith(("/docs", "/openapi", "/redoc")):
            seen.setdefault(p, set()).update(x for x in m if x not in ("HEAD", "OPTIONS"))
    for p in sorted(seen):
        rows.append(f"| `{p}` | {', '.join(sorted(seen[p]))} |")
    eng.shutdown()
    return "\n".join(rows) + "\n"


PAGES = {"supported_models.md": supported_models, "server_arguments.md": server_arguments, "environment_variables.md": environment, "http_api.md": http_api}


def main(check: bool = False) -> int:
    os.makedirs(os.path.join(ROOT, "docs"), exist_ok=True)
    stale = []
    for name, fn in PAGES.items():
        text = fn()
        path = os.path.join(ROOT, "docs", name)
        old = open(path).read() if os.path.exists(path) else None
        if old != text:
            stale.append(name)
            if not check:
                with open(path, "w") as f:
                    f.write(text)
    print(("stale: " if check else "wrote: ") + (", ".join(stale) or "nothing"))
    return 1 if (check and stale) else 0


if __name__ == "__main__":
    sys.exit(main("--check" in sys.argv))
