"""``python -m sglang_b200.cli {serve,generate,diffusion,bench-serving,router,killall,version}`` (reference ``python/sglang/cli/main.py:12-45``)."""
from __future__ import annotations

import json
import sys


def main(argv=None) -> None:
    argv = list(sys.argv[1:] if argv is None else argv)
    cmd = argv.pop(0) if argv else "help"
    if cmd == "serve":
        from .entrypoints.http_server import main as serve

        serve(argv)
    elif cmd == "generate":
        import argparse

        from .server_args import ServerArgs

        ap = argparse.ArgumentParser("sglang_b200 generate")
        ServerArgs.add_cli_args(ap)
        ap.add_argument("--prompt", required=True)
        ap.add_argument("--max-new-tokens", type=int, default=64)
        ap.add_argument("--temperature", type=float, default=0.0)
        ns = ap.parse_args(argv)
        from .entrypoints.engine import Engine

        eng = Engine(server_args=ServerArgs.from_cli_args(ns))
        out = eng.generate(ns.prompt, {"max_new_tokens": ns.max_new_tokens, "temperature": ns.temperature})
        print(json.dumps(out, indent=1))
    elif cmd == "bench-serving":
        from .bench_serving import main as bs

        bs(argv)
    elif cmd == "router":
        from .router.router import main as rt

        rt(argv)
    elif cmd in ("diffusion", "generate-image"):                  # the reference's `sglang generate` (multimodal_gen) product
        from .multimodal_gen.pipeline import main as dm

        dm(argv)
    elif cmd == "killall":                  # reference `killall_sglang`: here by pid registry, never by command-line pattern
        import argparse

        from .utils.process_tree import killall_servers, list_servers

        ap = argparse.ArgumentParser("sglang_b200 killall")
        ap.add_argument("--port", type=int, default=None, help="only the server listening on this port")
        ap.add_argument("--grace", type=float, default=10.0, help="seconds to wait for a graceful drain before SIGKILL")
        ap.add_argument("--list", action="store_true", help="only list the registered servers")
        ns = ap.parse_args(argv)
        if ns.list:
            print(json.dumps(list_servers(), indent=1))
        else:
            print(json.dumps({"stopped": killall_servers(ns.grace, ns.port)}))
    elif cmd == "version":
        from . import __version__

        print(__version__)
    else:
        print("usage: python -m sglang_b200.cli {serve,generate,diffusion,bench-serving,router,killall,version} [args]")


if __name__ == "__main__":
    main()
