"""``python -m sglang_b200.launch_server --model-path ... --tp 8`` (reference ``python/sglang/launch_server.py:16``)."""
import sys

from .entrypoints.http_server import launch_server
from .server_args import prepare_server_args

if __name__ == "__main__":
    launch_server(prepare_server_args(sys.argv[1:]))
