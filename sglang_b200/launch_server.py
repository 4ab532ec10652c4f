"""``python -m sglang_b200.launch_server --model-path ... --tp 8`` (reference ``python/sglang/launch_server.py:16``)."""
import sys

from .entrypoints.http_server import launch_server
from .server_args import prepare_server_args

if __name__ == "__main__":
    args = prepare_server_args(sys.argv[1:])
    if args.encoder_only:      # EPD: a vision-encoder-only server (reference launch_server.py dispatch to the encoder server)
        from .disaggregation.encoder import launch_encoder_server

        launch_encoder_server(args.model_path, args.host, args.port, args.device, seed=args.random_seed)
    else:
        launch_server(args)
