"""ServerArgs: the single config dataclass, with auto-derived CLI flags.

Parity: reference ``python/sglang/srt/server_args.py:473-3642`` (dataclass + auto CLI),
``:3643`` resolution pipeline, ``:9970`` PortArgs.  Ours keeps the same flag names where a
user-facing flag exists in the reference (``--model-path``, ``--tp-size``, ``--mem-fraction-static``,
``--chunked-prefill-size``, ``--max-running-requests``, ``--schedule-policy``, ...), and a small
ordered resolution pipeline (``_resolve``) instead of ~90 handlers because there is exactly one
hardware target (B200, sm_100a).
"""
from __future__ import annotations

import argparse
import dataclasses
import json
import os
import tempfile
from dataclasses import dataclass
from typing import Any, Dict, List, Optional


@dataclass
class ServerArgs:
    # --- model
    model_path: str = "dummy:tiny-llama"
    tokenizer_path: Optional[str] = None
    load_format: str = "auto"  # auto|dummy|safetensors|pt|sharded_state|gguf|remote_instance
    http_frontend: str = "uvicorn"   # uvicorn | native (C++ epoll front end, csrc/runtime/http_frontend.cpp: HTTP/1.1 + SSE + WebSocket)
    model_impl: str = "auto"   # auto: native class, else the Transformers fallback | sglang: native only | transformers: always the fallback
    remote_instance_weight_loader_seed_instance_ip: Optional[str] = None
    remote_instance_weight_loader_seed_instance_service_port: Optional[int] = None
    remote_instance_weight_loader_send_weights_group_port: int = 31500
    dtype: str = "bfloat16"
    kv_cache_dtype: str = "auto"  # auto|bf16|fp8_e4m3
    quantization: Optional[str] = None  # None|fp8|fp8_block|w4a16|w8a8_int8(unsupported)
    context_length: Optional[int] = None
    trust_remote_code: bool = False
    served_model_name: Optional[str] = None
    chat_template: Optional[str] = None
    is_embedding: bool = False
    skip_tokenizer_init: bool = False
    json_model_override_args: str = "{}"
    # --- server
    host: str = "127.0.0.1"
    port: int = 30000
    enable_multi_process_frontend: bool = False   # HTTP + tokenizer | scheduler | detokenizer in separate processes (managers/multiprocess.py)
    api_key: Optional[str] = None
    admin_api_key: Optional[str] = None
    log_level: str = "info"
    log_requests: bool = False
    enable_metrics: bool = False
    enable_trace: bool = False
    otlp_traces_endpoint: str = "localhost:4317"     # an http(s):// URL selects the OTLP/HTTP JSON exporter (observability/trace.py)
    trace_level: int = 2                             # 1 request spans, 2 + queue / prefill / decode stages, 3 + one span per scheduler step
    trace_file: Optional[str] = None                 # JSON-lines span log
    decode_log_interval: int = 40
    watchdog_timeout: float = 300.0
    graceful_shutdown_timeout: float = 30.0         # SIGTERM / SIGINT: stop accepting connections, let in-flight requests finish for this long, then tear down
    soft_watchdog_timeout: Optional[float] = None
    crash_dump_folder: Optional[str] = None
    # --- memory / scheduling
    mem_fraction_static: Optional[float] = None
    max_running_requests: Optional[int] = None
    max_queued_requests: Optional[int] = None
    enable_dynamic_batch_tokenizer: bool = False      # asyncio front half: batch-encode prompts that arrive within a short window, off the event loop
    dynamic_batch_tokenizer_batch_size: int = 32
    dynamic_batch_tokenizer_batch_timeout: float = 0.002
    request_waiting_timeout: Optional[float] = None   # seconds a request may sit in the waiting queue before it is aborted (408)
    request_running_timeout: Optional[float] = None   # seconds a request may run (from its first scheduling) before it is aborted (408)
    max_total_tokens: Optional[int] = None
    max_prefill_tokens: int = 16384
    chunked_prefill_size: Optional[int] = None
    page_size: int = 32
    schedule_policy: str = "fcfs"  # lpm|fcfs|dfs-weight|lof|random
    enable_priority_scheduling: bool = False
    schedule_conservativeness: float = 1.0
    enable_mixed_chunk: bool = False
    disable_radix_cache: bool = False
    radix_eviction_policy: str = "lru"  # lru|lfu|fifo|mru|filo|priority
    radix_cache_impl: str = "auto"  # auto|cpp|python
    disable_overlap_schedule: bool = False
    stream_interval: int = 1
    enable_prefill_delayer: bool = False
    prefill_delayer_max_delay_passes: int = 30
    prefill_delayer_token_usage_low_watermark: Optional[float] = None
    prefill_delayer_queue_min_ratio: Optional[float] = None
    prefill_delayer_max_delay_ms: Optional[float] = None
    prefill_max_requests: Optional[int] = None
    min_free_slots: Optional[int] = None
    # --- parallelism
    tp_size: int = 1
    pp_size: int = 1
    dp_size: int = 1
    ep_size: int = 1
    enable_dp_attention: bool = False
    enable_sp: bool = False                    # sequence parallelism for the token-wise half of TP prefills (layers/sp.py)
    enable_two_batch_overlap: bool = False     # TBO: split every batch in two micro-batches and interleave them layer by layer on two streams
    enable_prefill_cp: bool = False
    cp_size: int = 1
    dcp_size: int = 1
    load_balance_method: str = "round_robin"
    dist_init_addr: Optional[str] = None
    nnodes: int = 1
    node_rank: int = 0
    base_gpu_id: int = 0
    nccl_port: Optional[int] = None
    comm_backend: str = "auto"  # auto|nvl|nccl
    # --- kernels / runtime
    attention_backend: str = "auto"  # auto|tc|simt|torch_native
    sampling_backend: str = "auto"  # auto|native|pytorch|token_oracle
    gemm_backend: str = "auto"  # auto|tc|lib
    disable_cuda_graph: bool = False
    cuda_graph_max_bs: Optional[int] = None
    tokenizer_worker_num: int = 1                   # > 1: that many HTTP + tokenizer worker processes (SO_REUSEPORT) in front of ONE scheduler / detokenizer pair
    enable_eplb: bool = False                       # automatic expert re-placement from the recorded load (eplb/expert_distribution.py EPLBManager)
    eplb_rebalance_num_iterations: int = 1000       # forward passes between balance checks
    eplb_min_balancedness: float = 0.8              # re-plan when mean / max rank load of the worst layer drops below this
    eplb_window: int = 64                           # forward passes the windowed recorder modes look back
    expert_distribution_recorder_mode: str = "stat"  # stat | window | per_pass
    mamba_cache_slots: int = 0                      # > 0: state-space / hybrid models keep this many recurrent-state snapshots so that prefix caching works for them
    swa_full_tokens_ratio: float = 0.0              # > 0: hybrid sliding-window models keep the sliding layers in a second pool of ratio x the full pool's tokens (0 = one pool)
    prefill_cuda_graph_max_tokens: int = 0          # > 0: capture token-bucketed prefill graphs up to this many new tokens per step (max 4096)
    cuda_graph_bs: Optional[List[int]] = None
    enable_deterministic_inference: bool = False
    random_seed: int = 0
    device: str = "auto"
    # --- speculative decoding
    speculative_algorithm: Optional[str] = None  # EAGLE|NGRAM|STANDALONE
    speculative_ngram_index: str = "trie"        # NGRAM drafts from: trie (bounded n-gram depth, chain + tree) | suffix_automaton (unbounded matches)
    enable_dwdp: bool = False                    # DP-attention + EP prefill: prefetch peers' expert weights instead of dispatching tokens
    dwdp_min_tokens: int = 1024                  # DP-wide new-token count from which a forward uses DWDP
    dllm_algorithm: Optional[str] = None         # diffusion-LLM block decoding: LowConfidence | JointThreshold | a registered plug-in
    dllm_algorithm_config: Optional[str] = None  # YAML path or inline JSON: threshold, edit_threshold, block_size, mask_id, ...
    dllm_block_size: Optional[int] = None
    dllm_mask_id: Optional[int] = None
    speculative_draft_model_path: Optional[str] = None
    speculative_num_steps: int = 3
    speculative_eagle_topk: int = 1
    speculative_num_draft_tokens: int = 4
    # --- grammar
    grammar_backend: str = "auto"  # auto|xgrammar|regex_fsm|none
    enable_jump_forward: bool = True
    reasoning_parser: Optional[str] = None
    tool_call_parser: Optional[str] = None
    # --- LoRA
    enable_lora: bool = False
    lora_paths: Optional[List[str]] = None
    max_loras_per_batch: int = 8
    max_lora_rank: int = 64
    # --- HiCache
    enable_hierarchical_cache: bool = False
    hicache_ratio: float = 2.0
    hicache_write_policy: str = "write_through"
    hicache_mem_layout: str = "layer_first"   # layer_first | page_first (host pool layout; page_first = one contiguous record per page for L3 IO)
    hicache_storage_backend: Optional[str] = None  # file | shm | memory | remote | remote://host:port | <registered plugin backend>
    hicache_storage_backend_extra_config: Optional[str] = None   # JSON passed to the backend factory
    hicache_storage_dir: Optional[str] = None
    # --- PD disaggregation
    disaggregation_mode: str = "null"  # null|prefill|decode
    disaggregation_transfer_backend: str = "nvlink"  # nvlink|fake|tcp
    disaggregation_bootstrap_port: int = 8998
    # --- multimodal
    encoder_only: bool = False            # EPD: serve only the vision tower (disaggregation/encoder.py)
    language_only: bool = False           # EPD: do not load the vision tower; images require --encoder-urls
    encoder_urls: Optional[List[str]] = None
    # --- debugging
    kv_canary: str = "none"
    enable_kv_events: bool = False
    scripted_scheduler: bool = False

    # ------------------------------------------------------------------ resolution
    def __post_init__(self) -> None:
        self._resolve()

    def _resolve(self) -> None:
        if self.tokenizer_path is None:
            self.tokenizer_path = self.model_path
        if self.served_model_name is None:
            self.served_model_name = self.model_path
        if self.device == "auto":
            try:
                import torch

                self.device = "cuda" if torch.cuda.is_available() else "cpu"
            except Exception:
                self.device = "cpu"
        if self.load_format == "auto":
            self.load_format = "dummy" if self.model_path.startswith("dummy:") else ("gguf" if self.model_path.endswith(".gguf") else "safetensors")
        if self.http_frontend not in ("uvicorn", "native"):
            raise ValueError("--http-frontend must be uvicorn or native")
        if self.enable_deterministic_inference:
            # reduction orders must not depend on what else is in the batch: no split-K / stream-K GEMM variants (their split count follows
            # the token count), one attention split per request (model_runner._pick_splits), the library all-reduce for every size (the
            # size-based switch between the one-shot kernel and NCCL changes the summation order with the batch)
            import os as _os

            _os.environ["SGLB_SKINNY"] = "off"
            _os.environ["SGLB_STREAMK_MAX_M"] = "0"
            _os.environ["SGLB_TP_LIBRARY_AR_MIN_BYTES"] = "0"
            _os.environ["SGLB_RS_AG_MIN_BYTES"] = "0"
        if self.model_impl not in ("auto", "sglang", "transformers"):
            raise ValueError("--model-impl must be auto, sglang or transformers")
        if self.model_impl != "auto":          # travels with the other config overrides to every process that reads the model config
            import json as _json

            ov = _json.loads(self.json_model_override_args or "{}")
            ov["model_impl"] = self.model_impl
            self.json_model_override_args = _json.dumps(ov)
        if self.page_size < 1 or (self.page_size & (self.page_size - 1)):
            raise ValueError("--page-size must be a power of two")
        if self.chunked_prefill_size is None:
            # B200: 180 GB HBM3e -> large chunks keep the tcgen05 prefill kernels compute-bound
            self.chunked_prefill_size = 8192 if self.device == "cuda" else 512
        if self.chunked_prefill_size is not None and self.chunked_prefill_size <= 0:
            self.chunked_prefill_size = None  # disabled
        if self.mem_fraction_static is None:
            self.mem_fraction_static = 0.88 if self.tp_size == 1 else 0.85
        if self.schedule_policy not in ("lpm", "fcfs", "dfs-weight", "lof", "random"):
            raise ValueError(f"unknown schedule policy {self.schedule_policy}")
        if self.disable_radix_cache and self.schedule_policy in ("lpm", "dfs-weight"):
            self.schedule_policy = "fcfs"
        world = self.tp_size * self.pp_size
        if self.ep_size > 1 and self.tp_size % self.ep_size != 0:
            raise ValueError("ep_size must divide tp_size (experts are sharded inside the TP group)")
        if self.enable_dp_attention:
            if self.dp_size > self.tp_size or self.tp_size % max(1, self.dp_size):
                raise ValueError("dp attention requires dp_size to divide tp_size")
            # groups take different batches every step: one synchronous forward per iteration on every rank, no graph replay
            self.disable_overlap_schedule = True
            self.disable_cuda_graph = True
        if self.cp_size > 1:
            self.enable_prefill_cp = True
        if self.enable_prefill_cp:
            if self.cp_size <= 1 or self.tp_size % self.cp_size or self.enable_dp_attention or self.pp_size > 1 or self.dcp_size > 1:
                raise ValueError("prefill CP: cp_size (> 1) must divide tp_size; not combinable with DP-attention, PP or DCP")
        if self.dcp_size > 1:
            attn_tp = self.tp_size // (self.dp_size if self.enable_dp_attention else 1)
            if attn_tp % self.dcp_size:
                raise ValueError("dcp_size must divide the attention TP size")
            if self.page_size % self.dcp_size:
                raise ValueError("page_size must be a multiple of dcp_size (each rank stores page_size / dcp_size tokens of every page)")
        if self.cuda_graph_max_bs is None:
            self.cuda_graph_max_bs = 512 if self.tp_size == 1 else 256
        if self.device != "cuda":
            self.disable_cuda_graph = True
        if self.speculative_algorithm is not None:
            self.speculative_algorithm = self.speculative_algorithm.upper()
            self.disable_overlap_schedule = True
        if self.dllm_algorithm is not None:
            if self.speculative_algorithm is not None:
                raise ValueError("--dllm-algorithm and --speculative-algorithm are mutually exclusive")
            self.disable_overlap_schedule = True
        if self.disaggregation_mode not in ("null", "prefill", "decode"):
            raise ValueError("bad --disaggregation-mode")
        if self.attention_backend == "auto":
            self.attention_backend = "tc" if self.device == "cuda" else "torch_native"
        if self.sampling_backend == "auto":
            self.sampling_backend = "native" if self.device == "cuda" else "pytorch"
        self.world_size = world

    # ------------------------------------------------------------------ CLI
    @staticmethod
    def add_cli_args(parser: argparse.ArgumentParser) -> None:
        for f in dataclasses.fields(ServerArgs):
            flag = "--" + f.name.replace("_", "-")
            aliases = {"tp_size": ["--tp"], "pp_size": ["--pp"], "dp_size": ["--dp"], "ep_size": ["--ep"],
                       "model_path": ["--model"]}.get(f.name, [])
            t = f.type
            default = f.default if f.default is not dataclasses.MISSING else None
            if t in ("bool", bool):
                parser.add_argument(flag, *aliases, action="store_true", default=default)
            elif "List[int]" in str(t):
                parser.add_argument(flag, *aliases, type=int, nargs="+", default=default)
            elif "List[str]" in str(t):
                parser.add_argument(flag, *aliases, type=str, nargs="+", default=default)
            elif "int" in str(t):
                parser.add_argument(flag, *aliases, type=int, default=default)
            elif "float" in str(t):
                parser.add_argument(flag, *aliases, type=float, default=default)
            else:
                parser.add_argument(flag, *aliases, type=str, default=default)
        parser.add_argument("--config", type=str, default=None, help="YAML/JSON file merged under CLI flags")

    @classmethod
    def from_cli_args(cls, ns: argparse.Namespace) -> "ServerArgs":
        names = {f.name for f in dataclasses.fields(cls)}
        kw = {k: v for k, v in vars(ns).items() if k in names}
        cfg = getattr(ns, "config", None)
        if cfg:
            # the file provides defaults; a CLI flag wins only when it was actually given (differs from the built-in default)
            defaults = {f.name: f.default for f in dataclasses.fields(cls)}
            explicit = {k: v for k, v in kw.items() if v is not None and v != defaults.get(k)}
            kw = {**{k: v for k, v in kw.items() if v is not None}, **_load_config_file(cfg), **explicit}
        return cls(**kw)

    def to_dict(self) -> Dict[str, Any]:
        return dataclasses.asdict(self)

    def url(self) -> str:
        return f"http://{self.host}:{self.port}"


def _load_config_file(path: str) -> Dict[str, Any]:
    with open(path) as f:
        text = f.read()
    if path.endswith((".yaml", ".yml")):
        import yaml

        d = yaml.safe_load(text) or {}
    else:
        d = json.loads(text)
    return {k.replace("-", "_"): v for k, v in d.items()}


def _merge_config_file(argv: List[str]) -> List[str]:
    """``--config file.yaml``: keys become flags placed BEFORE the command line, so explicit flags win (reference server_args_config_parser.py)"""
    if "--config" not in argv:
        return argv
    i = argv.index("--config")
    path = argv[i + 1]
    rest = argv[:i] + argv[i + 2:]
    import yaml

    with open(path) as f:
        cfg = yaml.safe_load(f) or {}
    if not isinstance(cfg, dict):
        raise ValueError(f"{path}: expected a mapping of flag names to values")
    extra: List[str] = []
    for k, v in cfg.items():
        flag = "--" + str(k).replace("_", "-")
        if isinstance(v, bool):
            if v:
                extra.append(flag)
        elif isinstance(v, (list, tuple)):
            extra += [flag] + [str(x) for x in v]
        elif v is not None:
            extra += [flag, json.dumps(v) if isinstance(v, dict) else str(v)]
    return extra + rest


def prepare_server_args(argv: List[str]) -> ServerArgs:
    parser = argparse.ArgumentParser("sglang_b200")
    ServerArgs.add_cli_args(parser)
    from .plugins import load_plugins

    load_plugins()          # out-of-tree models / grammar backends / spec algorithms register themselves before the args are resolved
    return ServerArgs.from_cli_args(parser.parse_args(_merge_config_file(list(argv))))


@dataclass
class PortArgs:
    """IPC endpoint names (reference ``server_args.py:9970-10080``)."""

    tokenizer_ipc_name: str
    scheduler_input_ipc_name: str
    detokenizer_ipc_name: str
    rpc_ipc_name: str
    metrics_ipc_name: str
    nccl_port: int

    @staticmethod
    def init_new(server_args: ServerArgs) -> "PortArgs":
        d = tempfile.mkdtemp(prefix="sglb_ipc_")
        mk = lambda n: f"ipc://{os.path.join(d, n)}"
        port = server_args.nccl_port or (server_args.port + 1007)
        return PortArgs(mk("tok"), mk("sched"), mk("detok"), mk("rpc"), mk("metrics"), port)
