"""Typed environment-variable registry.

Parity target: reference ``python/sglang/srt/environ.py:41-120`` (EnvBool/EnvInt/...
descriptors with ``get/set/override/is_set``).  Ours is a small declarative table:
every variable the runtime reads is declared once here so ``envs.dump()`` can print
the effective configuration and tests can ``override`` values hermetically.
"""
from __future__ import annotations

import contextlib
import os
from typing import Any, Callable, Dict, Generic, Iterator, Optional, TypeVar

T = TypeVar("T")
_REGISTRY: Dict[str, "EnvVar"] = {}


class EnvVar(Generic[T]):
    def __init__(self, name: str, default: T, parse: Callable[[str], T], doc: str = ""):
        self.name, self.default, self._parse, self.doc = name, default, parse, doc
        _REGISTRY[name] = self

    def is_set(self) -> bool:
        return self.name in os.environ

    def get(self) -> T:
        raw = os.environ.get(self.name)
        if raw is None:
            return self.default
        try:
            return self._parse(raw)
        except Exception as e:  # a bad value must fail loudly, not silently default
            raise ValueError(f"invalid value {raw!r} for env {self.name}: {e}") from e

    def set(self, value: Any) -> None:
        os.environ[self.name] = str(value)

    def clear(self) -> None:
        os.environ.pop(self.name, None)

    @contextlib.contextmanager
    def override(self, value: Any) -> Iterator[None]:
        old = os.environ.get(self.name)
        os.environ[self.name] = str(value)
        try:
            yield
        finally:
            if old is None:
                os.environ.pop(self.name, None)
            else:
                os.environ[self.name] = old


def _bool(s: str) -> bool:
    s = s.strip().lower()
    if s in ("1", "true", "yes", "on"):
        return True
    if s in ("0", "false", "no", "off", ""):
        return False
    raise ValueError("expected a boolean")


def EnvBool(name: str, default: bool, doc: str = "") -> EnvVar[bool]:
    return EnvVar(name, default, _bool, doc)


def EnvInt(name: str, default: int, doc: str = "") -> EnvVar[int]:
    return EnvVar(name, default, int, doc)


def EnvFloat(name: str, default: float, doc: str = "") -> EnvVar[float]:
    return EnvVar(name, default, float, doc)


def EnvStr(name: str, default: Optional[str], doc: str = "") -> EnvVar[Optional[str]]:
    return EnvVar(name, default, str, doc)


class _Envs:
    # --- build / kernels
    SGLB_REQUIRE_NATIVE = EnvBool("SGLB_REQUIRE_NATIVE", True, "fail loudly if the CUDA extension is missing on a GPU box")
    SGLB_FORCE_TORCH_OPS = EnvBool("SGLB_FORCE_TORCH_OPS", False, "debug: route every op through the fp32 torch reference")
    SGLB_ATTN_IMPL = EnvStr("SGLB_ATTN_IMPL", "auto", "auto|tc (tcgen05)|simt (reference-grade CUDA)")
    SGLB_GEMM_IMPL = EnvStr("SGLB_GEMM_IMPL", "auto", "auto|tc (own tcgen05 GEMM)|lib (cuBLASLt)")
    SGLB_ATTN_PERSIST = EnvBool("SGLB_ATTN_PERSIST", True, "persistent-CTA attention with an atomic work queue (pipelines across requests)")
    SGLB_SKINNY = EnvStr("SGLB_SKINNY", "tp", "tp|on|off: decode-shape (<=256 tokens) swap-AB split-K tcgen05 GEMM — tp: only as the fused "
                         "GEMM+all-reduce producer; on: every decode GEMM with fused consumers")
    SGLB_STREAMK_MAX_M = EnvInt("SGLB_STREAMK_MAX_M", 0, "bf16 GEMMs with <= this many tokens run on the persistent stream-K tcgen05 kernel "
                                "(gemm_streamk.cu); 0 = never (cuBLAS).  Default 0: measured in-graph it does not beat the library yet "
                                "(profiles/gemm_streamk_r2.md)")
    SGLB_NVTX = EnvBool("SGLB_NVTX", False, "bracket every model forward with an NVTX range (mode, batch size)")
    SGLB_TP_LIBRARY_AR_MIN_BYTES = EnvInt("SGLB_TP_LIBRARY_AR_MIN_BYTES", 3 << 19, "TP > 2: row-parallel outputs of at least this many bytes (default 1.5 MB) are all-reduced by NCCL (NVLS) + a local add+norm kernel; smaller ones by the own fused one-shot kernel (measured crossover, profiles/mgpu_tp4_r2.md); 0 = never")
    SGLB_RS_AG_MIN_BYTES = EnvInt("SGLB_RS_AG_MIN_BYTES", 0, "TP > 2: row-parallel outputs of at least this many bytes use reduce-scatter + norm + all-gather with a row-sharded residual (0 bytes per rank growth with world) instead of the one-shot fused all-reduce + norm; 0 = never (default: measured no faster than the one-shot kernel at 4 ranks, profiles/mgpu_tp4_r2.md)")
    SGLB_TP_GREEDY_SHARD = EnvBool("SGLB_TP_GREEDY_SHARD", True, "TP > 1: all-greedy decode batches take a vocab-parallel argmax (8 bytes per rank and row through the arena) instead of the [B, V] logits all-gather")
    SGLB_ATTN_DECODE = EnvStr("SGLB_ATTN_DECODE", "v3", "v3 = keys-as-M decode kernel (attention_decode.cu) | v2 = query-major kernel")
    SGLB_DISABLE_CUDA_GRAPH = EnvBool("SGLB_DISABLE_CUDA_GRAPH", False)
    # --- scheduler test hooks (reference: SGLANG_TEST_RETRACT etc., environ.py:384-393)
    SGLB_TEST_RETRACT = EnvBool("SGLB_TEST_RETRACT", False, "force a decode retraction every N steps")
    SGLB_TEST_RETRACT_INTERVAL = EnvInt("SGLB_TEST_RETRACT_INTERVAL", 3)
    SGLB_TEST_STUCK_SCHEDULER = EnvFloat("SGLB_TEST_STUCK_SCHEDULER", 0.0, "sleep seconds inside a forward (watchdog tests)")
    SGLB_TEST_DISAGG_FAILURE_PROB = EnvFloat("SGLB_TEST_DISAGG_FAILURE_PROB", 0.0)
    SGLB_STRICT_MEM_CHECK = EnvBool("SGLB_STRICT_MEM_CHECK", False, "run the pool invariant checker every busy step")
    # --- KV canary (reference: S/kv_canary)
    # --- distributed
    SGLB_COMM_IMPL = EnvStr("SGLB_COMM_IMPL", "auto", "auto|nvl (own P2P/multicast kernels)|nccl")
    SGLB_FUSED_AR = EnvStr("SGLB_FUSED_AR", "pull", "pull (GEMM writes its partial into the local arena slot, the next norm kernel reduces across "
                           "ranks via peer loads / NVLS) | push (GEMM epilogue stores into every peer, norm reduces locally) | off.  Measured: "
                           "pull wins at 2 and 8 GPUs (push issues O(world) remote stores per rank; profiles/mgpu_tp8_r1.md)")
    SGLB_SP_MIN_TOKENS = EnvInt("SGLB_SP_MIN_TOKENS", 512, "sequence parallelism: prefills with fewer new tokens keep the fused all-reduce + norm path")
    SGLB_CP_MIN_TOKENS = EnvInt("SGLB_CP_MIN_TOKENS", 1024, "prefill CP: batches with fewer new tokens run the plain TP path (KV all-gather would dominate)")
    SGLB_MOE_A2A = EnvBool("SGLB_MOE_A2A", True, "DP-attention + expert parallel: all-to-all token dispatch/combine instead of gather-all-tokens")
    SGLB_SHM_BROADCAST = EnvBool("SGLB_SHM_BROADCAST", True, "single-node control-plane fan-out over a shared-memory ring (else Gloo)")
    SGLB_NVLS = EnvBool("SGLB_NVLS", True, "use NVSwitch multicast (multimem.ld_reduce / multimem.st) in the NVL collectives when available")
    SGLB_NVL_MAX_BYTES = EnvInt("SGLB_NVL_MAX_BYTES", 512 << 20, "symmetric arena bytes per rank")
    # --- observability
    SGLB_TORCH_PROFILER_DIR = EnvStr("SGLB_TORCH_PROFILER_DIR", "/tmp/sglb_profiles")

    def dump(self) -> Dict[str, Any]:
        return {k: v.get() for k, v in _REGISTRY.items()}


envs = _Envs()
