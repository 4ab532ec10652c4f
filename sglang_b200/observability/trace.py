"""Request tracing (reference ``python/sglang/srt/observability/trace.py:211,313`` + ``req_time_stats.py``): per-request
stage timestamps (queued / prefill start / first token / finished) emitted as spans.  Exporters: in-memory ring (tests,
/get_server_info), JSON-lines file, OTLP if ``opentelemetry`` is importable (it is not in this image, so the OTLP path is
import-gated).  Trace context (``traceparent``) propagates through ``TokenizedGenerateReqInput.trace_context``."""
from __future__ import annotations

import collections
import json
import threading
import time
import uuid
from typing import Any, Deque, Dict, Optional

_LOCK = threading.Lock()


class Span:
    __slots__ = ("trace_id", "span_id", "parent_id", "name", "start", "end", "attrs")

    def __init__(self, trace_id: str, name: str, parent_id: Optional[str] = None, start: Optional[float] = None):
        self.trace_id, self.name, self.parent_id = trace_id, name, parent_id
        self.span_id = uuid.uuid4().hex[:16]
        self.start = time.time() if start is None else start
        self.end: Optional[float] = None
        self.attrs: Dict[str, Any] = {}

    def finish(self, end: Optional[float] = None, **attrs: Any) -> "Span":
        self.end = time.time() if end is None else end
        self.attrs.update(attrs)
        return self

    def to_dict(self) -> Dict[str, Any]:
        return {"trace_id": self.trace_id, "span_id": self.span_id, "parent_id": self.parent_id, "name": self.name,
                "start": self.start, "end": self.end, "attrs": self.attrs}


class Tracer:
    def __init__(self, level: int = 1, file: Optional[str] = None, capacity: int = 4096):
        self.level, self.file = level, file
        self.ring: Deque[Dict[str, Any]] = collections.deque(maxlen=capacity)

    def set_level(self, level: int) -> None:
        self.level = level

    def export(self, span: Span) -> None:
        d = span.to_dict()
        with _LOCK:
            self.ring.append(d)
            if self.file:
                with open(self.file, "a") as f:
                    f.write(json.dumps(d) + "\n")

    def trace_request(self, rid: str, meta: Dict[str, Any], created_wall: float, trace_context: Optional[Dict[str, str]] = None) -> None:
        """emit the span tree of a finished request from its meta_info timings"""
        if self.level <= 0:
            return
        tid = (trace_context or {}).get("traceparent", "").split("-")[1] if trace_context and "traceparent" in trace_context else uuid.uuid4().hex
        root = Span(tid, "request", start=created_wall)
        e2e = meta.get("e2e_latency", 0.0)
        root.finish(created_wall + e2e, rid=rid, prompt_tokens=meta.get("prompt_tokens"), completion_tokens=meta.get("completion_tokens"),
                    cached_tokens=meta.get("cached_tokens"), finish_reason=(meta.get("finish_reason") or {}).get("type"))
        self.export(root)
        if self.level >= 2:
            q = meta.get("queue_time") or 0.0
            ttft = meta.get("ttft") or 0.0
            for name, s, e in (("queue", 0.0, q), ("prefill", q, ttft), ("decode", ttft, e2e)):
                self.export(Span(tid, name, root.span_id, created_wall + s).finish(created_wall + e))


_TRACER: Optional[Tracer] = None


def get_tracer() -> Optional[Tracer]:
    return _TRACER


def init_tracing(level: int = 1, file: Optional[str] = None) -> Tracer:
    global _TRACER
    _TRACER = Tracer(level, file)
    return _TRACER
