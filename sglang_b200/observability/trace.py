"""Request tracing (reference ``python/sglang/srt/observability/trace.py:211,313`` + ``req_time_stats.py``): per-request
stage timestamps (queued / prefill start / first token / finished) emitted as spans, plus (level 3) one span per scheduler step.
Exporters: in-memory ring (tests, /get_server_info), JSON-lines file, and **OTLP/HTTP with the protocol's JSON encoding** to any collector
(``--otlp-traces-endpoint http://host:4318``; no SDK needed: ``OtlpHttpExporter`` batches spans on a background thread and POSTs
``/v1/traces``).  Trace context (``traceparent``) propagates through ``TokenizedGenerateReqInput.trace_context``.  With ``SGLB_NVTX=1`` the
model runner brackets every forward with an NVTX range (``nsys`` / ``ncu --nvtx`` timelines)."""
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


def _otlp_value(v: Any) -> Dict[str, Any]:
    if isinstance(v, bool):
        return {"boolValue": v}
    if isinstance(v, int):
        return {"intValue": str(v)}
    if isinstance(v, float):
        return {"doubleValue": v}
    return {"stringValue": "" if v is None else str(v)}


def to_otlp(spans, service_name: str = "sglang_b200") -> Dict[str, Any]:
    """OTLP/JSON ``ExportTraceServiceRequest`` for a list of span dicts (trace ids padded to 32 hex digits, times in unix nanoseconds)"""
    out = []
    for d in spans:
        out.append({"traceId": str(d["trace_id"]).rjust(32, "0")[-32:], "spanId": str(d["span_id"]).rjust(16, "0")[-16:],
                    **({"parentSpanId": str(d["parent_id"]).rjust(16, "0")[-16:]} if d.get("parent_id") else {}),
                    "name": d["name"], "kind": 2, "startTimeUnixNano": str(int(d["start"] * 1e9)),
                    "endTimeUnixNano": str(int((d["end"] if d["end"] is not None else d["start"]) * 1e9)),
                    "attributes": [{"key": k, "value": _otlp_value(v)} for k, v in (d.get("attrs") or {}).items() if v is not None]})
    return {"resourceSpans": [{"resource": {"attributes": [{"key": "service.name", "value": {"stringValue": service_name}}]},
                               "scopeSpans": [{"scope": {"name": "sglang_b200.observability.trace"}, "spans": out}]}]}


class OtlpHttpExporter:
    """batches spans and POSTs them to ``<endpoint>/v1/traces`` (Content-Type: application/json); never blocks or fails the caller"""

    def __init__(self, endpoint: str, batch: int = 128, interval_s: float = 1.0, timeout_s: float = 5.0, service_name: str = "sglang_b200"):
        self.url = endpoint.rstrip("/") + ("" if endpoint.rstrip("/").endswith("/v1/traces") else "/v1/traces")
        self.batch, self.interval_s, self.timeout_s, self.service_name = batch, interval_s, timeout_s, service_name
        self.q: Deque[Dict[str, Any]] = collections.deque(maxlen=65536)
        self.sent = self.failed = 0
        self._wake = threading.Event()
        self._stop = False
        self._t = threading.Thread(target=self._run, daemon=True, name="sglb-otlp")
        self._t.start()

    def add(self, d: Dict[str, Any]) -> None:
        self.q.append(d)
        if len(self.q) >= self.batch:
            self._wake.set()

    def flush(self) -> None:
        import urllib.request

        while self.q:
            chunk = [self.q.popleft() for _ in range(min(self.batch, len(self.q)))]
            try:
                req = urllib.request.Request(self.url, data=json.dumps(to_otlp(chunk, self.service_name)).encode(),
                                             headers={"Content-Type": "application/json"}, method="POST")
                urllib.request.urlopen(req, timeout=self.timeout_s).read()
                self.sent += len(chunk)
            except Exception:      # noqa: BLE001 — a dead collector must not disturb serving
                self.failed += len(chunk)

    def _run(self) -> None:
        while not self._stop:
            self._wake.wait(self.interval_s)
            self._wake.clear()
            self.flush()

    def shutdown(self) -> None:
        self._stop = True
        self._wake.set()
        self._t.join(timeout=self.timeout_s + 1)
        self.flush()


class Tracer:
    def __init__(self, level: int = 1, file: Optional[str] = None, capacity: int = 4096, otlp_endpoint: Optional[str] = None):
        self.level, self.file = level, file
        self.ring: Deque[Dict[str, Any]] = collections.deque(maxlen=capacity)
        self.otlp = OtlpHttpExporter(otlp_endpoint) if otlp_endpoint else None

    def step_span(self, name: str, start: float, end: float, **attrs: Any) -> None:
        """level >= 3: one span per scheduler step (mode, batch size, tokens)"""
        if self.level >= 3:
            self.export(Span("0" * 16 + "5c4ed01e5c4ed01e", name, start=start).finish(end, **attrs))

    def set_level(self, level: int) -> None:
        self.level = level

    def export(self, span: Span) -> None:
        d = span.to_dict()
        with _LOCK:
            self.ring.append(d)
            if self.file:
                with open(self.file, "a") as f:
                    f.write(json.dumps(d) + "\n")
        if self.otlp is not None:
            self.otlp.add(d)

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


def init_tracing(level: int = 1, file: Optional[str] = None, otlp_endpoint: Optional[str] = None) -> Tracer:
    global _TRACER
    _TRACER = Tracer(level, file, otlp_endpoint=otlp_endpoint)
    return _TRACER
