"""Prometheus metrics (names follow the reference's ``sglang:`` family, ``observability/metrics_collector.py:237,1480``:
num_running_reqs, num_queue_reqs, token_usage, gen_throughput, cache_hit_rate, TTFT / e2e / ITL histograms...)."""
from __future__ import annotations

from typing import Any, Dict

from prometheus_client import CollectorRegistry, Counter, Gauge, Histogram, generate_latest

_TTFT_BUCKETS = (0.01, 0.025, 0.05, 0.1, 0.2, 0.4, 0.8, 1.6, 3.2, 6.4, 12.8, 25.6)
_E2E_BUCKETS = (0.1, 0.2, 0.5, 1, 2, 5, 10, 20, 40, 80, 160, 320)
_TPOT_BUCKETS = (0.002, 0.005, 0.01, 0.02, 0.04, 0.08, 0.16, 0.32, 0.64)


class MetricsCollector:
    def __init__(self, model_name: str):
        self.reg = CollectorRegistry()
        lab = {"model_name": model_name}
        self.labels = lab
        g = lambda n, d: Gauge(n, d, list(lab), registry=self.reg).labels(**lab)
        self.running = g("sglang:num_running_reqs", "running requests")
        self.queued = g("sglang:num_queue_reqs", "queued requests")
        self.token_usage = g("sglang:token_usage", "fraction of KV pages in use")
        self.gen_throughput = g("sglang:gen_throughput", "generation throughput (token/s)")
        self.cache_hit_rate = g("sglang:cache_hit_rate", "prefix cache hit rate")
        self.new_token_ratio = g("sglang:new_token_ratio", "scheduler new-token ratio")
        self.retracted = g("sglang:num_retracted_reqs", "retracted requests (cumulative)")
        self.spec_accept = g("sglang:spec_accept_length", "speculative accept length")
        c = lambda n, d: Counter(n, d, list(lab), registry=self.reg).labels(**lab)
        self.prompt_tokens = c("sglang:prompt_tokens_total", "prefill tokens")
        self.gen_tokens = c("sglang:generation_tokens_total", "generated tokens")
        self.cached_tokens = c("sglang:cached_tokens_total", "prefix-cache hit tokens")
        self.requests = c("sglang:num_requests_total", "finished requests")
        self.aborted = c("sglang:num_aborted_requests_total", "aborted requests")
        h = lambda n, d, b: Histogram(n, d, list(lab), buckets=b, registry=self.reg).labels(**lab)
        self.ttft = h("sglang:time_to_first_token_seconds", "TTFT", _TTFT_BUCKETS)
        self.e2e = h("sglang:e2e_request_latency_seconds", "end-to-end latency", _E2E_BUCKETS)
        self.tpot = h("sglang:time_per_output_token_seconds", "TPOT", _TPOT_BUCKETS)

    def log_stats(self, st: Dict[str, Any]) -> None:
        self.running.set(st["num_running_reqs"])
        self.queued.set(st["num_waiting_reqs"])
        self.token_usage.set(st["token_usage"])
        self.gen_throughput.set(st["gen_throughput"])
        self.cache_hit_rate.set(st["cache_hit_rate"])
        self.new_token_ratio.set(st["new_token_ratio"])
        self.retracted.set(st["num_retracted_reqs"])

    def observe_request(self, meta: Dict[str, Any]) -> None:
        self.requests.inc()
        self.prompt_tokens.inc(meta.get("prompt_tokens", 0))
        self.gen_tokens.inc(meta.get("completion_tokens", 0))
        self.cached_tokens.inc(meta.get("cached_tokens", 0))
        fr = meta.get("finish_reason") or {}
        if fr.get("type") == "abort":
            self.aborted.inc()
        if "ttft" in meta:
            self.ttft.observe(meta["ttft"])
        if "e2e_latency" in meta:
            self.e2e.observe(meta["e2e_latency"])
            n = meta.get("completion_tokens", 0)
            if n > 1 and "ttft" in meta:
                self.tpot.observe((meta["e2e_latency"] - meta["ttft"]) / (n - 1))
        if "spec_accept_length" in meta:
            self.spec_accept.set(meta["spec_accept_length"])

    def render(self, state: Dict[str, Any] = None) -> bytes:
        if state:
            self.log_stats(state)
        return generate_latest(self.reg)
