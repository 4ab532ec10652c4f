"""Prometheus metrics (names follow the reference's ``sglang:`` family, ``observability/metrics_collector.py:237,1480``:
num_running_reqs, num_queue_reqs, token_usage, gen_throughput, cache_hit_rate, TTFT / e2e / ITL histograms...)."""
from __future__ import annotations

from typing import Any, Dict

from prometheus_client import CollectorRegistry, Counter, Gauge, Histogram, generate_latest

_TTFT_BUCKETS = (0.01, 0.025, 0.05, 0.1, 0.2, 0.4, 0.8, 1.6, 3.2, 6.4, 12.8, 25.6)
_E2E_BUCKETS = (0.1, 0.2, 0.5, 1, 2, 5, 10, 20, 40, 80, 160, 320)
_TPOT_BUCKETS = (0.002, 0.005, 0.01, 0.02, 0.04, 0.08, 0.16, 0.32, 0.64)
_LEN_BUCKETS = (1, 8, 32, 64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384, 32768, 131072)
_HANDLED = {"num_running_reqs", "num_waiting_reqs", "token_usage", "gen_throughput", "cache_hit_rate", "new_token_ratio", "num_retracted_reqs",
            "spec_accept_length"}


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
        self.itl = h("sglang:inter_token_latency_seconds", "inter-token latency (per streamed chunk)", _TPOT_BUCKETS)
        self.queue_time = h("sglang:queue_time_seconds", "time a request waited before its first prefill chunk", _TTFT_BUCKETS)
        self.prompt_len = h("sglang:prompt_tokens_histogram", "prompt length", _LEN_BUCKETS)
        self.gen_len = h("sglang:generation_tokens_histogram", "generation length", _LEN_BUCKETS)
        self.cached_len = h("sglang:cached_tokens_histogram", "prefix-cache hit length", _LEN_BUCKETS)
        # finish reasons, one counter per kind (stop / length / abort / ...)
        self._finish = Counter("sglang:num_finished_requests_total", "finished requests by reason", list(lab) + ["reason"], registry=self.reg)
        # every numeric key of Scheduler.get_internal_state() that has no hand-written metric above becomes a gauge (KV pages, graph hits,
        # sliding-window pool, speculative acceptance, LoRA pool, HiCache tiers, NVL collectives, H2D / D2H bytes, ...)
        self._auto: Dict[str, Any] = {}
        self._doc = {
            "forward_ct": "model forward passes", "decode_ct": "decode steps", "available_pages": "free + evictable KV pages", "total_pages": "KV pages",
            "evictable_pages": "KV pages held by unlocked prefix-tree nodes", "protected_pages": "KV pages locked by running requests",
            "num_generated_tokens": "tokens generated since start", "num_prefill_tokens": "tokens prefilled since start",
            "num_cached_tokens": "prefix-cache hit tokens since start", "max_running_requests": "request slots", "max_context_len": "context limit",
            "graph_launches": "kernels replayed through CUDA graphs", "prefill_graph_hits": "prefill steps served by a captured graph",
            "h2d_bytes": "host-to-device bytes (batch metadata)", "d2h_bytes": "device-to-host bytes (sampled ids / logprobs)",
            "scheduler_iterations": "scheduler loop iterations", "scheduler_wall_s": "seconds spent inside the scheduler loop",
            "swa_total_pages": "sliding-window pool pages", "swa_available_pages": "free sliding-window pool pages",
            "swa_slid_pages": "sliding-window pages released by running requests", "swa_tombstoned_pages": "sliding-window pages reclaimed from cached prefixes",
            "spec_accept_rate": "accepted / drafted tokens", "spec_verify_ct": "speculative verify passes", "lora_slots_used": "LoRA pool slots in use",
            "lora_adapters_loaded": "registered LoRA adapters", "nvl_collective_calls": "NVLink arena collectives issued", "nvl_rs_calls": "reduce-scatter + norm + all-gather calls",
            "grammar_queue": "requests waiting for a grammar to compile", "num_sessions": "open sessions", "pending_results": "launched steps whose results are not processed yet",
        }

    def log_stats(self, st: Dict[str, Any]) -> None:
        self.running.set(st["num_running_reqs"])
        self.queued.set(st["num_waiting_reqs"])
        self.token_usage.set(st["token_usage"])
        self.gen_throughput.set(st["gen_throughput"])
        self.cache_hit_rate.set(st["cache_hit_rate"])
        self.new_token_ratio.set(st["new_token_ratio"])
        self.retracted.set(st["num_retracted_reqs"])
        if "spec_accept_length" in st:
            self.spec_accept.set(st["spec_accept_length"])
        for k, v in st.items():
            if k in _HANDLED or isinstance(v, bool) or not isinstance(v, (int, float)):
                continue
            g = self._auto.get(k)
            if g is None:
                g = self._auto[k] = Gauge(f"sglang:{k}", self._doc.get(k, k.replace("_", " ")), list(self.labels), registry=self.reg).labels(**self.labels)
            g.set(v)

    def observe_request(self, meta: Dict[str, Any]) -> None:
        self.requests.inc()
        self.prompt_tokens.inc(meta.get("prompt_tokens", 0))
        self.gen_tokens.inc(meta.get("completion_tokens", 0))
        self.cached_tokens.inc(meta.get("cached_tokens", 0))
        fr = meta.get("finish_reason") or {}
        if fr.get("type") == "abort":
            self.aborted.inc()
        self._finish.labels(reason=str(fr.get("type") or "unknown"), **self.labels).inc()
        self.prompt_len.observe(meta.get("prompt_tokens", 0))
        self.gen_len.observe(meta.get("completion_tokens", 0))
        self.cached_len.observe(meta.get("cached_tokens", 0))
        if meta.get("queue_time") is not None:
            self.queue_time.observe(meta["queue_time"])
        for dt in meta.get("itl", ()) or ():
            self.itl.observe(dt)
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
