"""Fleet router / gateway: OpenAI-compatible reverse proxy in front of N engine instances.

Parity: reference ``sgl-model-gateway`` (Rust; ``src/policies/*.rs``: random, round_robin, cache_aware (approximate
radix tree per worker), power_of_two, prefix_hash / consistent_hashing; worker registry + health checks + circuit
breaker ``src/core/circuit_breaker.rs``; retries with backoff; ``/workers`` control plane; PD routing
``routers/http/pd_router.rs``).  No Rust toolchain exists in the image: the policy hot path (prefix matching) runs in
the C++ radix tree from ``sglb_rt`` (one tree per worker over the request bytes), the proxy itself is asyncio/aiohttp.

    python -m sglang_b200.router.router --worker-urls http://h1:30000 http://h2:30000 --policy cache_aware --port 8080
"""
from __future__ import annotations

import argparse
import asyncio
import hashlib
import json
import logging
import random
import time
from dataclasses import dataclass
from typing import Any, Dict, List, Optional

import numpy as np

logger = logging.getLogger(__name__)


@dataclass
class Worker:
    url: str
    role: str = "regular"            # regular | prefill | decode
    healthy: bool = True
    inflight: int = 0
    fails: int = 0
    opened_at: float = 0.0           # circuit breaker: time the circuit opened (0 = closed)
    served: int = 0
    tree: Any = None
    index: int = 0                   # stable id used by the KV indexer

    def available(self, now: float, cooldown: float) -> bool:
        if not self.healthy:
            return False
        if self.opened_at and now - self.opened_at < cooldown:
            return False
        return True


class Policy:
    def __init__(self, name: str, cache_threshold: float = 0.5, balance_abs: int = 32, balance_rel: float = 1.5,
                 max_tree_nodes: int = 1 << 20):
        self.name = name
        self.cache_threshold, self.balance_abs, self.balance_rel = cache_threshold, balance_abs, balance_rel
        self.max_tree_nodes = max_tree_nodes
        self._rr = 0

    def _ensure_tree(self, w: Worker) -> None:
        if w.tree is None:
            from .. import sglb_rt

            w.tree = sglb_rt.RadixTree(1, 0)

    def select(self, workers: List[Worker], text: str, routing_key: Optional[str] = None) -> Worker:
        n = self.name
        if n == "random":
            return random.choice(workers)
        if n == "round_robin":
            self._rr += 1
            return workers[self._rr % len(workers)]
        if n == "power_of_two":
            a, b = random.sample(workers, 2) if len(workers) > 1 else (workers[0], workers[0])
            return a if a.inflight <= b.inflight else b
        if n in ("prefix_hash", "consistent_hashing"):
            key = routing_key or text[:256]
            h = int.from_bytes(hashlib.md5(key.encode()).digest()[:8], "little")
            ring = sorted(workers, key=lambda w: hashlib.md5(w.url.encode()).hexdigest())
            return ring[h % len(ring)]
        if n == "cache_aware":
            loads = [w.inflight for w in workers]
            lo, hi = min(loads), max(loads)
            imbalanced = (hi - lo) > self.balance_abs and hi > self.balance_rel * max(1, lo)
            data = np.frombuffer(text.encode("utf-8", errors="ignore")[:8192], dtype=np.uint8).astype(np.int32)
            best, best_len = None, -1
            if not imbalanced and data.size:
                for w in workers:
                    self._ensure_tree(w)
                    m, _ = w.tree.match_prefix(data, 0)
                    if m.size > best_len:
                        best, best_len = w, int(m.size)
            if best is None or best_len < self.cache_threshold * max(1, data.size):
                best = min(workers, key=lambda w: w.inflight)
            if data.size:
                self._ensure_tree(best)
                best.tree.insert(data, np.zeros(data.size, dtype=np.int32), 0, 0)
                if best.tree.num_nodes() > self.max_tree_nodes:
                    best.tree.evict(best.tree.evictable_pages() // 4)
            return best
        if n == "kv_aware":
            return self._select_kv_aware(workers, text)
        raise ValueError(f"unknown policy {n}")

    # ---- KV-event driven routing (reference: experimental/sgl-router + sgl-kv-indexer): exact longest cached prefix per worker
    def attach_kv_index(self, tokenizer: Any, page_size: int) -> None:
        from .. import sglb_rt

        self.kv_index = sglb_rt.KvIndexer()
        self.kv_tokenizer, self.kv_page_size = tokenizer, page_size

    def apply_kv_events(self, worker_idx: int, events: List[Any]) -> None:
        for kind, payload in events:
            if kind == "BlockStored":
                self.kv_index.stored(worker_idx, [int(h) for h in payload["hashes"]])
            elif kind == "BlockRemoved":
                self.kv_index.removed(worker_idx, [int(h) for h in payload["hashes"]])
            elif kind == "AllBlocksCleared":
                self.kv_index.clear_worker(worker_idx)

    def _select_kv_aware(self, workers: List[Worker], text_or_ids: Any) -> Worker:
        from ..mem_cache.radix_cache import chain_hashes

        loads = [w.inflight for w in workers]
        lo, hi = min(loads), max(loads)
        imbalanced = (hi - lo) > self.balance_abs and hi > self.balance_rel * max(1, lo)
        ids = text_or_ids if not isinstance(text_or_ids, str) else self.kv_tokenizer(text_or_ids)
        hashes = chain_hashes(np.asarray(ids, dtype=np.int32), self.kv_page_size, 0).tolist()
        best = None
        if hashes and not imbalanced:
            by_idx = {w.index: w for w in workers}
            m = [(depth, -by_idx[wi].inflight, wi) for wi, depth in self.kv_index.match(hashes) if wi in by_idx]
            if m:
                depth, _, wi = max(m)
                if depth >= self.cache_threshold * len(hashes):
                    best = by_idx[wi]
        return best if best is not None else min(workers, key=lambda w: w.inflight)


def make_engine_tokenizer(tokenizer_path: Optional[str]):
    """text -> token ids exactly as a worker tokenises a `/generate` text prompt (so prefix hashes line up with the workers' events)"""
    from ..configs.model_config import load_model_config
    from ..utils.tokenizer import get_tokenizer

    path = tokenizer_path or "dummy:tiny-llama"
    cfg = load_model_config(path) if path.startswith("dummy:") or __import__("os").path.isdir(path) else None
    tok = get_tokenizer(path, cfg.vocab_size if cfg else 32000, cfg.eos_token_id if cfg else 2, cfg.bos_token_id if cfg else 1, False)
    byte_tok = type(tok).__name__ == "ByteTokenizer"

    def enc(text: str) -> List[int]:
        ids = tok.encode(text)
        return ([tok.bos_token_id] + ids) if byte_tok and getattr(tok, "bos_token_id", None) is not None else ids

    return enc


class Router:
    def __init__(self, worker_urls: List[str], policy: str = "cache_aware", prefill_urls: Optional[List[str]] = None,
                 decode_urls: Optional[List[str]] = None, max_retries: int = 3, cb_threshold: int = 5, cb_cooldown: float = 10.0,
                 health_interval: float = 5.0, rate_limit: Optional[float] = None, api_key: Optional[str] = None):
        self.workers: List[Worker] = [Worker(u.rstrip("/")) for u in worker_urls]
        self.workers += [Worker(u.rstrip("/"), "prefill") for u in (prefill_urls or [])]
        self.workers += [Worker(u.rstrip("/"), "decode") for u in (decode_urls or [])]
        for i, w in enumerate(self.workers):
            w.index = i
        self.policy = Policy(policy)
        self.max_retries, self.cb_threshold, self.cb_cooldown = max_retries, cb_threshold, cb_cooldown
        self.health_interval = health_interval
        self.api_key = api_key
        self.tokens = rate_limit            # token bucket (requests / s)
        self._bucket, self._bucket_t = float(rate_limit or 0), time.monotonic()
        self.session = None
        self.stats = {"requests": 0, "retries": 0, "failed": 0, "rate_limited": 0}

    # ------------------------------------------------------------------ helpers
    def _candidates(self, role: str = "regular") -> List[Worker]:
        now = time.monotonic()
        return [w for w in self.workers if w.role == role and w.available(now, self.cb_cooldown)]

    def _record(self, w: Worker, ok: bool) -> None:
        if ok:
            w.fails, w.opened_at = 0, 0.0
            w.served += 1
        else:
            w.fails += 1
            if w.fails >= self.cb_threshold:
                w.opened_at = time.monotonic()
                logger.warning("circuit opened for %s", w.url)

    def _admit(self) -> bool:
        if not self.tokens:
            return True
        now = time.monotonic()
        self._bucket = min(self.tokens, self._bucket + (now - self._bucket_t) * self.tokens)
        self._bucket_t = now
        if self._bucket >= 1:
            self._bucket -= 1
            return True
        return False

    @staticmethod
    def _text_of(body: Dict[str, Any]) -> str:
        if "text" in body:
            t = body["text"]
            return t if isinstance(t, str) else (t[0] if t else "")
        if "prompt" in body:
            p = body["prompt"]
            return p if isinstance(p, str) else json.dumps(p)[:4096]
        if "messages" in body:
            return json.dumps(body["messages"])[:8192]
        if "input_ids" in body:
            return json.dumps(body["input_ids"])[:8192]
        return ""

    # ------------------------------------------------------------------ app
    def build_app(self):
        from aiohttp import ClientSession, ClientTimeout, web

        router = self

        async def on_start(app):
            router.session = ClientSession(timeout=ClientTimeout(total=3600))
            app["health"] = asyncio.create_task(router._health_loop())
            if router.policy.name == "kv_aware":
                app["kv"] = asyncio.create_task(router._kv_event_loop())

        async def on_stop(app):
            app["health"].cancel()
            if "kv" in app:
                app["kv"].cancel()
            await router.session.close()

        async def proxy(request: "web.Request"):
            if router.api_key and request.headers.get("Authorization", "") != f"Bearer {router.api_key}":
                return web.json_response({"error": "invalid API key"}, status=401)
            if not router._admit():
                router.stats["rate_limited"] += 1
                return web.json_response({"error": "rate limited"}, status=429)
            raw = await request.read()
            try:
                body = json.loads(raw) if raw else {}
            except Exception:
                body = {}
            text = router._text_of(body)
            router.stats["requests"] += 1
            last_err = "no healthy worker"
            for attempt in range(router.max_retries):
                cands = router._candidates("regular")
                if not cands:
                    break
                w = router.policy.select(cands, text, body.get("routing_key"))
                w.inflight += 1
                try:
                    async with router.session.request(request.method, w.url + request.path_qs, data=raw,
                                                      headers={"Content-Type": "application/json"}) as resp:
                        if resp.status >= 500:
                            raise RuntimeError(f"upstream {resp.status}")
                        out = web.StreamResponse(status=resp.status, headers={"Content-Type": resp.headers.get("Content-Type", "application/json")})
                        await out.prepare(request)
                        async for chunk in resp.content.iter_any():
                            await out.write(chunk)
                        await out.write_eof()
                        router._record(w, True)
                        return out
                except Exception as e:
                    last_err = f"{w.url}: {e}"
                    router._record(w, False)
                    router.stats["retries"] += 1
                    await asyncio.sleep(min(1.0, 0.05 * 2 ** attempt) * (0.5 + random.random()))
                finally:
                    w.inflight -= 1
            router.stats["failed"] += 1
            return web.json_response({"error": f"all retries failed: {last_err}"}, status=503)

        async def list_workers(_):
            return web.json_response([{"url": w.url, "role": w.role, "healthy": w.healthy, "inflight": w.inflight, "served": w.served,
                                       "circuit_open": bool(w.opened_at)} for w in router.workers])

        async def add_worker(request):
            b = await request.json()
            router.workers.append(Worker(b["url"].rstrip("/"), b.get("role", "regular")))
            return web.json_response({"ok": True})

        async def remove_worker(request):
            b = await request.json()
            router.workers = [w for w in router.workers if w.url != b["url"].rstrip("/")]
            return web.json_response({"ok": True})

        async def health(_):
            ok = any(w.healthy for w in router.workers)
            return web.Response(status=200 if ok else 503)

        async def stats(_):
            return web.json_response(router.stats)

        app = web.Application(client_max_size=1 << 28)
        app.on_startup.append(on_start)
        app.on_cleanup.append(on_stop)
        app.router.add_get("/health", health)
        app.router.add_get("/workers", list_workers)
        app.router.add_post("/workers", add_worker)
        app.router.add_delete("/workers", remove_worker)
        app.router.add_get("/router_stats", stats)
        app.router.add_route("*", "/{tail:.*}", proxy)
        return app

    async def _kv_event_loop(self, period: float = 0.05) -> None:
        """pull every worker's KV events (``GET /kv_events``, workers run with --enable-kv-events) into the shared index"""
        while True:
            for w in list(self.workers):
                if not w.healthy:
                    continue
                try:
                    async with self.session.get(w.url + "/kv_events", timeout=__import__("aiohttp").ClientTimeout(total=3)) as r:
                        if r.status == 200:
                            d = await r.json()
                            self.policy.kv_page_size = int(d.get("page_size", self.policy.kv_page_size))
                            self.policy.apply_kv_events(w.index, d.get("events", []))
                except Exception:
                    pass
            await asyncio.sleep(period)

    async def _health_loop(self) -> None:
        while True:
            for w in list(self.workers):
                try:
                    async with self.session.get(w.url + "/health", timeout=__import__("aiohttp").ClientTimeout(total=3)) as r:
                        w.healthy = r.status == 200
                except Exception:
                    w.healthy = False
            await asyncio.sleep(self.health_interval)


def main(argv: Optional[List[str]] = None) -> None:
    from aiohttp import web

    ap = argparse.ArgumentParser("sglang_b200.router")
    ap.add_argument("--worker-urls", nargs="*", default=[])
    ap.add_argument("--prefill-urls", nargs="*", default=[])
    ap.add_argument("--decode-urls", nargs="*", default=[])
    ap.add_argument("--policy", default="cache_aware", choices=["random", "round_robin", "cache_aware", "power_of_two", "prefix_hash",
                                                                "consistent_hashing", "kv_aware"])
    ap.add_argument("--tokenizer-path", default=None, help="kv_aware: tokenizer used to hash request prefixes (dummy:<preset> = byte tokenizer)")
    ap.add_argument("--page-size", type=int, default=32, help="kv_aware: KV page size of the workers (refreshed from /kv_events)")
    ap.add_argument("--host", default="127.0.0.1")
    ap.add_argument("--port", type=int, default=8080)
    ap.add_argument("--max-retries", type=int, default=3)
    ap.add_argument("--rate-limit", type=float, default=None)
    ap.add_argument("--api-key", default=None)
    a = ap.parse_args(argv)
    r = Router(a.worker_urls, a.policy, a.prefill_urls, a.decode_urls, a.max_retries, rate_limit=a.rate_limit, api_key=a.api_key)
    if a.policy == "kv_aware":
        r.policy.attach_kv_index(make_engine_tokenizer(a.tokenizer_path), a.page_size)
    web.run_app(r.build_app(), host=a.host, port=a.port)


if __name__ == "__main__":
    main()
