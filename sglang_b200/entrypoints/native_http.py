"""ASGI host on the native C++ HTTP front end (``csrc/runtime/http_frontend.cpp``).

Parity: the reference embeds a Rust (axum) API server in the scheduler process (``rust/sglang-server``, enabled by ``SGLANG_RUST_SERVER``;
``python/sglang/srt/managers/rust_server.py``) so that socket handling never competes with Python for the GIL.  Here the same split is done
with C++: one epoll thread accepts, parses HTTP/1.1, keeps connections alive, answers health probes and other registered static paths by
itself, frames chunked / SSE output and notices disconnects; Python receives *complete requests* from a queue and returns response bytes.
The existing FastAPI application (every route in ``docs/http_api.md``) runs unchanged on top through the ASGI protocol, so
``--http-frontend native`` swaps uvicorn's Python socket layer for the native one and nothing else.

WebSocket routes (``/v1/realtime``) run on the same front end: handshake, masking, fragment reassembly, ping / pong and close are handled by
the I/O thread, Python sees whole messages through the ASGI ``websocket`` scope.  Not handled natively: TLS, HTTP/2, permessage-deflate.
"""
from __future__ import annotations

import asyncio
import logging
import signal
import threading
from typing import Any, Dict, List, Optional, Tuple
from urllib.parse import unquote

logger = logging.getLogger(__name__)


def _rt():
    try:
        from .. import sglb_rt as rt
    except ImportError as e:
        raise RuntimeError("the native HTTP front end needs the C++ runtime (python -m sglang_b200.build)") from e
    if not hasattr(rt, "HttpFrontend"):
        raise RuntimeError("sglb_rt was built before the native HTTP front end existed: rebuild (python -m sglang_b200.build)")
    return rt


class NativeAsgiServer(object):
    def __init__(self, app: Any, host: str = "127.0.0.1", port: int = 30000, *, reuse_port: bool = False, max_body: int = 1 << 30,
                 graceful_timeout: float = 30.0, static: Optional[Dict[str, Tuple[int, str, bytes]]] = None):
        self.app = app
        self.fe = _rt().HttpFrontend(host, port, 2048, max_body, reuse_port)
        self.host, self.port = host, self.fe.port
        self.graceful_timeout = graceful_timeout
        for path, (status, ctype, body) in (static or {}).items():
            self.fe.set_static(path, status, ctype, body)
        self._disc: Dict[int, asyncio.Event] = {}
        self._wsq: Dict[int, "asyncio.Queue"] = {}
        self._tasks: set = set()
        self._stop: Optional[asyncio.Event] = None
        self._loop: Optional[asyncio.AbstractEventLoop] = None
        self.started = threading.Event()

    # ---- life cycle
    def run(self, install_signal_handlers: bool = True) -> None:
        asyncio.run(self.serve(install_signal_handlers))

    def request_stop(self) -> None:
        if self._loop is not None and self._stop is not None:
            self._loop.call_soon_threadsafe(self._stop.set)

    async def serve(self, install_signal_handlers: bool = True) -> None:
        self._loop = asyncio.get_running_loop()
        self._stop = asyncio.Event()
        if install_signal_handlers:
            for sig in (signal.SIGINT, signal.SIGTERM):
                try:
                    self._loop.add_signal_handler(sig, self._stop.set)
                except (NotImplementedError, RuntimeError, ValueError):      # not the main thread
                    pass
        life = _Lifespan(self.app)
        await life.startup()
        self.fe.start()
        # the C++ side raises an eventfd while requests / disconnects are queued: the loop drains them itself (a helper thread would have
        # to win the GIL from the loop thread for every request, i.e. wait out the interpreter's switch interval)
        self._loop.add_reader(self.fe.notify_fd, self._on_ready)
        logger.info("native HTTP front end listening on http://%s:%d", self.host, self.port)
        self.started.set()
        try:
            await self._stop.wait()
        finally:
            # drain: the listener closes at once, in-flight requests (= running generations) may finish within the grace period
            stopper = self._loop.run_in_executor(None, self.fe.stop, float(self.graceful_timeout))
            while not stopper.done():
                await asyncio.sleep(0.02)
            self._loop.remove_reader(self.fe.notify_fd)
            for t in list(self._tasks):
                t.cancel()
            await life.shutdown()

    # ---- C++ -> asyncio
    def _on_ready(self) -> None:
        evs = self.fe.poll(0, 256)
        if evs:
            self._dispatch(evs)

    def _dispatch(self, evs: List[tuple]) -> None:
        for ev in evs:
            if ev[0] in (2, 3):          # WebSocket message / peer close
                q = self._wsq.get(ev[1])
                if q is not None:
                    q.put_nowait(ev)
                continue
            if ev[0] == 1:
                d = self._disc.get(ev[1])
                if d is not None:
                    d.set()
                q = self._wsq.get(ev[1])
                if q is not None:
                    q.put_nowait((3, ev[1], "", b"", "", "", [], b""))
                continue
            t = asyncio.ensure_future(self._handle(ev))
            self._tasks.add(t)
            t.add_done_callback(self._tasks.discard)

    async def _handle(self, ev: tuple) -> None:
        _, rid, method, target, version, peer, headers, body = ev
        fe = self.fe
        path, _, qs = target.partition(b"?")
        hdrs = [(k.lower(), v) for k, v in headers]
        ip, _, port = peer.rpartition(":")
        if any(k == b"upgrade" and v.lower() == b"websocket" for k, v in hdrs):
            await self._handle_ws(rid, path, qs, hdrs, (ip, int(port or 0)), version)
            return
        scope = {"type": "http", "asgi": {"version": "3.0", "spec_version": "2.3"}, "http_version": version.split("/")[-1], "method": method,
                 "scheme": "http", "path": unquote(path.decode("latin-1")), "raw_path": path, "query_string": qs, "root_path": "",
                 "headers": hdrs, "client": (ip, int(port or 0)), "server": (self.host, self.port)}
        disc = asyncio.Event()
        self._disc[rid] = disc
        st: Dict[str, Any] = {"got_body": False, "status": 200, "headers": [], "started": False, "streaming": False, "done": False}

        async def receive():
            if not st["got_body"]:
                st["got_body"] = True
                return {"type": "http.request", "body": body, "more_body": False}
            await disc.wait()
            return {"type": "http.disconnect"}

        async def send(msg):
            t = msg["type"]
            if t == "http.response.start":
                st["status"] = int(msg["status"])
                st["headers"] = [(k.decode("latin-1"), v.decode("latin-1")) for k, v in msg.get("headers", [])]
                st["started"] = True
            elif t == "http.response.body":
                if st["done"]:
                    return
                data, more = bytes(msg.get("body", b"")), bool(msg.get("more_body", False))
                if method == "HEAD":
                    data = b""
                if not st["streaming"]:
                    if not more:
                        ok = fe.respond(rid, st["status"], st["headers"], data)
                        st["done"] = True
                    else:
                        ok = fe.start_stream(rid, st["status"], st["headers"]) and fe.send_chunk(rid, data)
                        st["streaming"] = True
                else:
                    ok = fe.send_chunk(rid, data)
                    if not more:
                        fe.end_stream(rid)
                        st["done"] = True
                if not ok:           # the client is gone: let the application see the disconnect
                    disc.set()

        try:
            await self.app(scope, receive, send)
            if not st["done"]:
                if st["streaming"]:
                    fe.end_stream(rid)
                elif st["started"]:
                    fe.respond(rid, st["status"], st["headers"], b"")
                else:
                    fe.respond(rid, 500, [("Content-Type", "application/json")], b'{"error": "no response"}')
        except asyncio.CancelledError:
            raise
        except Exception as e:          # the application failed outside its own error handling
            logger.exception("unhandled error in %s %s", method, path)
            if st["streaming"]:
                fe.end_stream(rid)
            elif not st["done"]:
                fe.respond(rid, 500, [("Content-Type", "application/json")], ('{"error": "%s"}' % type(e).__name__).encode())
        finally:
            self._disc.pop(rid, None)


class _WsMixin:
    async def _handle_ws(self, rid: int, path: bytes, qs: bytes, hdrs, client, version: str) -> None:
        """ASGI ``websocket`` scope over the native frames: the C++ thread does the handshake hash, (un)masking, fragment reassembly, ping / pong
        and close echo; here only whole messages move"""
        fe = self.fe
        q: asyncio.Queue = asyncio.Queue()
        self._wsq[rid] = q
        sub = [v.decode("latin-1") for k, v in hdrs if k == b"sec-websocket-protocol"]
        scope = {"type": "websocket", "asgi": {"version": "3.0", "spec_version": "2.3"}, "http_version": version.split("/")[-1], "scheme": "ws",
                 "path": unquote(path.decode("latin-1")), "raw_path": path, "query_string": qs, "root_path": "", "headers": hdrs, "client": client,
                 "server": (self.host, self.port), "subprotocols": [p.strip() for s_ in sub for p in s_.split(",")]}
        st = {"connected": False, "accepted": False, "closed": False}

        async def receive():
            if not st["connected"]:
                st["connected"] = True
                return {"type": "websocket.connect"}
            ev = await q.get()
            if ev[0] == 2:
                return {"type": "websocket.receive", "text": ev[7].decode("utf-8")} if ev[2] == "text" else {"type": "websocket.receive", "bytes": ev[7]}
            st["closed"] = True
            code = int.from_bytes(ev[7][:2], "big") if len(ev[7]) >= 2 else 1005
            return {"type": "websocket.disconnect", "code": code}

        async def send(msg):
            t = msg["type"]
            if t == "websocket.accept":
                hs = [(k.decode("latin-1"), v.decode("latin-1")) for k, v in msg.get("headers", [])]
                if msg.get("subprotocol"):
                    hs.append(("Sec-WebSocket-Protocol", msg["subprotocol"]))
                st["accepted"] = fe.ws_accept(rid, hs)
            elif t == "websocket.send":
                if msg.get("text") is not None:
                    fe.ws_send(rid, False, msg["text"].encode("utf-8"))
                else:
                    fe.ws_send(rid, True, bytes(msg.get("bytes") or b""))
            elif t == "websocket.close":
                if not st["accepted"]:
                    fe.respond(rid, 403, [("Content-Type", "application/json")], b'{"error": "websocket rejected"}')
                elif not st["closed"]:
                    fe.ws_close(rid, int(msg.get("code", 1000)))
                st["closed"] = True

        try:
            await self.app(scope, receive, send)
            if st["accepted"] and not st["closed"]:
                fe.ws_close(rid, 1000)
            elif not st["accepted"] and not st["closed"]:
                fe.respond(rid, 403, [("Content-Type", "application/json")], b'{"error": "websocket rejected"}')
        except asyncio.CancelledError:
            raise
        except Exception:
            logger.exception("unhandled error in websocket %s", path)
            if st["accepted"] and not st["closed"]:
                fe.ws_close(rid, 1011)
        finally:
            self._wsq.pop(rid, None)


class _Lifespan:
    """ASGI lifespan protocol (startup / shutdown events of the FastAPI app)"""

    def __init__(self, app: Any):
        self.app = app
        self.rx: asyncio.Queue = asyncio.Queue()
        self.tx: asyncio.Queue = asyncio.Queue()
        self.task: Optional[asyncio.Task] = None
        self.supported = True

    async def _run(self):
        try:
            await self.app({"type": "lifespan", "asgi": {"version": "3.0", "spec_version": "2.0"}}, self.rx.get, self.tx.put)
        except Exception:
            self.supported = False
        finally:
            await self.tx.put({"type": "lifespan.closed"})

    async def startup(self):
        self.task = asyncio.ensure_future(self._run())
        await self.rx.put({"type": "lifespan.startup"})
        msg = await self.tx.get()
        if msg["type"] == "lifespan.startup.failed":
            raise RuntimeError(f"application start-up failed: {msg.get('message')}")

    async def shutdown(self):
        if self.task is None or self.task.done():
            return
        await self.rx.put({"type": "lifespan.shutdown"})
        try:
            await asyncio.wait_for(self.tx.get(), timeout=10)
        except asyncio.TimeoutError:
            self.task.cancel()


NativeAsgiServer._handle_ws = _WsMixin._handle_ws
