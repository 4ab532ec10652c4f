"""Native C++ HTTP front end (csrc/runtime/http_frontend.cpp + entrypoints/native_http.py) serving the real FastAPI app over real sockets:
keep-alive, pipelining, SSE streaming, static paths answered by the I/O thread, client-disconnect aborts, limits, graceful drain.
(reference: the embedded Rust server, rust/sglang-server, is exercised by the same OpenAI-route tests with SGLANG_RUST_SERVER=1)"""
import json
import socket
import threading
import time

import pytest
import requests

import sglang_b200 as sgl
from sglang_b200.entrypoints.http_server import create_app
from sglang_b200.entrypoints.native_http import NativeAsgiServer


@pytest.fixture(scope="module")
def server():
    eng = sgl.Engine(model_path="dummy:tiny-llama", page_size=4, enable_metrics=True)
    eng.start_background_loop()
    app = create_app(eng)
    srv = NativeAsgiServer(app, "127.0.0.1", 0, graceful_timeout=5.0, max_body=1 << 20,
                           static={"/health": (200, "application/json", b"{}")})
    app.state.native_frontend = srv.fe
    th = threading.Thread(target=srv.run, kwargs={"install_signal_handlers": False}, daemon=True)
    th.start()
    assert srv.started.wait(30)
    yield srv, eng, f"http://127.0.0.1:{srv.port}"
    srv.request_stop()
    th.join(20)
    eng.shutdown()


def test_routes_keepalive_and_static(server):
    srv, eng, url = server
    s = requests.Session()
    before = srv.fe.stats()
    for _ in range(3):
        assert s.get(url + "/health").status_code == 200
    r = s.post(url + "/generate", json={"text": "hello", "sampling_params": {"max_new_tokens": 5, "temperature": 0}})
    assert r.status_code == 200 and r.json()["meta_info"]["completion_tokens"] == 5
    assert s.get(url + "/v1/models").json()["data"][0]["id"] == "dummy:tiny-llama"
    assert s.get(url + "/no/such/route").status_code == 404
    assert s.post(url + "/generate", json={"text": "x", "sampling_params": {"top_p": 7}}).status_code == 400
    st = srv.fe.stats()
    assert st["static_hits"] - before["static_hits"] == 3               # liveness probes never reached Python
    assert st["accepted"] - before["accepted"] == 1                     # one TCP connection for the whole session
    m = s.get(url + "/metrics").text
    assert "http_static_hits" in m and "http_accepted" in m          # the socket layer's counters are exported next to the scheduler's


def test_sse_stream_matches_unary(server):
    _, _, url = server
    body = {"text": "hello", "sampling_params": {"max_new_tokens": 6, "temperature": 0}}
    ref = requests.post(url + "/generate", json=body).json()
    with requests.post(url + "/generate", json={**body, "stream": True}, stream=True) as r:
        assert r.headers.get("transfer-encoding") == "chunked" and r.headers["content-type"].startswith("text/event-stream")
        lines = [l.decode() for l in r.iter_lines() if l.startswith(b"data:")]
    assert lines[-1] == "data: [DONE]" and json.loads(lines[-2][5:])["output_ids"] == ref["output_ids"]
    with requests.post(url + "/v1/chat/completions", json={"model": "m", "messages": [{"role": "user", "content": "ping"}], "max_tokens": 4,
                                                           "stream": True, "stream_options": {"include_usage": True}}, stream=True) as r:
        chunks = [json.loads(l[5:]) for l in r.iter_lines() if l.startswith(b"data:") and b"[DONE]" not in l]
    assert chunks[0]["choices"][0]["delta"]["role"] == "assistant" and chunks[-1]["usage"]["completion_tokens"] == 4


def _raw(port, payload: bytes, read_until: bytes = None, timeout=10.0) -> bytes:
    s = socket.create_connection(("127.0.0.1", port), timeout=timeout)
    s.sendall(payload)
    buf = b""
    try:
        while True:
            d = s.recv(65536)
            if not d:
                break
            buf += d
            if read_until is not None and buf.count(read_until) >= 1 and buf.endswith(read_until):
                break
    except socket.timeout:
        pass
    s.close()
    return buf


def test_pipelining_expect_continue_and_limits(server):
    srv, _, _ = server
    body = json.dumps({"text": "hi", "sampling_params": {"max_new_tokens": 2, "temperature": 0}}).encode()
    req = b"POST /generate HTTP/1.1\r\nHost: x\r\nContent-Type: application/json\r\nContent-Length: %d\r\n\r\n" % len(body) + body
    # two pipelined requests + a closing one on a single connection: answered in order
    out = _raw(srv.port, req + req + b"GET /health HTTP/1.1\r\nHost: x\r\nConnection: close\r\n\r\n")
    assert out.count(b"HTTP/1.1 200 OK") == 3 and out.rstrip().endswith(b"{}") and b"Connection: close" in out
    # Expect: 100-continue is acknowledged before the body is sent
    s = socket.create_connection(("127.0.0.1", srv.port), timeout=10)
    s.sendall(b"POST /generate HTTP/1.1\r\nHost: x\r\nContent-Type: application/json\r\nExpect: 100-continue\r\nContent-Length: %d\r\n\r\n" % len(body))
    assert s.recv(100).startswith(b"HTTP/1.1 100 Continue")
    s.sendall(body)
    buf = b""
    while b"output_ids" not in buf:
        buf += s.recv(65536)
    s.close()
    assert b"HTTP/1.1 200 OK" in buf
    # limits and malformed input are refused by the I/O thread itself
    assert _raw(srv.port, b"POST /generate HTTP/1.1\r\nContent-Length: 99999999\r\n\r\n").startswith(b"HTTP/1.1 413")
    assert _raw(srv.port, b"POST /generate HTTP/1.1\r\nTransfer-Encoding: chunked\r\n\r\n5\r\nhello\r\n0\r\n\r\n").startswith(b"HTTP/1.1 411")
    assert _raw(srv.port, b"garbage\r\n\r\n").startswith(b"HTTP/1.1 400")


def test_client_disconnect_aborts_generation(server):
    srv, eng, url = server
    body = json.dumps({"text": "hi", "stream": True, "sampling_params": {"max_new_tokens": 100000, "temperature": 0, "ignore_eos": True}}).encode()
    s = socket.create_connection(("127.0.0.1", srv.port), timeout=10)
    s.sendall(b"POST /generate HTTP/1.1\r\nHost: x\r\nContent-Type: application/json\r\nContent-Length: %d\r\n\r\n" % len(body) + body)
    buf = b""
    while buf.count(b"data:") < 3:
        buf += s.recv(65536)
    d0 = srv.fe.stats()["disconnects"]
    s.close()                                  # walk away mid-stream
    t0 = time.time()
    while time.time() - t0 < 20:
        if srv.fe.stats()["disconnects"] > d0 and requests.get(url + "/get_load").json()["num_running_reqs"] == 0:
            break
        time.sleep(0.1)
    assert srv.fe.stats()["disconnects"] > d0
    assert requests.get(url + "/get_load").json()["num_running_reqs"] == 0          # the scheduler dropped the abandoned request


def test_concurrent_clients(server):
    _, _, url = server
    outs = [None] * 16

    def one(i):
        outs[i] = requests.post(url + "/generate", json={"text": "p%d" % i, "sampling_params": {"max_new_tokens": 4, "temperature": 0}}).json()

    ths = [threading.Thread(target=one, args=(i,)) for i in range(16)]
    [t.start() for t in ths]
    [t.join(60) for t in ths]
    assert all(o is not None and o["meta_info"]["completion_tokens"] == 4 for o in outs)


def test_graceful_drain_finishes_inflight_request():
    eng = sgl.Engine(model_path="dummy:tiny-llama", page_size=4)
    eng.start_background_loop()
    srv = NativeAsgiServer(create_app(eng), "127.0.0.1", 0, graceful_timeout=20.0)
    th = threading.Thread(target=srv.run, kwargs={"install_signal_handlers": False}, daemon=True)
    th.start()
    assert srv.started.wait(30)
    url = f"http://127.0.0.1:{srv.port}"
    res = {}

    def long_req():
        res["r"] = requests.post(url + "/generate", json={"text": "hi", "sampling_params": {"max_new_tokens": 200, "temperature": 0, "ignore_eos": True}})

    t = threading.Thread(target=long_req)
    t.start()
    time.sleep(0.3)
    srv.request_stop()                          # SIGTERM path: stop accepting, let the running generation finish
    t.join(30)
    th.join(30)
    assert res["r"].status_code == 200 and res["r"].json()["meta_info"]["completion_tokens"] == 200
    with pytest.raises(requests.exceptions.ConnectionError):
        requests.get(url + "/health", timeout=2)
    eng.shutdown()


def test_bench_one_batch_server_against_a_live_server(server, tmp_path):
    """the server-side static-batch benchmark (reference bench_one_batch_server.py) drives a real socket end to end"""
    from sglang_b200 import bench_one_batch_server as B

    _, _, url = server
    out = tmp_path / "rows.jsonl"
    rows = B.main(["--base-url", url, "--batch-size", "1", "3", "--input-len", "12", "--output-len", "5", "--result-file", str(out)])
    assert [r["batch_size"] for r in rows] == [1, 3] and all(r["completed_tokens"] == r["batch_size"] * 5 for r in rows)
    assert all(r["ttft_s"] > 0 and r["latency_s"] >= r["ttft_s"] and r["itl_ms"] > 0 for r in rows)
    assert len(out.read_text().strip().splitlines()) == 2


def test_websocket_echo_and_realtime_route_over_the_native_front_end(server):
    """RFC 6455 on the C++ I/O thread (handshake hash, masking, fragments, ping/pong, close) + the ASGI websocket scope: a plain echo app,
    then the engine's /v1/realtime text session"""
    from websockets.sync.client import connect

    async def echo(scope, receive, send):
        assert scope["type"] == "websocket" and scope["path"] == "/echo"
        assert (await receive())["type"] == "websocket.connect"
        await send({"type": "websocket.accept"})
        while True:
            m = await receive()
            if m["type"] == "websocket.disconnect":
                return
            if m.get("text") == "bye":
                await send({"type": "websocket.close", "code": 1000})
                return
            await send({"type": "websocket.send", **({"text": m["text"].upper()} if m.get("text") is not None else {"bytes": m["bytes"][::-1]})})

    srv2 = NativeAsgiServer(echo, "127.0.0.1", 0)
    th = threading.Thread(target=srv2.run, kwargs={"install_signal_handlers": False}, daemon=True)
    th.start()
    assert srv2.started.wait(10)
    try:
        with connect(f"ws://127.0.0.1:{srv2.port}/echo") as ws:
            ws.send("hello")
            assert ws.recv() == "HELLO"
            ws.send(b"\x01\x02\x03")
            assert ws.recv() == b"\x03\x02\x01"
            big = "x" * 200000                          # 64-bit length header, several TCP segments
            ws.send(big)
            assert ws.recv() == big.upper()
            ws.ping().wait(5)                           # answered by the I/O thread
            ws.send("bye")
            with pytest.raises(Exception):
                ws.recv()
    finally:
        srv2.request_stop()
        th.join(10)
    srv, eng, url = server
    with connect(f"ws://127.0.0.1:{srv.port}/v1/realtime") as ws:
        first = json.loads(ws.recv())
        assert first["type"] == "session.created"
        ws.send(json.dumps({"type": "conversation.item.create", "item": {"type": "message", "role": "user", "content": [{"type": "input_text", "text": "hi"}]}}))
        ws.send(json.dumps({"type": "response.create", "response": {"max_output_tokens": 4}}))
        kinds = []
        for _ in range(50):
            ev = json.loads(ws.recv())
            kinds.append(ev["type"])
            if ev["type"] in ("response.done", "error"):
                break
        assert "response.done" in kinds and any(k.endswith("delta") for k in kinds), kinds


def test_bench_serving_random_image_dataset_against_a_vlm_server(tmp_path):
    """bench_serving --dataset-name random-image: requests carry random PNGs as data URLs next to placeholder tokens; a tiny Qwen2-VL served by
    the native front end answers all of them (preprocessing + tower + splice under concurrency)"""
    from sglang_b200 import bench_serving as B

    eng = sgl.Engine(model_path="dummy:tiny-qwen2-vl", device="cpu", page_size=4, log_level="error")
    eng.start_background_loop()
    srv = NativeAsgiServer(create_app(eng), "127.0.0.1", 0, graceful_timeout=3.0)
    th = threading.Thread(target=srv.run, kwargs={"install_signal_handlers": False}, daemon=True)
    th.start()
    assert srv.started.wait(30)
    try:
        m = B.main(["--base-url", f"http://127.0.0.1:{srv.port}", "--dataset-name", "random-image", "--num-prompts", "5", "--random-input-len", "12",
                    "--random-output-len", "4", "--image-token-id", str(eng.model_config.image_token_id), "--image-count", "2",
                    "--image-resolution", "70x56", "--vocab-size", str(eng.model_config.vocab_size), "--request-rate", "inf"])
        assert m["completed"] == 5 and m["total_output_tokens"] == 20, m
        assert m["total_input_tokens"] == 5 * (12 + 2)                    # as sent: 12 text ids + 2 placeholders (the server expands them)
    finally:
        srv.request_stop()
        th.join(20)
        eng.shutdown()
