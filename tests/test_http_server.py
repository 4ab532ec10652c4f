"""HTTP + OpenAI-compatible surface (in-process ASGI client; reference: test/registered/openai_server/*)."""
import json

import pytest

import sglang_b200 as sgl


class _SocketClient:
    """the slice of the TestClient API these tests use, over real sockets (requests) — lets the same route tests run against the native C++ HTTP
    front end (entrypoints/native_http.py)"""

    def __init__(self, base):
        import requests

        self.base, self.s = base, requests.Session()

    def get(self, path, **kw):
        return self.s.get(self.base + path, **kw)

    def post(self, path, **kw):
        return self.s.post(self.base + path, **kw)

    def stream(self, method, path, **kw):
        import contextlib

        @contextlib.contextmanager
        def cm():
            r = self.s.request(method, self.base + path, stream=True, **kw)

            class _R:
                status_code = r.status_code
                headers = r.headers

                @staticmethod
                def iter_lines():
                    for l in r.iter_lines():
                        yield l.decode() if isinstance(l, bytes) else l
            try:
                yield _R
            finally:
                r.close()
        return cm()


@pytest.fixture(scope="module", params=["asgi", "native-socket"])
def client(request):
    from fastapi.testclient import TestClient

    from sglang_b200.entrypoints.http_server import create_app

    eng = sgl.Engine(model_path="dummy:tiny-llama", page_size=4, enable_metrics=True, tool_call_parser="qwen25",
                     reasoning_parser="qwen3", grammar_backend="regex_fsm")
    eng.start_background_loop()
    if request.param == "asgi":
        with TestClient(create_app(eng)) as c:
            yield c
    else:          # every route test again through the C++ epoll front end over a real TCP connection
        import threading

        from sglang_b200.entrypoints.native_http import NativeAsgiServer

        srv = NativeAsgiServer(create_app(eng), "127.0.0.1", 0, graceful_timeout=3.0)
        th = threading.Thread(target=srv.run, kwargs={"install_signal_handlers": False}, daemon=True)
        th.start()
        assert srv.started.wait(30)
        yield _SocketClient(f"http://127.0.0.1:{srv.port}")
        srv.request_stop()
        th.join(20)
    eng.shutdown()


def test_health_and_info(client):
    assert client.get("/health").status_code == 200
    assert client.get("/health_generate").status_code == 200
    info = client.get("/get_server_info").json()
    assert info["internal_states"][0]["page_size"] == 4
    assert client.get("/v1/models").json()["data"][0]["id"] == "dummy:tiny-llama"
    assert client.get("/get_load").json()["num_running_reqs"] == 0


def test_native_generate_and_stream(client):
    r = client.post("/generate", json={"text": "hello", "sampling_params": {"max_new_tokens": 5, "temperature": 0}})
    out = r.json()
    assert out["meta_info"]["completion_tokens"] == 5 and len(out["output_ids"]) == 5
    r2 = client.post("/generate", json={"text": ["a", "bb"], "sampling_params": {"max_new_tokens": 3}})
    assert len(r2.json()) == 2
    with client.stream("POST", "/generate", json={"text": "hello", "stream": True, "sampling_params": {"max_new_tokens": 5, "temperature": 0}}) as s:
        lines = [l for l in s.iter_lines() if l.startswith("data:")]
    assert lines[-1] == "data: [DONE]"
    last = json.loads(lines[-2][5:])
    assert last["output_ids"] == out["output_ids"]
    assert client.post("/generate", json={"text": "x", "sampling_params": {"top_p": 7}}).status_code == 400


def test_openai_completions_and_chat(client):
    r = client.post("/v1/completions", json={"model": "m", "prompt": "hi there", "max_tokens": 6, "temperature": 0, "logprobs": 2})
    j = r.json()
    assert j["usage"]["completion_tokens"] == 6 and len(j["choices"][0]["logprobs"]["token_logprobs"]) == 6
    r = client.post("/v1/chat/completions", json={"model": "m", "messages": [{"role": "user", "content": "ping"}], "max_tokens": 4,
                                                  "temperature": 0})
    j = r.json()
    assert j["choices"][0]["message"]["role"] == "assistant" and j["choices"][0]["finish_reason"] == "length"
    with client.stream("POST", "/v1/chat/completions", json={"model": "m", "messages": [{"role": "user", "content": "ping"}],
                                                             "max_tokens": 4, "stream": True, "stream_options": {"include_usage": True}}) as s:
        chunks = [json.loads(l[5:]) for l in s.iter_lines() if l.startswith("data:") and "[DONE]" not in l]
    assert chunks[0]["choices"][0]["delta"]["role"] == "assistant"
    assert chunks[-1]["usage"]["completion_tokens"] == 4
    # response_format -> grammar
    r = client.post("/v1/chat/completions", json={"model": "m", "messages": [{"role": "user", "content": "json"}], "max_tokens": 60,
                                                  "response_format": {"type": "regex", "pattern": "(ab)+c"}, "temperature": 0.5,
                                                  "chat_template_kwargs": {"enable_thinking": False}})
    import re

    assert re.fullmatch("(ab)+c", r.json()["choices"][0]["message"]["content"])


def test_tokenize_embeddings_metrics_admin(client):
    t = client.post("/v1/tokenize", json={"prompt": "abc"}).json()
    assert t["count"] == 3
    assert client.post("/v1/detokenize", json={"tokens": t["tokens"]}).json()["text"] == "abc"
    e = client.post("/v1/embeddings", json={"input": ["x", "yy"]}).json()
    assert len(e["data"]) == 2 and abs(sum(v * v for v in e["data"][0]["embedding"]) - 1) < 1e-3
    assert b"sglang:num_requests_total" in client.get("/metrics").content
    assert client.post("/flush_cache").status_code == 200
    w = client.post("/get_weights_by_name", json={"name": "model.norm.weight", "truncate_size": 3}).json()
    assert w == [1.0, 1.0, 1.0]
    up = client.post("/update_weights_from_tensor", json={"named_tensors": [["model.norm.weight", [2.0] * 64]]}).json()
    assert up["success"]
    assert client.post("/get_weights_by_name", json={"name": "model.norm.weight", "truncate_size": 2}).json() == [2.0, 2.0]


def test_anthropic_messages(client):
    body = {"model": "m", "max_tokens": 5, "temperature": 0, "system": "be brief",
            "messages": [{"role": "user", "content": [{"type": "text", "text": "ping"}]}]}
    j = client.post("/v1/messages", json=body).json()
    assert j["type"] == "message" and j["role"] == "assistant" and j["stop_reason"] == "max_tokens"
    assert j["usage"]["output_tokens"] == 5 and j["content"][0]["type"] == "text"
    with client.stream("POST", "/v1/messages", json={**body, "stream": True}) as s:
        events = [l[7:] for l in s.iter_lines() if l.startswith("event: ")]
    assert events[0] == "message_start" and events[-1] == "message_stop" and "content_block_delta" in events
    assert client.post("/v1/messages/count_tokens", json=body).json()["input_tokens"] > 0


def test_ollama_api(client):
    assert client.get("/api/tags").json()["models"][0]["name"] == "dummy:tiny-llama"
    j = client.post("/api/generate", json={"model": "m", "prompt": "hello", "stream": False, "options": {"num_predict": 4, "temperature": 0}}).json()
    assert j["done"] and j["eval_count"] == 4 and isinstance(j["response"], str)
    with client.stream("POST", "/api/generate", json={"model": "m", "prompt": "hello", "options": {"num_predict": 4, "temperature": 0}}) as s:
        lines = [json.loads(l) for l in s.iter_lines() if l]
    assert lines[-1]["done"] and "".join(x["response"] for x in lines) == j["response"]
    j = client.post("/api/chat", json={"model": "m", "messages": [{"role": "user", "content": "hi"}], "stream": False,
                                       "options": {"num_predict": 3}}).json()
    assert j["done"] and j["message"]["role"] == "assistant" and j["eval_count"] == 3
    with client.stream("POST", "/api/chat", json={"model": "m", "messages": [{"role": "user", "content": "hi"}], "options": {"num_predict": 3}}) as s:
        lines = [json.loads(l) for l in s.iter_lines() if l]
    assert lines[-1]["done"]


def test_sagemaker_vertex_responses(client):
    assert client.get("/ping").status_code == 200
    j = client.post("/invocations", json={"model": "m", "messages": [{"role": "user", "content": "x"}], "max_tokens": 3}).json()
    assert j["usage"]["completion_tokens"] == 3
    j = client.post("/invocations", json={"text": "abc", "sampling_params": {"max_new_tokens": 2}}).json()
    assert j["meta_info"]["completion_tokens"] == 2
    j = client.post("/vertex_generate", json={"instances": [{"text": "a"}, {"text": "bb"}], "parameters": {"max_new_tokens": 2}}).json()
    assert len(j["predictions"]) == 2 and j["predictions"][0]["meta_info"]["completion_tokens"] == 2
    j = client.post("/v1/responses", json={"model": "m", "input": "hello", "max_output_tokens": 4, "instructions": "short"}).json()
    assert j["object"] == "response" and j["status"] == "completed" and j["usage"]["output_tokens"] == 4
    with client.stream("POST", "/v1/responses", json={"model": "m", "input": [{"role": "user", "content": [{"type": "input_text", "text": "yo"}]}],
                                                      "max_output_tokens": 4, "stream": True}) as s:
        events = [l[7:] for l in s.iter_lines() if l.startswith("event: ")]
    assert events[0] == "response.created" and events[-1] == "response.completed"
