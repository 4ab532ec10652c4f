"""Authorisation policy: deny by default once a key is configured (reference: srt/utils/auth.py decide_request_auth + test/registered/
openai_server auth cases).  Regression for ADVICE r1: /invocations, /flush_cache, /get_weights_by_name, /pause_generation answered 200
without a key when only --api-key was set."""
import types

import grpc
import pytest

import sglang_b200 as sgl
from sglang_b200.utils import auth


def test_policy_table():
    d = auth.decide
    # no keys: everything open
    assert d("POST", "/flush_cache", None, None, None)[0]
    assert d("POST", "/update_weights_from_disk", None, None, None)[0]
    # api key only
    for path in ("/invocations", "/api/generate", "/api/chat", "/vertex_generate", "/tokenize", "/generate", "/v1/chat/completions",
                 "/flush_cache", "/pause_generation", "/abort_request", "/open_session", "/hicache/storage-backend", "/eplb_rebalance"):
        assert d("POST", path, None, "k", None) == (False, 401, "invalid API key"), path
        assert d("POST", path, "Bearer k", "k", None)[0], path
        assert not d("POST", path, "Bearer wrong", "k", None)[0]
        assert not d("POST", path, "k", "k", None)[0]          # not a bearer header
    # weight mutation / read-back / fault injection are refused outright without an admin key
    for path in ("/update_weights_from_disk", "/get_weights_by_name", "/release_memory_occupation", "/slow_down",
                 "/init_weights_update_group", "/send_weights_to_remote_instance"):
        assert d("POST", path, "Bearer k", "k", None)[:2] == (False, 403), path
        assert d("POST", path, "Bearer a", "k", "a")[0]
        assert not d("POST", path, "Bearer k", "k", "a")[0]
    # admin key configured: control-plane routes need it, the API key is not enough
    assert not d("POST", "/flush_cache", "Bearer k", "k", "a")[0]
    assert d("POST", "/flush_cache", "Bearer a", "k", "a")[0]
    # data plane accepts either key
    assert d("POST", "/generate", "Bearer a", "k", "a")[0] and d("POST", "/generate", "Bearer k", "k", "a")[0]
    # probes stay open
    for path in ("/health", "/health_generate", "/metrics", "/ping"):
        assert d("GET", path, None, "k", "a")[0]
    assert d("OPTIONS", "/generate", None, "k", None)[0]


@pytest.fixture(scope="module")
def keyed():
    from fastapi.testclient import TestClient

    from sglang_b200.entrypoints.http_server import create_app

    eng = sgl.Engine(model_path="dummy:tiny-llama", page_size=4, log_level="error", api_key="secret")
    eng.start_background_loop()
    with TestClient(create_app(eng)) as c:
        yield c, eng
    eng.shutdown()


def test_http_routes_need_the_key(keyed):
    c, _ = keyed
    h = {"Authorization": "Bearer secret"}
    assert c.get("/health").status_code == 200
    for method, path, body in (("post", "/invocations", {"text": "hi", "sampling_params": {"max_new_tokens": 1}}),
                               ("post", "/flush_cache", None), ("post", "/pause_generation", None),
                               ("post", "/tokenize", {"prompt": "hi"}), ("post", "/api/generate", {"model": "m", "prompt": "hi"}),
                               ("get", "/get_model_info", None), ("get", "/v1/models", None)):
        r = getattr(c, method)(path, **({"json": body} if body is not None else {}))
        assert r.status_code == 401, (path, r.status_code)
    assert c.post("/get_weights_by_name", json={"name": "x"}, headers=h).status_code == 403
    assert c.post("/flush_cache", headers=h).status_code == 200
    assert c.post("/continue_generation", headers=h).status_code == 200
    r = c.post("/generate", json={"text": "hi", "sampling_params": {"max_new_tokens": 2}}, headers=h)
    assert r.status_code == 200 and r.json()["meta_info"]["completion_tokens"] == 2


def test_grpc_needs_the_key(keyed):
    from sglang_b200.entrypoints.grpc_server import MSG, SglbGrpcClient, make_server

    _, eng = keyed
    server, port = make_server(eng, "127.0.0.1", 0)
    server.start()
    try:
        anon = SglbGrpcClient(f"127.0.0.1:{port}")
        assert anon.call("HealthCheck", MSG["Empty"]()).healthy
        with pytest.raises(grpc.RpcError) as ei:
            anon.call("FlushCache", MSG["Empty"]())
        assert ei.value.code() == grpc.StatusCode.UNAUTHENTICATED
        with pytest.raises(grpc.RpcError) as ei:
            list(anon.call("TextGenerate", MSG["TextGenerateRequest"](text="hi")))
        assert ei.value.code() == grpc.StatusCode.UNAUTHENTICATED
        anon.close()
        good = SglbGrpcClient(f"127.0.0.1:{port}", api_key="secret")
        assert good.call("FlushCache", MSG["Empty"]()).success
        good.close()
    finally:
        server.stop(0)
