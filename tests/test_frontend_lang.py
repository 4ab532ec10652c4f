"""sgl frontend language on the in-process engine backend and over HTTP (reference: test/registered/lang/*)."""
import re
import threading

import pytest

import sglang_b200 as sgl
from sglang_b200.lang.backend.runtime_endpoint import EngineBackend


@pytest.fixture(scope="module")
def backend():
    eng = sgl.Engine(model_path="dummy:tiny-llama", page_size=4, grammar_backend="regex_fsm")
    b = EngineBackend(eng)
    sgl.set_default_backend(b)
    yield b
    eng.shutdown()


def test_gen_select_roles_fork(backend):
    @sgl.function
    def qa(s, question):
        s += sgl.system("You are terse.")
        s += sgl.user(question)
        s += sgl.assistant(sgl.gen("answer", max_tokens=6, temperature=0))
        s += sgl.user("pick one")
        s += sgl.assistant(sgl.select("choice", choices=["alpha", "beta", "gamma"]))

    st = qa.run(question="what is 1+1?")
    assert st["choice"] in ("alpha", "beta", "gamma")
    assert isinstance(st["answer"], str) and st.get_meta_info("answer")["completion_tokens"] == 6
    msgs = st.messages()
    assert [m["role"] for m in msgs] == ["system", "user", "assistant", "user", "assistant"]
    assert msgs[-1]["content"] == st["choice"]

    @sgl.function
    def branches(s):
        s += "Root text. "
        forks = s.fork(3)
        for i, f in enumerate(forks):
            f += f"branch {i}: " + sgl.gen("x", max_tokens=3, temperature=0)
        forks.join()
        s += "done"

    st = branches.run()
    assert len(st["x"]) == 3 and st.text().endswith("done")


def test_regex_gen_and_batch_and_stream(backend):
    @sgl.function
    def ip(s, host):
        s += f"The IP of {host} is " + sgl.gen("ip", max_tokens=24, regex=r"((25[0-5]|2[0-4]\d|[01]?\d\d?)\.){3}(25[0-5]|2[0-4]\d|[01]?\d\d?)", temperature=0.6)

    states = ip.run_batch([{"host": "a"}, {"host": "b"}, {"host": "c"}], num_threads=1)
    for st in states:
        assert re.fullmatch(r"(\d{1,3}\.){3}\d{1,3}", st["ip"]), st["ip"]

    @sgl.function
    def story(s):
        s += "Once " + sgl.gen("t", max_tokens=8, temperature=0)

    st = story.run(stream=True)
    chunks = list(st.text_iter("t"))
    assert "".join(chunks) == st["t"]
    # radix cache served the shared prefix of the batch
    info = backend.get_server_info()["internal_states"][0]
    assert info["num_cached_tokens"] > 0


def test_runtime_endpoint_over_http(backend):
    """RuntimeEndpoint against a real uvicorn server thread"""
    import socket
    import time

    import uvicorn

    from sglang_b200.entrypoints.http_server import create_app
    from sglang_b200.lang.backend.runtime_endpoint import RuntimeEndpoint, _wait_http

    eng = backend.engine
    eng.start_background_loop()
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    server = uvicorn.Server(uvicorn.Config(create_app(eng), host="127.0.0.1", port=port, log_level="error"))
    t = threading.Thread(target=server.run, daemon=True)
    t.start()
    try:
        _wait_http(f"http://127.0.0.1:{port}", 30)
        ep = RuntimeEndpoint(f"http://127.0.0.1:{port}")

        @sgl.function
        def f(s):
            s += "Hello " + sgl.gen("w", max_tokens=4, temperature=0) + sgl.select("c", choices=["yes", "no"])

        st = f.run(backend=ep)
        assert st["c"] in ("yes", "no") and st.get_meta_info("w")["completion_tokens"] == 4
        assert ep.flush_cache() in (True, False)
    finally:
        server.should_exit = True
        t.join(timeout=10)
        eng.shutdown()


def test_openai_and_anthropic_backends_against_compatible_endpoints(backend):
    """sgl programs through the OpenAI-compatible and Anthropic-compatible backends (lang/backend/openai.py), served by this engine's own
    /v1/chat/completions, /v1/completions and /v1/messages routes on a real uvicorn thread"""
    import socket

    import uvicorn

    from sglang_b200.entrypoints.http_server import create_app
    from sglang_b200.lang.backend.runtime_endpoint import _wait_http

    eng = backend.engine
    eng.start_background_loop()
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    server = uvicorn.Server(uvicorn.Config(create_app(eng), host="127.0.0.1", port=port, log_level="error"))
    t = threading.Thread(target=server.run, daemon=True)
    t.start()
    try:
        _wait_http(f"http://127.0.0.1:{port}", 30)

        @sgl.function
        def chat(s, q):
            s += sgl.system("You are terse.")
            s += sgl.user(q)
            s += sgl.assistant(sgl.gen("a", max_tokens=6, temperature=0))
            s += sgl.user("and again?")
            s += sgl.assistant(sgl.gen("b", max_tokens=4, temperature=0))

        oa = sgl.OpenAI("dummy", base_url=f"http://127.0.0.1:{port}/v1")
        st = chat.run(q="hi", backend=oa)
        assert isinstance(st["a"], str) and isinstance(st["b"], str)
        assert st.get_meta_info("a")["completion_tokens"] <= 6 and oa.token_usage["completion_tokens"] > 0
        assert [m["role"] for m in st.messages()] == ["system", "user", "assistant", "user", "assistant"]
        # the same answer as the chat route called by hand (the backend sends role-tagged messages, the server applies the template)
        import requests

        want = requests.post(f"http://127.0.0.1:{port}/v1/chat/completions", json={"model": "dummy", "max_tokens": 6, "temperature": 0, "messages": [
            {"role": "system", "content": "You are terse."}, {"role": "user", "content": "hi"}]}).json()["choices"][0]["message"]["content"]
        assert st["a"] == want
        # streaming + completion-style model
        st2 = chat.run(q="hi", backend=oa, stream=True)
        assert "".join(st2.text_iter("a")) == st["a"]
        comp = sgl.OpenAI("dummy", base_url=f"http://127.0.0.1:{port}/v1", is_chat_model=False)

        @sgl.function
        def plain(s):
            s += "Hello " + sgl.gen("w", max_tokens=4, temperature=0)

        assert isinstance(plain.run(backend=comp)["w"], str)
        # Anthropic Messages API
        an = sgl.Anthropic("dummy", base_url=f"http://127.0.0.1:{port}")
        st3 = chat.run(q="hi", backend=an)
        assert isinstance(st3["a"], str) and st3.get_meta_info("a")["completion_tokens"] <= 6
    finally:
        server.should_exit = True
        t.join(timeout=10)
        eng.shutdown()
