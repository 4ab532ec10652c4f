"""Front end | scheduler process | detokenizer process over ZMQ (reference: engine.py:905 _launch_subprocesses, tokenizer_manager.py,
detokenizer_manager.py): same answers as the in-process engine, streaming, control RPCs, clean shutdown."""
import json

import pytest

import sglang_b200 as sgl


@pytest.fixture(scope="module")
def front():
    from sglang_b200.managers.multiprocess import FrontendEngine

    f = FrontendEngine(model_path="dummy:tiny-llama", page_size=4, log_level="error", enable_metrics=True)
    yield f
    f.shutdown()
    assert not f._p_sched.is_alive() and not f._p_detok.is_alive()


def test_generate_matches_in_process_engine(front):
    ref = sgl.Engine(model_path="dummy:tiny-llama", page_size=4, log_level="error")
    sp = {"max_new_tokens": 8, "temperature": 0, "ignore_eos": True}
    prompts = ["hello world", "the quick brown fox", "a"]
    want = ref.generate(prompts, sp)
    ref.shutdown()
    got = front.generate(prompts, sp)
    assert [g["output_ids"] for g in got] == [w["output_ids"] for w in want]
    assert [g["text"] for g in got] == [w["text"] for w in want]
    assert all(g["meta_info"]["completion_tokens"] == 8 and g["meta_info"]["finish_reason"]["type"] == "length" for g in got)


def test_http_stream_and_control_over_processes(front):
    from fastapi.testclient import TestClient

    from sglang_b200.entrypoints.http_server import create_app

    with TestClient(create_app(front)) as c:
        assert c.get("/health_generate").status_code == 200
        r = c.post("/generate", json={"text": "stream me", "sampling_params": {"max_new_tokens": 6, "temperature": 0, "ignore_eos": True}, "stream": True})
        chunks = [json.loads(ln[6:]) for ln in r.text.splitlines() if ln.startswith("data: {")]
        assert chunks and chunks[-1]["meta_info"]["completion_tokens"] == 6
        assert chunks[-1]["text"].startswith(chunks[0]["text"])
        r = c.post("/v1/chat/completions", json={"model": "m", "messages": [{"role": "user", "content": "hi"}], "max_tokens": 4, "temperature": 0})
        assert r.status_code == 200 and r.json()["usage"]["completion_tokens"] <= 4
        info = c.get("/get_server_info").json()
        assert info["internal_states"][0]["page_size"] == 4 and info["internal_states"][0]["num_generated_tokens"] > 0
        assert c.post("/flush_cache").status_code == 200
        assert "sglang:" in c.get("/metrics").text
        assert c.get("/v1/models").json()["data"][0]["max_model_len"] > 0
    assert front.get_load()["num_running_reqs"] == 0
