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


def _children_of(pid: int):
    import psutil

    try:
        return [c.pid for c in psutil.Process(pid).children(recursive=True)]
    except psutil.NoSuchProcess:
        return []


@pytest.mark.parametrize("sig", ["TERM", "KILL"])
def test_server_leaves_no_orphans(sig, tmp_path):
    """SIGTERM runs the orderly shutdown (uvicorn re-raises the captured signal; utils/process_tree.graceful_sigterm turns it into the finally
    path); SIGKILL of the front end still takes the scheduler / detokenizer down through PR_SET_PDEATHSIG (reference watchdog.py:166-223)"""
    import os
    import signal
    import socket
    import subprocess
    import sys
    import time
    import urllib.request

    import psutil

    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = subprocess.Popen([sys.executable, "-m", "sglang_b200.launch_server", "--model-path", "dummy:tiny-llama", "--device", "cpu", "--port", str(port),
                          "--host", "127.0.0.1", "--log-level", "error", "--enable-multi-process-frontend"], cwd=root,
                         env=dict(os.environ, PYTHONPATH=root), stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    try:
        for _ in range(120):
            try:
                if urllib.request.urlopen(f"http://127.0.0.1:{port}/health", timeout=1).status == 200:
                    break
            except Exception:
                time.sleep(0.5)
        else:
            pytest.fail("server did not come up")
        kids = _children_of(p.pid)
        assert len(kids) >= 2, kids
        p.send_signal(signal.SIGTERM if sig == "TERM" else signal.SIGKILL)
        p.wait(timeout=30)
        t_end = time.time() + 15
        while time.time() < t_end and any(psutil.pid_exists(k) and psutil.Process(k).status() != psutil.STATUS_ZOMBIE for k in kids):
            time.sleep(0.25)
        left = [k for k in kids if psutil.pid_exists(k) and psutil.Process(k).status() != psutil.STATUS_ZOMBIE]
        assert not left, f"orphaned children after SIG{sig}: {left}"
    finally:
        if p.poll() is None:
            p.kill()
        for k in _children_of(p.pid):
            try:
                os.kill(k, signal.SIGKILL)
            except OSError:
                pass


def test_subprocess_watchdog_reports_dead_child():
    import subprocess
    import sys
    import time

    from sglang_b200.utils.process_tree import SubprocessWatchdog

    ok = subprocess.Popen([sys.executable, "-c", "import time; time.sleep(30)"])
    bad = subprocess.Popen([sys.executable, "-c", "import sys; sys.exit(3)"])
    seen = []
    dog = SubprocessWatchdog([ok, bad], ["ok", "bad"], interval_s=0.05, on_death=lambda n, c: seen.append((n, c)))
    dog.start()
    t0 = time.time()
    while not seen and time.time() - t0 < 10:
        time.sleep(0.05)
    dog.stop(); ok.kill(); ok.wait()
    assert seen == [("bad", 3)]


@pytest.mark.parametrize("frontend", ["uvicorn", "native"])
def test_sigterm_drains_in_flight_requests(tmp_path, frontend):
    """graceful shutdown (both socket layers: uvicorn and the C++ epoll front end): a generation that is running when SIGTERM arrives still gets its full answer; new connections are refused"""
    import os
    import signal
    import socket
    import subprocess
    import sys
    import threading
    import time
    import urllib.request

    import requests

    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = subprocess.Popen([sys.executable, "-m", "sglang_b200.launch_server", "--model-path", "dummy:tiny-llama", "--device", "cpu", "--port", str(port),
                          "--host", "127.0.0.1", "--log-level", "error", "--graceful-shutdown-timeout", "20", "--http-frontend", frontend], cwd=root,
                         env=dict(os.environ, PYTHONPATH=root), stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    try:
        for _ in range(120):
            try:
                if urllib.request.urlopen(f"http://127.0.0.1:{port}/health", timeout=1).status == 200:
                    break
            except Exception:
                time.sleep(0.5)
        else:
            pytest.fail("server did not come up")
        out = {}

        def call():
            r = requests.post(f"http://127.0.0.1:{port}/generate", json={"text": "drain me", "sampling_params": {"max_new_tokens": 200, "temperature": 0,
                                                                                                              "ignore_eos": True}}, timeout=60)
            out["status"], out["body"] = r.status_code, r.json()

        t = threading.Thread(target=call)
        t.start()
        time.sleep(0.4)                              # the request is decoding now
        p.send_signal(signal.SIGTERM)
        t.join(timeout=40)
        assert out.get("status") == 200 and out["body"]["meta_info"]["completion_tokens"] == 200, out
        p.wait(timeout=30)
        with pytest.raises(Exception):
            requests.get(f"http://127.0.0.1:{port}/health", timeout=2)
    finally:
        if p.poll() is None:
            p.kill()


@pytest.mark.parametrize("frontend", ["uvicorn", "native"])
def test_multi_tokenizer_workers_share_one_scheduler(tmp_path, frontend):
    """--tokenizer-worker-num 2 (with either socket layer): two HTTP + tokenizer processes on one port (SO_REUSEPORT), one scheduler, one detokenizer that routes results by
    worker; answers equal the single-process engine, both workers serve traffic, aborts and streaming reach the right worker"""
    import json
    import os
    import signal
    import socket
    import subprocess
    import sys
    import time
    import urllib.request
    from concurrent.futures import ThreadPoolExecutor

    import requests

    import sglang_b200 as sgl

    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = subprocess.Popen([sys.executable, "-m", "sglang_b200.launch_server", "--model-path", "dummy:tiny-llama", "--device", "cpu", "--port", str(port),
                          "--host", "127.0.0.1", "--log-level", "error", "--tokenizer-worker-num", "2", "--page-size", "4", "--http-frontend", frontend], cwd=root,
                         env=dict(os.environ, PYTHONPATH=root), stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    try:
        for _ in range(160):
            try:
                if urllib.request.urlopen(f"http://127.0.0.1:{port}/health", timeout=1).status == 200:
                    break
            except Exception:
                time.sleep(0.5)
        else:
            pytest.fail("server did not come up")
        time.sleep(1.0)                                   # let the second worker finish binding
        ref = sgl.Engine(model_path="dummy:tiny-llama", page_size=4, log_level="error")
        prompts = [f"prompt number {i} about things" for i in range(24)]
        sp = {"max_new_tokens": 6, "temperature": 0, "ignore_eos": True}
        want = [o["output_ids"] for o in ref.generate(prompts, sp)]
        ref.shutdown()

        def call(i):
            with requests.Session() as c:                 # a fresh connection per request: the kernel picks a worker each time
                r = c.post(f"http://127.0.0.1:{port}/generate", json={"text": prompts[i], "sampling_params": sp, "rid": f"user-{i}"}, timeout=60)
                return r.headers.get("x-sglb-worker"), r.json()

        with ThreadPoolExecutor(8) as ex:
            res = list(ex.map(call, range(len(prompts))))
        assert [r[1]["output_ids"] for r in res] == want
        assert all(r[1]["meta_info"]["id"] == f"user-{i}" for i, r in enumerate(res)), "the worker prefix must not leak into request ids"
        assert {r[0] for r in res} == {"0", "1"}, "both workers must have answered"
        # streaming through whichever worker takes the connection
        r = requests.post(f"http://127.0.0.1:{port}/generate", json={"text": "stream me", "sampling_params": sp, "stream": True}, stream=True, timeout=60)
        chunks = [json.loads(l[6:]) for l in r.iter_lines(decode_unicode=True) if l and l.startswith("data: {")]
        assert chunks and chunks[-1]["meta_info"]["completion_tokens"] == 6
        info = requests.get(f"http://127.0.0.1:{port}/get_server_info", timeout=10).json()
        assert info["internal_states"][0]["num_generated_tokens"] >= 24 * 6
        p.send_signal(signal.SIGTERM)
        p.wait(timeout=40)
    finally:
        if p.poll() is None:
            p.kill()
        for k in _children_of(p.pid):
            try:
                os.kill(k, signal.SIGKILL)
            except OSError:
                pass


def test_cli_killall_stops_registered_servers_by_pid(tmp_path):
    """`cli killall`: servers register pid + port in a directory; killall sends SIGTERM to exactly those pids (graceful drain) — no pattern kills"""
    import json
    import os
    import socket
    import subprocess
    import sys
    import time
    import urllib.request

    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, PYTHONPATH=root, SGLB_PID_DIR=str(tmp_path / "pids"))
    p = subprocess.Popen([sys.executable, "-m", "sglang_b200.launch_server", "--model-path", "dummy:tiny-llama", "--device", "cpu", "--port", str(port),
                          "--host", "127.0.0.1", "--log-level", "error"], cwd=root, env=env, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    try:
        for _ in range(120):
            try:
                if urllib.request.urlopen(f"http://127.0.0.1:{port}/health", timeout=1).status == 200:
                    break
            except Exception:
                time.sleep(0.5)
        else:
            pytest.fail("server did not come up")
        ls = json.loads(subprocess.run([sys.executable, "-m", "sglang_b200.cli", "killall", "--list"], cwd=root, env=env, capture_output=True, text=True).stdout)
        assert [r["pid"] for r in ls] == [p.pid] and ls[0]["port"] == port
        other = subprocess.run([sys.executable, "-m", "sglang_b200.cli", "killall", "--port", str(port + 1)], cwd=root, env=env, capture_output=True, text=True)
        assert json.loads(other.stdout) == {"stopped": 0} and p.poll() is None
        out = subprocess.run([sys.executable, "-m", "sglang_b200.cli", "killall", "--grace", "20"], cwd=root, env=env, capture_output=True, text=True)
        assert json.loads(out.stdout) == {"stopped": 1}
        p.wait(timeout=30)
        assert json.loads(subprocess.run([sys.executable, "-m", "sglang_b200.cli", "killall", "--list"], cwd=root, env=env, capture_output=True,
                                         text=True).stdout) == []
    finally:
        if p.poll() is None:
            p.kill()
