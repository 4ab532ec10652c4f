"""Auxiliary / debugging subsystems: tensor invariants, rank consensus + slow-rank detection, tensor dumper/comparator, weights
checker, /slow_down, /set_trace_level, YAML config merge, crash dump, plugins, device timer."""
import json
import os
import sys
import time

import numpy as np
import pytest
import torch

import sglang_b200 as sgl


def test_invariants_levels(monkeypatch):
    from sglang_b200.utils import invariants as I

    t = torch.tensor([0.0, float("nan")])
    assert I.expect(t, I.NotNaN()) is t                          # off: nothing happens
    monkeypatch.setenv("SGLB_INVARIANTS", "crash")
    with pytest.raises(I.InvariantError):
        I.expect(t, I.NotNaN(), name="logits")
    with pytest.raises(I.InvariantError):
        I.expect(torch.tensor([0, 5, 9]), I.InRange(0, 9), name="slots")
    with pytest.raises(I.InvariantError):
        I.expect(torch.tensor([3, 3]), I.Unique())
    I.expect(torch.tensor([0, 32, 64]), I.PageAligned(32), I.InRange(0, 65), I.Unique(), I.Shape(3))
    monkeypatch.setenv("SGLB_INVARIANTS", "log")
    I.expect(torch.tensor([1, 33]), I.PageAligned(32))           # logs, does not raise


def test_digest_and_slow_rank_detector():
    from sglang_b200.utils.rank_consensus import SlowRankDetector, digest

    a = {"ids": np.arange(4), "t": torch.ones(2), "x": [1, "a"]}
    b = {"x": [1, "a"], "t": torch.ones(2), "ids": np.arange(4)}
    assert digest(a) == digest(b) != digest({**a, "x": [2, "a"]})

    class G:
        world_size, rank = 4, 0

        def all_gather_pyobj(self, v):
            return [v, v, v * 2.0, v]
    d = SlowRankDetector(G(), window=4, every=4, ratio=1.5)
    rep = None
    for _ in range(4):
        d.step_begin()
        time.sleep(0.001)
        rep = d.step_end() or rep
    assert rep and rep["slow_ranks"] == [2]


def test_dumper_and_comparator(tmp_path):
    from sglang_b200.debug_utils.dumper import Dumper, compare_dumps, main

    m = torch.nn.Sequential(torch.nn.Linear(4, 4), torch.nn.ReLU(), torch.nn.Linear(4, 2))
    x = torch.randn(3, 4)
    for name, scale in (("a", 1.0), ("b", 1.0), ("c", 1.5)):
        d = Dumper(str(tmp_path / name))
        assert d.hook(m, r"^[02]$") == 2
        with torch.no_grad():
            m[2].weight.mul_(scale)
            m(x)
            m[2].weight.div_(scale)
        d.dump("input", x)
        d.unhook()
    assert all(r["ok"] for r in compare_dumps(str(tmp_path / "a"), str(tmp_path / "b")))
    rows = compare_dumps(str(tmp_path / "a"), str(tmp_path / "c"))
    assert [r["ok"] for r in rows] == [True, False, True]          # layer 0 equal, layer 2 differs, input equal
    assert main([str(tmp_path / "a"), str(tmp_path / "c")]) == 1


def test_weights_checker_slow_down_trace_level_and_crash_dump(tmp_path):
    from fastapi.testclient import TestClient

    from sglang_b200.entrypoints.http_server import create_app

    eng = sgl.Engine(model_path="dummy:tiny-llama", device="cpu", page_size=4, log_level="error", crash_dump_folder=str(tmp_path / "dumps"))
    with TestClient(create_app(eng)) as c:
        j = c.post("/weights_checker", json={"action": "checksum"}).json()
        fp = json.loads(j["message"])["fingerprint"]
        assert j["success"] and fp > 0
        assert c.post("/weights_checker", json={"action": "snapshot"}).json()["success"]
        assert c.post("/weights_checker", json={"action": "compare"}).json()["success"]
        name, p = next(iter(eng.runner.model.named_parameters()))
        p.data.view(-1)[3] += 1.0
        r = c.post("/weights_checker", json={"action": "compare"})
        assert r.status_code == 400 and name in r.json()["message"]
        assert c.post("/slow_down", json={"forward_sleep_time": 0.05}).json()["success"]
        t0 = time.perf_counter()
        assert c.post("/generate", json={"input_ids": [5, 6, 7], "sampling_params": {"max_new_tokens": 3}}).status_code == 200
        assert time.perf_counter() - t0 > 0.15                    # prefill + decode steps each slept
        c.post("/slow_down", json={})
        assert c.post("/set_trace_level", json={"level": 2}).json() == {"success": True, "level": 2}
    # crash dump: in-flight + waiting requests land in a JSON file (the background loop is paused so it cannot finish the request first)
    eng.scheduler.paused = True
    eng.scheduler.handle_request(eng._tokenize(sgl.managers.io_struct.GenerateReqInput(input_ids=[1, 2, 3], sampling_params={"max_new_tokens": 4}).normalize()[0]))
    path = eng.dump_requests_before_crash()
    d = json.load(open(path))
    assert [r["input_ids"] for r in d["requests"]] == [[1, 2, 3]] and d["requests"][0]["state"] == "waiting" and "server_args" in d
    eng.scheduler.paused = False
    eng.shutdown()


def test_yaml_config_merge_and_plugins(tmp_path, monkeypatch):
    from sglang_b200 import plugins
    from sglang_b200.server_args import prepare_server_args

    cfg = tmp_path / "serve.yaml"
    cfg.write_text("model_path: dummy:tiny-llama\npage_size: 8\ndisable_radix_cache: true\ntp_size: 1\nlora_paths: [a, b]\n")
    a = prepare_server_args(["--config", str(cfg), "--page-size", "16", "--device", "cpu"])
    assert a.model_path == "dummy:tiny-llama" and a.page_size == 16 and a.disable_radix_cache and a.lora_paths == ["a", "b"]
    # a plugin module registers a model architecture and a hook
    (tmp_path / "my_sglb_plugin.py").write_text(
        "from sglang_b200.models.llama import LlamaForCausalLM\n"
        "seen = []\n"
        "def register(reg):\n"
        "    reg.register_model('MyLlamaForCausalLM', LlamaForCausalLM)\n"
        "    reg.add_hook('on_server_start', lambda **kw: seen.append(kw))\n")
    monkeypatch.syspath_prepend(str(tmp_path))
    monkeypatch.setenv("SGLB_PLUGINS", "my_sglb_plugin")
    reg = plugins.load_plugins(force=True)
    assert "my_sglb_plugin" in reg.loaded
    from sglang_b200.models.registry import get_model_class

    assert get_model_class("MyLlamaForCausalLM").__name__ == "LlamaForCausalLM"
    reg.run("on_server_start", port=1)
    assert sys.modules["my_sglb_plugin"].seen == [{"port": 1}]


def test_device_timer_cpu():
    from sglang_b200.utils.device_timer import DeviceTimer

    t = DeviceTimer("cpu")
    for _ in range(3):
        with t("mm"):
            torch.randn(64, 64) @ torch.randn(64, 64)
    s = t.summary()
    assert s["mm"]["calls"] == 3 and s["mm"]["ms"] > 0


def test_engine_watchdog_fires_on_a_stuck_forward(tmp_path, monkeypatch):
    """--watchdog-timeout is wired into the engine's background loop: a forward pass that hangs longer than the hard timeout produces the
    soft warning first, then the crash dump of the in-flight requests, and marks the engine wedged (servers additionally exit)"""
    import asyncio
    import os

    monkeypatch.setenv("SGLB_TEST_STUCK_SCHEDULER", "2.5")
    eng = sgl.Engine(model_path="dummy:tiny-llama", device="cpu", page_size=4, log_level="error", watchdog_timeout=1.0, soft_watchdog_timeout=0.3,
                     crash_dump_folder=str(tmp_path / "dumps"))
    try:
        async def go():
            return await eng.async_generate(input_ids=[5, 6, 7], sampling_params={"max_new_tokens": 2, "temperature": 0})

        out = asyncio.run(go())                       # the stuck forward eventually completes (2.5 s per pass)
        assert len(out["output_ids"]) == 2
        wd = eng._watchdog
        assert wd.fired_soft >= 1 and wd.fired_hard == 1 and eng.wedged
        dumps = os.listdir(tmp_path / "dumps")
        assert dumps and json.load(open(tmp_path / "dumps" / dumps[0]))["requests"][0]["input_ids"] == [5, 6, 7]
    finally:
        eng.shutdown()


def test_deterministic_inference_flag_pins_reduction_orders(monkeypatch):
    import os

    for k in ("SGLB_SKINNY", "SGLB_STREAMK_MAX_M", "SGLB_TP_LIBRARY_AR_MIN_BYTES"):
        monkeypatch.delenv(k, raising=False)
    from sglang_b200.server_args import ServerArgs

    ServerArgs(model_path="dummy:tiny-llama", enable_deterministic_inference=True)
    assert os.environ["SGLB_SKINNY"] == "off" and os.environ["SGLB_STREAMK_MAX_M"] == "0" and os.environ["SGLB_TP_LIBRARY_AR_MIN_BYTES"] == "0"
    eng = sgl.Engine(model_path="dummy:tiny-llama", device="cpu", page_size=4, log_level="error", enable_deterministic_inference=True)
    try:
        eng.runner.is_cuda = True                     # the split heuristic is a pure function of its arguments
        assert eng.runner._pick_splits(1, 1 << 16) == 1
    finally:
        eng.runner.is_cuda = False
        eng.shutdown()
