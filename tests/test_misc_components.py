"""Small components: CLI, IO-struct codec, server args, EPLB recorder/rebalancer, watchdog, tracing, router policies."""
import json
import time

import numpy as np
import pytest

import sglang_b200 as sgl
from sglang_b200.managers import io_struct as ios
from sglang_b200.server_args import ServerArgs, prepare_server_args


def test_io_struct_roundtrip():
    m = ios.TokenizedGenerateReqInput(rid="a", input_text="hi", input_ids=[1, 2, 3], sampling_params={"temperature": 0.5}, stream=True)
    back = ios.decode(ios.encode(m))
    assert back == m
    out = ios.BatchTokenIDOutput(rids=["a"], finished_reasons=[{"type": "stop", "matched": 2}], output_ids=[[5, 6]], read_offsets=[0],
                                 prompt_tokens=[3], completion_tokens=[2], cached_tokens=[0], skip_special_tokens=[True],
                                 no_stop_trim=[False], stop_strs=[[]])
    assert ios.decode(ios.encode(out)) == out
    g = ios.GenerateReqInput(text=["a", "b"], sampling_params={"n": 2, "max_new_tokens": 4})
    assert len(g.normalize()) == 4
    with pytest.raises(ValueError):
        ios.GenerateReqInput(text="a", input_ids=[1]).normalize()


def test_server_args_cli_and_validation(tmp_path):
    a = prepare_server_args(["--model-path", "dummy:tiny-llama", "--tp", "2", "--page-size", "16", "--schedule-policy", "lpm",
                             "--disable-radix-cache"])
    assert a.tp_size == 2 and a.page_size == 16 and a.schedule_policy == "fcfs"   # lpm needs the radix cache
    cfg = tmp_path / "c.yaml"
    cfg.write_text("page-size: 8\nmax-running-requests: 7\n")
    b = prepare_server_args(["--model-path", "dummy:tiny-llama", "--config", str(cfg)])
    assert b.page_size == 8 and b.max_running_requests == 7
    with pytest.raises(ValueError):
        ServerArgs(page_size=3)
    with pytest.raises(ValueError):
        ServerArgs(tp_size=4, ep_size=3)


def test_eplb_recorder_and_rebalance():
    eng = sgl.Engine(model_path="dummy:tiny-mixtral", page_size=4)
    assert eng._control(ios.ExpertDistributionReq("start")).success
    eng.generate(["a b c d e f g", "hello there"], {"max_new_tokens": 6})
    d = eng._control(ios.ExpertDistributionReq("dump")).payload
    counts = np.asarray(d["counts"])
    assert counts.shape == (2, 4) and counts.sum() > 0
    from sglang_b200.eplb.expert_distribution import rebalance

    skew = np.array([[100, 1, 1, 100], [5, 5, 5, 5]])
    perm = rebalance(skew, ep_size=2)
    loads = skew[0][perm[0]].reshape(2, 2).sum(1)
    assert abs(loads[0] - loads[1]) <= 1                       # the two hot experts land on different ranks


def test_watchdog_and_tracing():
    from sglang_b200.environ import envs
    from sglang_b200.observability.trace import init_tracing
    from sglang_b200.utils.watchdog import SchedulerWatchdog

    eng = sgl.Engine(model_path="dummy:tiny-llama", page_size=4)
    fired = []
    wd = SchedulerWatchdog(eng.scheduler, hard_timeout=0.6, soft_timeout=0.2, interval=0.05, on_hard=lambda: fired.append(1)).start()
    with envs.SGLB_TEST_STUCK_SCHEDULER.override(1.0):
        eng.generate("x", {"max_new_tokens": 2})
    time.sleep(0.2)
    wd.stop()
    assert fired and wd.fired_soft >= 1
    tr = init_tracing(level=2)
    out = eng.generate("trace me", {"max_new_tokens": 3})
    tr.trace_request(out["meta_info"]["id"], out["meta_info"], time.time() - out["meta_info"]["e2e_latency"])
    names = [s["name"] for s in tr.ring]
    assert names[0] == "request" and {"queue", "prefill", "decode"} <= set(names)


def test_router_policies():
    from sglang_b200.router.router import Policy, Worker

    ws = [Worker("http://a"), Worker("http://b"), Worker("http://c")]
    p = Policy("cache_aware")
    first = p.select(ws, "SYSTEM: you are helpful. USER: question one")
    again = p.select(ws, "SYSTEM: you are helpful. USER: question one and more")
    assert first is again                                         # shared prefix -> same worker
    rr = Policy("round_robin")
    assert len({rr.select(ws, "").url for _ in range(3)}) == 3
    ch = Policy("consistent_hashing")
    assert ch.select(ws, "x", "user-42") is ch.select(ws, "y", "user-42")
    ws[0].inflight = 50
    assert Policy("power_of_two").select(ws[:2], "").url == "http://b"
