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


def test_more_tool_call_formats():
    import json

    from sglang_b200.function_call.function_call_parser import DETECTORS, FunctionCallParser

    tools = [{"type": "function", "function": {"name": "get_weather", "parameters": {"type": "object", "properties": {
        "city": {"type": "string"}, "days": {"type": "integer"}, "detail": {"type": "boolean"}}}}}]
    want = {"city": "Paris", "days": 3}
    samples = {
        "hermes": 'ok <tool_call>\n{"name": "get_weather", "arguments": {"city": "Paris", "days": 3}}\n</tool_call>',
        "internlm": 'hm<|action_start|><|plugin|>{"name": "get_weather", "parameters": {"city": "Paris", "days": 3}}<|action_end|>',
        "cohere_command4": '<|START_ACTION|>[{"tool_call_id": "0", "tool_name": "get_weather", "parameters": {"city": "Paris", "days": 3}}]<|END_ACTION|>',
        "lfm2": 'sure <|tool_call_start|>[get_weather(city="Paris", days=3)]<|tool_call_end|>',
        "deepseekv31": '<｜tool▁calls▁begin｜><｜tool▁call▁begin｜>get_weather<｜tool▁sep｜>{"city": "Paris", "days": 3}<｜tool▁call▁end｜><｜tool▁calls▁end｜>',
        "kimi_k2": '<|tool_calls_section_begin|><|tool_call_begin|>functions.get_weather:0<|tool_call_argument_begin|>{"city": "Paris", "days": 3}<|tool_call_end|><|tool_calls_section_end|>',
        "gpt-oss": '<|start|>assistant<|channel|>commentary to=functions.get_weather <|constrain|>json<|message|>{"city": "Paris", "days": 3}<|call|>',
        "qwen3_coder": '<tool_call>\n<function=get_weather>\n<parameter=city>\nParis\n</parameter>\n<parameter=days>\n3\n</parameter>\n</function>\n</tool_call>',
        "glm45": '<tool_call>get_weather\n<arg_key>city</arg_key>\n<arg_value>Paris</arg_value>\n<arg_key>days</arg_key>\n<arg_value>3</arg_value>\n</tool_call>',
        "minimax-m2": '<minimax:tool_call>\n<invoke name="get_weather">\n<parameter name="city">Paris</parameter>\n<parameter name="days">3</parameter>\n</invoke>\n</minimax:tool_call>',
        "deepseekv32": '<｜DSML｜function_calls>\n<｜DSML｜invoke name="get_weather">\n<｜DSML｜parameter name="city" string="true">Paris</｜DSML｜parameter>\n<｜DSML｜parameter name="days" string="false">3</｜DSML｜parameter>\n</｜DSML｜invoke>\n</｜DSML｜function_calls>',
        "step3": '<｜tool_calls_begin｜><｜tool_call_begin｜>function<｜tool_sep｜><steptml:invoke name="get_weather">\n<steptml:parameter name="city">Paris</steptml:parameter>\n<steptml:parameter name="days">3</steptml:parameter>\n</steptml:invoke><｜tool_call_end｜><｜tool_calls_end｜>',
        "dots": '<dots_function_call>\n<invoke name="get_weather">\n<parameter name="city">Paris</parameter>\n<parameter name="days">3</parameter>\n</invoke>\n</dots_function_call>',
        "minicpm5": '<function name="get_weather"><param name="city">Paris</param><param name="days">3</param></function>',
    }
    assert len(DETECTORS) >= 25
    for fmt, text in samples.items():
        p = FunctionCallParser(tools, fmt)
        assert p.has_tool_call(text), fmt
        normal, calls = p.parse_non_stream(text)
        assert len(calls) == 1 and calls[0].name == "get_weather", (fmt, calls)
        assert json.loads(calls[0].parameters) == want, (fmt, calls[0].parameters)
        assert "get_weather" not in normal, (fmt, normal)
        # streaming: feed char by char, exactly one call must come out and no marker text may leak
        sp = FunctionCallParser(tools, fmt)
        out_text, out_calls = "", []
        for ch in text:
            d, cs = sp.parse_stream_chunk(ch)
            out_text += d
            out_calls += cs
        assert len(out_calls) == 1 and json.loads(out_calls[0].parameters) == want, (fmt, out_calls)
        assert "get_weather" not in out_text, (fmt, out_text)
        st = p.get_structural_constraint("required")
        assert st["structures"][0]["begin"] and st["structures"][0]["schema"]["type"] == "object"
    # unknown function names are ignored
    p = FunctionCallParser(tools, "hermes")
    assert p.parse_non_stream('<tool_call>{"name": "nope", "arguments": {}}</tool_call>')[1] == []


def test_jit_build_cache_and_weak_ref(tmp_path, monkeypatch):
    import torch

    from sglang_b200 import jit

    monkeypatch.setenv("SGLB_JIT_CACHE", str(tmp_path))
    jit._LOADED.clear()
    so = jit.build("sglb_jit_weak_ref", [jit.CSRC_JIT / "weak_ref.cpp"])
    assert so.exists() and so.parent.parent == tmp_path
    t0 = so.stat().st_mtime
    assert jit.build("sglb_jit_weak_ref", [jit.CSRC_JIT / "weak_ref.cpp"]) == so and so.stat().st_mtime == t0      # cache hit, no rebuild
    x = torch.arange(12, dtype=torch.float32).view(3, 4)
    w = jit.weak_ref_tensor(x[:, 1:3])
    assert w.data_ptr() == x[:, 1:3].data_ptr() and w.shape == (3, 2) and w.stride() == (4, 1)
    x[0, 1] = 42.0
    assert w[0, 0] == 42.0
    # a different flag set is a different cache entry
    so2 = jit.build("sglb_jit_weak_ref", [jit.CSRC_JIT / "weak_ref.cpp"], extra_cflags=["-DSGLB_JIT_TEST_FLAG=1"])
    assert so2 != so and so2.exists()
    # inline source: a tiny CPU op
    src = '#include <torch/library.h>\n#include <torch/types.h>\nstatic at::Tensor twice(const at::Tensor& t) { return t * 2; }\n' \
          'TORCH_LIBRARY(sglb_jit_t, m) { m.def("twice(Tensor t) -> Tensor", &twice); }\n'
    jit.load_inline("sglb_jit_t", src, is_cuda=False)
    assert torch.equal(torch.ops.sglb_jit_t.twice(torch.ones(3)), torch.full((3,), 2.0))


def test_update_weights_from_distributed_group():
    """RL weight sync: an external 'trainer' (rank 0 of a private gloo group) broadcasts a tensor, the engine loads it"""
    import datetime
    import socket
    import threading

    import torch
    import torch.distributed as dist

    import sglang_b200 as sgl

    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    eng = sgl.Engine(model_path="dummy:tiny-llama", page_size=4, log_level="error")
    name = "model.layers.0.mlp.down_proj.weight"
    shape = list(dict(eng.runner.model.named_parameters())[name].shape)
    new = torch.full(shape, 0.125)
    before = eng.generate(input_ids=[5, 6, 7, 8], sampling_params={"max_new_tokens": 4, "temperature": 0})["output_ids"]
    err = []

    def trainer():
        try:
            store = dist.TCPStore("127.0.0.1", port, 2, is_master=True, timeout=datetime.timedelta(seconds=60))
            pg = dist.ProcessGroupGloo(dist.PrefixStore("weight_update_group", store), 0, 2, datetime.timedelta(seconds=60))
            opts = dist.BroadcastOptions(); opts.rootRank = 0
            pg.broadcast([new.clone()], opts).wait()
        except Exception as e:  # pragma: no cover
            err.append(e)

    th = threading.Thread(target=trainer); th.start()
    r = eng.init_weights_update_group("127.0.0.1", port, 1, 2, backend="gloo")
    assert r.success, r.message
    r = eng.update_weights_from_distributed([name], ["float32"], [shape])
    th.join(timeout=60)
    assert not err and r.success, (err, r.message)
    got = torch.tensor(eng.get_weights_by_name(name, 8))
    assert torch.allclose(got, torch.full((8,), 0.125))
    after = eng.generate(input_ids=[5, 6, 7, 8], sampling_params={"max_new_tokens": 4, "temperature": 0})
    assert after["meta_info"]["cached_tokens"] == 0          # the radix cache was flushed with the old weights
    assert eng.update_weights_from_distributed([name], ["float32"], [shape], group_name="nope").success is False
    assert eng.destroy_weights_update_group().success
    eng.shutdown()


def test_kv_indexer_and_kv_aware_routing():
    """C++ KvIndexer + KV-event-driven routing: the worker that really holds a request's prefix (per its radix-cache events) is chosen"""
    import numpy as np

    import sglang_b200 as sgl
    from sglang_b200 import sglb_rt
    from sglang_b200.mem_cache.radix_cache import chain_hashes
    from sglang_b200.router.router import Policy, Worker, make_engine_tokenizer

    idx = sglb_rt.KvIndexer()
    h = chain_hashes(np.arange(64, dtype=np.int32), 4, 0).tolist()
    idx.stored(0, h[:10]); idx.stored(1, h[:4]); idx.stored(1, [123456789])
    assert dict(idx.match(h)) == {0: 10, 1: 4} and idx.worker_blocks(1) == 5
    idx.removed(0, h[6:10])
    assert dict(idx.match(h)) == {0: 6, 1: 4}
    idx.clear_worker(1)
    assert dict(idx.match(h)) == {0: 6} and idx.worker_blocks(1) == 0 and idx.num_blocks() == 6
    assert idx.match([42]) == []
    # chained: the same page content under a different prefix hashes differently
    a = chain_hashes(np.asarray([1, 2, 3, 4, 9, 9, 9, 9], dtype=np.int32), 4, 0)
    b = chain_hashes(np.asarray([5, 6, 7, 8, 9, 9, 9, 9], dtype=np.int32), 4, 0)
    assert a[1] != b[1] and chain_hashes(np.asarray([1, 2, 3, 4], dtype=np.int32), 4, 0)[0] == a[0]

    # two engines with KV events on; worker 0 serves a long document, worker 1 something else
    engs = [sgl.Engine(model_path="dummy:tiny-llama", page_size=4, enable_kv_events=True, log_level="error") for _ in range(2)]
    pol = Policy("kv_aware")
    pol.attach_kv_index(make_engine_tokenizer("dummy:tiny-llama"), 4)
    workers = [Worker("http://w0", index=0), Worker("http://w1", index=1)]
    doc = "the quick brown fox jumps over the lazy dog " * 6
    sp = {"max_new_tokens": 2, "temperature": 0}
    engs[0].generate(doc + "question one", sp)
    engs[1].generate("completely different text that shares nothing with it " * 4, sp)
    for i, e in enumerate(engs):
        ev = e.take_kv_events()
        assert ev["page_size"] == 4 and any(k == "BlockStored" for k, _ in ev["events"])
        pol.apply_kv_events(i, ev["events"])
    assert pol._select_kv_aware(workers, doc + "question two") is workers[0]
    assert pol._select_kv_aware(workers, "completely different text that shares nothing with it " * 4 + "tail") is workers[1]
    workers[1].inflight = 3
    assert pol._select_kv_aware(workers, "never seen before prompt") is workers[0]         # no match -> least loaded
    # eviction / flush events remove the prefix from the index
    engs[0].flush_cache()
    pol.apply_kv_events(0, engs[0].take_kv_events()["events"])
    workers[1].inflight = 0; workers[0].inflight = 2
    assert pol._select_kv_aware(workers, doc + "question three") is workers[1]
    for e in engs:
        e.shutdown()


def test_remote_instance_loader_clones_a_running_server():
    """load_format=remote_instance: a new engine pulls every parameter from a running seed instance over a private 2-rank group (no disk)"""
    import socket
    import threading

    import torch
    import uvicorn

    import sglang_b200 as sgl
    from sglang_b200.entrypoints.http_server import create_app
    from sglang_b200.lang.backend.runtime_endpoint import _wait_http

    def free_port():
        s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
        return p

    seed = sgl.Engine(model_path="dummy:tiny-llama", device="cpu", page_size=4, random_seed=11, log_level="error")
    seed.start_background_loop()
    port = free_port()
    server = uvicorn.Server(uvicorn.Config(create_app(seed), host="127.0.0.1", port=port, log_level="error"))
    t = threading.Thread(target=server.run, daemon=True)
    t.start()
    try:
        _wait_http(f"http://127.0.0.1:{port}", 30)
        new = sgl.Engine(model_path="dummy:tiny-llama", device="cpu", page_size=4, random_seed=0, log_level="error", load_format="remote_instance",
                         remote_instance_weight_loader_seed_instance_ip="127.0.0.1", remote_instance_weight_loader_seed_instance_service_port=port,
                         remote_instance_weight_loader_send_weights_group_port=free_port())
        a, b = dict(seed.runner.model.named_parameters()), dict(new.runner.model.named_parameters())
        assert a.keys() == b.keys() and all(torch.equal(a[k], b[k]) for k in a)
        sp = {"max_new_tokens": 5, "temperature": 0, "ignore_eos": True}
        assert new.generate(input_ids=[5, 6, 7, 8], sampling_params=sp)["output_ids"] == seed.generate(input_ids=[5, 6, 7, 8], sampling_params=sp)["output_ids"]
        new.shutdown()
    finally:
        server.should_exit = True
        t.join(timeout=10)
        seed.shutdown()


def test_reasoner_grammar_constrains_only_after_think_end():
    import numpy as np

    from sglang_b200.constrained.base_grammar import BaseGrammar
    from sglang_b200.constrained.reasoner_grammar import ReasonerGrammar

    class OnlyToken7(BaseGrammar):
        def __init__(self):
            self.seen = []

        def accept_token(self, t):
            assert t == 7
            self.seen.append(t)

        def fill_vocab_mask(self, row, V):
            row[:] = 0
            row[0] = 1 << 7

        def is_terminated(self):
            return len(self.seen) >= 2

        def copy(self):
            c = OnlyToken7()
            c.seen = list(self.seen)
            return c

    g = ReasonerGrammar(OnlyToken7(), end_ids=[50, 51])
    row = np.zeros(2, dtype=np.int32)
    g.fill_vocab_mask(row, 64)
    assert (row == -1).all()                                  # thinking: everything allowed
    for t in (3, 50, 9, 50, 51):                              # "50 9" is not the marker; "50 51" is
        g.accept_token(t)
    assert not g.in_reasoning
    g.fill_vocab_mask(row, 64)
    assert row[0] == 1 << 7 and row[1] == 0                   # answering: the inner grammar rules
    c = g.copy()
    g.accept_token(7)
    g.accept_token(7)
    assert g.is_terminated() and not c.is_terminated()


def test_collective_rpc_and_batch_tokenizer():
    import asyncio

    from sglang_b200.utils.batch_tokenizer import AsyncDynamicBatchTokenizer

    eng = sgl.Engine(model_path="dummy:tiny-llama", device="cpu", page_size=4, log_level="error")
    assert eng.collective_rpc("get_memory_info")["num_pages"] > 0          # a model-runner method
    assert eng.collective_rpc("has_work") is False                          # a scheduler method
    with pytest.raises(Exception):
        r = eng._control(sgl.managers.io_struct.RpcReq("_private", {}))
        assert not r.success
        raise RuntimeError(r.message)
    tk = eng.tokenizer

    async def run():
        bt = AsyncDynamicBatchTokenizer(tk, max_batch_size=8, batch_wait_timeout_s=0.02)
        outs = await asyncio.gather(*[bt.encode(f"hello {i}") for i in range(10)])
        return outs, bt.batches

    outs, batches = asyncio.run(run())
    assert outs == [tk.encode(f"hello {i}") for i in range(10)]
    assert sum(batches) == 10 and max(batches) > 1                          # concurrent requests were coalesced
    eng.shutdown()


def test_grammar_manager_wraps_only_requests_that_will_reason():
    from types import SimpleNamespace

    from sglang_b200.constrained.grammar_manager import GrammarManager
    from sglang_b200.constrained.reasoner_grammar import ReasonerGrammar

    class G:
        def copy(self):
            return G()

    old = getattr(GrammarManager, "think_end", None)
    GrammarManager.think_end = [42]
    try:
        plain = SimpleNamespace(sampling_params=SimpleNamespace(custom_params=None), grammar=None, to_abort=None)
        think = SimpleNamespace(sampling_params=SimpleNamespace(custom_params={"require_reasoning": True}), grammar=None, to_abort=None)
        GrammarManager._bind(plain, G())
        GrammarManager._bind(think, G())
        assert isinstance(plain.grammar, G) and isinstance(think.grammar, ReasonerGrammar)
    finally:
        GrammarManager.think_end = old


def test_expert_placement_planner_replicates_hot_experts_and_skips_dead_ranks():
    import numpy as np

    from sglang_b200.eplb.placement import ExpertPlacement, imbalance, plan

    load = np.asarray([900.0, 10, 10, 10, 10, 10, 10, 40])
    base = ExpertPlacement.default(8, 4)
    pl = plan(load, ep_size=4, slots=3, live=None)                     # 12 slots for 8 experts: 4 replicas to hand out
    reps = pl.replicas(8)
    assert len(reps[0]) >= 3 and all(len(r) >= 1 for r in reps)         # the hot expert is replicated, nobody is dropped
    assert len({p // 3 for p in reps[0]}) == len(reps[0])               # its replicas sit on different ranks
    assert imbalance(load, pl) < 0.6 * imbalance(load, base)
    dead = plan(load, ep_size=4, slots=3, live=[0, 2, 3])               # rank 1 lost: 9 slots still hold all 8 experts
    assert (dead.phys_to_logical[1] == -1).all() and all(dead.replicas(8))
    tab, cnt = dead.tables(8, "cpu")
    assert tab.shape[0] == 8 and int(cnt.min()) == 1 and not any(int(p) // 3 == 1 for p in tab.flatten())
    import pytest

    with pytest.raises(ValueError):
        plan(load, ep_size=4, slots=2, live=[0])                        # 2 slots cannot hold 8 experts


def test_tracing_spans_and_otlp_http_export():
    """request spans (root + queue / prefill / decode), scheduler-step spans at level 3, and the OTLP/HTTP JSON exporter posting well-formed
    ExportTraceServiceRequest bodies to a collector (observability/trace.py; reference observability/trace.py:211,313)"""
    import http.server
    import json
    import threading
    import time

    import sglang_b200 as sgl
    from sglang_b200.observability import trace as T

    got = []

    class H(http.server.BaseHTTPRequestHandler):
        def do_POST(self):
            body = self.rfile.read(int(self.headers["Content-Length"]))
            got.append((self.path, self.headers.get("Content-Type"), json.loads(body)))
            self.send_response(200); self.end_headers(); self.wfile.write(b"{}")

        def log_message(self, *a):
            pass

    srv = http.server.HTTPServer(("127.0.0.1", 0), H)
    threading.Thread(target=srv.serve_forever, daemon=True).start()
    T._TRACER = None
    eng = sgl.Engine(model_path="dummy:tiny-llama", device="cpu", log_level="error", enable_trace=True, trace_level=3,
                     otlp_traces_endpoint=f"http://127.0.0.1:{srv.server_address[1]}")
    try:
        eng.generate(["hello there", "another prompt"], {"max_new_tokens": 4, "temperature": 0})
        tr = T.get_tracer()
        names = [d["name"] for d in tr.ring]
        assert names.count("request") == 2 and {"queue", "prefill", "decode"} <= set(names) and "scheduler_step" in names
        root = next(d for d in tr.ring if d["name"] == "request")
        assert root["attrs"]["completion_tokens"] == 4 and root["end"] >= root["start"]
        tr.otlp.flush()
        t0 = time.time()
        while not got and time.time() - t0 < 5:
            time.sleep(0.05)
        path, ctype, body = got[0]
        assert path == "/v1/traces" and ctype == "application/json"
        spans = body["resourceSpans"][0]["scopeSpans"][0]["spans"]
        assert body["resourceSpans"][0]["resource"]["attributes"][0]["value"]["stringValue"] == "sglang_b200"
        assert all(len(s["traceId"]) == 32 and len(s["spanId"]) == 16 and int(s["endTimeUnixNano"]) >= int(s["startTimeUnixNano"]) for s in spans)
        assert tr.otlp.sent >= len(spans) and tr.otlp.failed == 0
    finally:
        eng.shutdown()
        if T.get_tracer() is not None and T.get_tracer().otlp is not None:
            T.get_tracer().otlp.shutdown()
        T._TRACER = None
        srv.shutdown()


def test_reasoning_parser_families_stream_equals_non_stream():
    """every reasoning family: the streamed deltas (any chunking, markers split across chunks) concatenate to the non-streaming split"""
    import random

    from sglang_b200.parser.reasoning_parser import _FAMILIES, ReasoningParser

    samples = {
        "deepseek-r1": "let me think\nabout it</think>\nThe answer is 4.",
        "qwen3": "<think>step 1\nstep 2</think>\n\nDone.",
        "qwen3-thinking": "no opener here</think>Final.",
        "kimi": "◁think▷hmm◁/think▷ok",
        "kimi-k2": "<think>need a tool<|tool_calls_section_begin|>functions.f:0<|tool_call_argument_begin|>{}",
        "glm45": "<think>plan</think>\n<tool_call>get_weather\n</tool_call>",
        "mistral": "[THINK]chain[/THINK]Answer",
        "hunyuan": "<think>t</think>\n<answer>42</answer>",
        "seed-oss": "<seed:think>budget</seed:think>hi",
        "minimax-append-think": "implicit start</think>text",
        "gpt-oss": "<|channel|>analysis<|message|>User wants 2+2.<|end|><|start|>assistant<|channel|>final<|message|>4<|return|>",
    }
    rng = random.Random(0)
    for fam in _FAMILIES:
        text = samples.get(fam, "<think>x y z</think>answer")
        want_r, want_c = ReasoningParser(fam).parse_non_stream(text)
        for _ in range(20):
            p = ReasoningParser(fam)
            r, c, i = "", "", 0
            while i < len(text):
                n = rng.randint(1, 7)
                dr, dc = p.parse_stream_chunk(text[i:i + n])
                r, c, i = r + dr, c + dc, i + n
            assert (r.strip() or None) == (want_r or None) and c.strip() == want_c.strip(), (fam, r, c, want_r, want_c)
    r, c = ReasoningParser("gpt-oss").parse_non_stream(samples["gpt-oss"])
    assert r == "User wants 2+2." and c == "4"
    r, c = ReasoningParser("gpt-oss").parse_non_stream("<|channel|>analysis<|message|>call it<|end|><|start|>assistant<|channel|>commentary to=functions.f "
                                                       '<|constrain|>json<|message|>{"a":1}<|call|>')
    assert r == "call it" and "to=functions.f" in c and c.endswith("<|call|>")
    r, c = ReasoningParser("hunyuan").parse_non_stream(samples["hunyuan"])
    assert r == "t" and c.strip() == "42"
    r, c = ReasoningParser("kimi-k2").parse_non_stream(samples["kimi-k2"])
    assert r == "need a tool" and c.startswith("<|tool_calls_section_begin|>")


def test_expert_recorder_modes_and_eplb_manager():
    """device-side accumulation, windowed load, balancedness, and the automatic re-plan trigger (eplb/expert_distribution.py)"""
    import numpy as np
    import torch

    from sglang_b200.eplb.expert_distribution import EPLBManager, ExpertDistributionRecorder

    rec = ExpertDistributionRecorder(2, 8, mode="window", window=3)
    rec.enabled = True
    skew = torch.tensor([[0, 0], [0, 1], [0, 0], [0, 2]])                 # expert 0 is hot
    for i in range(5):
        rec.record(0, skew)
        rec.record(1, torch.tensor([[i % 8, (i + 1) % 8]]))
        rec.end_pass()
    load = rec.load()
    assert load.shape == (2, 8) and load[0, 0] == 3 * 6 and load[0].sum() == 3 * 8     # only the last 3 passes count
    assert rec.dev_counts[0, 0].item() == 5 * 6                                          # the cumulative tensor keeps everything
    bal = rec.balancedness(load, 4)
    assert bal[0] < 0.5 and 0 < bal[1] <= 1.0
    calls = []

    class R:
        expert_placement = None

        def eplb_rebalance(self, redundant, load):
            calls.append((redundant, load.copy()))

    rec2 = ExpertDistributionRecorder(1, 8, mode="window", window=4)
    mgr = EPLBManager(R(), rec2, ep_size=4, interval=4, threshold=0.8)
    for i in range(8):
        rec2.record(0, torch.tensor([[0, 0, 0, 1]]))
        mgr.on_forward_pass()
    assert mgr.rebalances == 2 and len(calls) == 2 and calls[0][1][0, 0] == 12 and mgr.last_balancedness < 0.8
    even = ExpertDistributionRecorder(1, 8, mode="window", window=4)
    mgr2 = EPLBManager(R(), even, ep_size=4, interval=4, threshold=0.8)
    for i in range(8):
        even.record(0, torch.arange(8).view(1, 8))
        mgr2.on_forward_pass()
    assert mgr2.rebalances == 0 and abs(mgr2.last_balancedness - 1.0) < 1e-9


def test_async_dynamic_batch_tokenizer_batches_concurrent_prompts():
    """prompts arriving together are encoded by one batched call on a worker thread; results equal plain encode; failures reach every waiter"""
    import asyncio
    import threading

    from sglang_b200.managers.async_batch_tokenizer import AsyncDynamicBatchTokenizer

    class Tok:
        def __init__(self):
            self.calls, self.threads = [], set()

        def encode(self, t):
            return [ord(c) for c in t]

        def __call__(self, texts):
            self.calls.append(len(texts))
            self.threads.add(threading.current_thread().name)
            if any(t == "boom" for t in texts):
                raise RuntimeError("bad batch")
            return {"input_ids": [self.encode(t) for t in texts]}

    tok = Tok()
    dt = AsyncDynamicBatchTokenizer(tok, max_batch_size=8, batch_wait_timeout_s=0.02)

    async def go():
        texts = [f"prompt {i}" for i in range(20)]
        out = await asyncio.gather(*[dt.encode(t) for t in texts])
        assert out == [tok.encode(t) for t in texts]
        assert max(tok.calls) == 8 and sum(tok.calls) >= 19 and dt.batches <= 4              # 20 prompts -> 3 batched calls, not 20
        assert all(n.startswith("sglb-tokenize") for n in tok.threads)                       # never on the event loop thread
        single = await dt.encode("solo")
        assert single == tok.encode("solo")

    asyncio.run(go())
