"""SPMD worker for the multi-process CPU tests (gloo): every rank builds the same Engine and drives the same requests; rank 0
prints the result as one JSON line.   env: RANK WORLD_SIZE MASTER_ADDR MASTER_PORT;  argv: json config"""
import faulthandler
import json
import os
import sys

if os.environ.get("SGLB_HANG_DUMP_S"):
    faulthandler.dump_traceback_later(int(os.environ["SGLB_HANG_DUMP_S"]), exit=True)

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
cfg = json.loads(sys.argv[1])
import sglang_b200 as sgl  # noqa: E402

eng = sgl.Engine(log_level="error", device="cpu", **cfg["engine"])
outs = eng.generate(input_ids=cfg["prompts"], sampling_params=cfg["sampling"], **cfg.get("gen_kw", {}))
rpc_out = []
for step in cfg.get("rpc", []):              # control actions between two generations (every rank issues them: SPMD)
    if step.get("zero_rank") is not None:    # fault injection: wipe a "dead" rank's expert memory so nothing can silently read it
        if int(os.environ.get("RANK", 0)) == step["zero_rank"]:
            for m in eng.runner.model.modules():
                if m.__class__.__name__ == "FusedMoE":
                    m.w13.data.zero_(), m.w2.data.zero_()
        continue
    rpc_out.append(eng.collective_rpc(step["method"], **step.get("kwargs", {})))
outs2 = eng.generate(input_ids=cfg["prompts"], sampling_params=cfg["sampling"], **cfg.get("gen_kw", {})) if cfg.get("rpc") else None
if int(os.environ.get("RANK", 0)) == 0:      # only the driver rank receives outputs
    res = {"ids": [o["output_ids"] for o in outs]}
    if outs2 is not None:
        res["ids2"], res["rpc"] = [o["output_ids"] for o in outs2], rpc_out
    if cfg.get("gen_kw", {}).get("return_logprob"):
        res["lp"] = [[float(t[0]) for t in o["meta_info"]["input_token_logprobs"] if t[0] is not None] for o in outs]
    if getattr(eng.runner, "dwdp", None) is not None:
        m = eng.runner.dwdp
        res["dwdp"] = {"forwards": m.num_forwards, "prefetches": m.num_prefetches, "bytes": m.bytes_moved,
                       "a2a_calls": sum(getattr(l, "_a2a_calls", 0) for l in m.layers)}
    print("RESULT " + json.dumps(res), flush=True)
if os.environ.get("SGLB_TEST_GREEDY_SHARD"):
    import torch
    from sglang_b200.layers.linear import vocab_parallel_greedy

    r = eng.runner
    head = r.greedy_head()
    assert head is not None, "vocab-parallel greedy head expected at TP > 1"
    g = torch.Generator().manual_seed(7)
    hid = torch.randn(5, head.weight.shape[1], generator=g).to(head.weight.dtype)
    head.weight.data[:, :] = torch.randn(head.weight.shape, generator=torch.Generator().manual_seed(11 + r.tp.rank)).to(head.weight.dtype) * 0.3
    want = r.model.compute_logits(hid).float().argmax(dim=-1).to(torch.int32)
    buf = torch.zeros(r.tp.world_size, 6, 4, dtype=torch.bfloat16)
    nt = torch.zeros(16, dtype=torch.int32)
    slots = torch.tensor([3, 1, 4, 9, 2], dtype=torch.int32)
    got = vocab_parallel_greedy(head, hid, buf, slots, nt, torch.zeros(5, dtype=torch.int32))
    assert torch.equal(got, want), (got, want)
    assert torch.equal(nt[slots.long()], want)
if os.environ.get("SGLB_TEST_EXPECT_A2A"):
    ex = eng.runner.model.model.layers[0].mlp.experts
    assert getattr(ex, "_a2a_calls", 0) > 0, "the all-to-all dispatcher was expected to run"
if os.environ.get("SGLB_TEST_EXPECT_DCP"):
    assert eng.runner.dcp is not None and eng.runner.kv_pool.page_size * eng.runner.dcp.world_size == eng.runner.page_size
    assert eng.runner.dcp.comm_calls > 0, "DCP collectives were expected to run"
if os.environ.get("SGLB_TEST_EXPECT_CP"):
    assert eng.runner.cp is not None and eng.runner.cp_forwards > 0, "context-parallel prefills were expected to run"
if os.environ.get("SGLB_TEST_EXPECT_SP"):
    from sglang_b200.layers.sp import sp_context

    assert sp_context().forwards > 0, "sequence-parallel prefills were expected to run"
eng.shutdown()
os._exit(0)
