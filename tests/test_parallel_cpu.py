"""Tensor / pipeline / expert parallelism on the CPU tier: world_size 2 over gloo, one process per rank, against the single-process
engine on the same synthetic checkpoint (reference: test/registered/distributed/* run on 2-8 GPUs; the host-side logic — rank
layout, sharded weight loading, SPMD scheduling, PP send/recv, EP expert ownership — is identical on gloo)."""
import json
import os
import socket
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
PROMPTS = [[5, 6, 7, 8, 9, 10, 11, 12, 13], [3, 4, 5], list(range(20, 75)), [7] * 30]
SAMPLING = {"max_new_tokens": 8, "temperature": 0, "ignore_eos": True}


def _free_port() -> int:
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run(engine_kw, world, gen_kw=None, extra_env=None, prompts=None, rpc=None):
    cfg = json.dumps({"engine": engine_kw, "prompts": prompts or PROMPTS, "sampling": SAMPLING, "gen_kw": gen_kw or {}, "rpc": rpc or []})
    port = _free_port()
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   PYTHONPATH=ROOT, OMP_NUM_THREADS="2", **(extra_env or {}))
        procs.append(subprocess.Popen([sys.executable, os.path.join(HERE, "_mp_worker.py"), cfg], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.PIPE, text=True, cwd=ROOT))
    outs = []
    try:
        for p in procs:
            o, e = p.communicate(timeout=240)
            outs.append((p.returncode, o, e))
    finally:
        for p in procs:
            if p.poll() is None:
                p.kill()
    for rc, o, e in outs:
        assert rc == 0, e[-3000:]
    line = [l for l in outs[0][1].splitlines() if l.startswith("RESULT ")][-1]
    return json.loads(line[7:])


def _check(par, ref, atol=2e-3):
    assert par["ids"] == ref["ids"]
    if "lp" in ref:
        for a, b in zip(par["lp"], ref["lp"]):
            assert len(a) == len(b) and max(abs(x - y) for x, y in zip(a, b)) < atol


LP = {"return_logprob": True, "logprob_start_len": 0}


def test_tensor_parallel_2_matches_single():
    kw = dict(model_path="dummy:tiny-llama", page_size=4, chunked_prefill_size=32)
    _check(_run({**kw, "tp_size": 2}, 2, LP), _run(kw, 1, LP))


def test_pipeline_parallel_2_matches_single():
    kw = dict(model_path="dummy:tiny-llama", page_size=4, chunked_prefill_size=32)
    _check(_run({**kw, "pp_size": 2}, 2, LP), _run(kw, 1, LP))


def test_expert_parallel_2_matches_single():
    kw = dict(model_path="dummy:tiny-mixtral", page_size=4)
    ref = _run(kw, 1, LP)
    _check(_run({**kw, "tp_size": 2, "ep_size": 2}, 2, LP), ref)     # whole experts per rank
    _check(_run({**kw, "tp_size": 2}, 2, LP), ref)                    # experts' intermediate dim split (moe-TP)


def test_tp2_qwen_bias_and_gqa_replication():
    # qkv bias sharding + kv-head replication (num_kv_heads=1 < tp): tiny-gemma has a single KV head
    for preset in ("tiny-qwen2", "tiny-gemma"):
        kw = dict(model_path=f"dummy:{preset}", page_size=4)
        _check(_run({**kw, "tp_size": 2}, 2, LP), _run(kw, 1, LP))


def test_world4_layouts():
    """TP=4 with KV-head replication, EP=4 / EP=2 x moe-TP=2, and TP=2 x PP=2 on four gloo ranks"""
    kw = dict(model_path="dummy:tiny-llama", page_size=4, chunked_prefill_size=32)
    ref = _run(kw, 1, LP)
    _check(_run({**kw, "tp_size": 4}, 4, LP), ref)
    _check(_run({**kw, "tp_size": 2, "pp_size": 2}, 4, LP), ref)
    kw = dict(model_path="dummy:tiny-mixtral", page_size=4)
    ref = _run(kw, 1, LP)
    _check(_run({**kw, "tp_size": 4, "ep_size": 4}, 4, LP), ref)
    _check(_run({**kw, "tp_size": 4, "ep_size": 2}, 4, LP), ref)


def test_shm_broadcast_ring_multi_reader():
    import pickle

    from sglang_b200.parallel.shm_broadcast import ShmBroadcaster

    w = ShmBroadcaster(2, max_chunk_bytes=4096, n_chunks=4)
    h = w.handle()
    arg = pickle.dumps((h.name, h.n_readers, h.max_chunk_bytes, h.n_chunks)).hex()
    procs = [subprocess.Popen([sys.executable, os.path.join(HERE, "_shm_reader.py"), arg, str(r)], stdout=subprocess.PIPE, text=True,
                              env=dict(os.environ, PYTHONPATH=ROOT)) for r in range(2)]
    want = 0
    for i in range(300):                      # far more messages than chunks; some span several chunks
        blob = b"x" * (10 if i % 7 else 20000)
        o = {"x": list(range(i % 50)), "blob": blob}
        want += sum(o["x"]) + len(blob)
        w.broadcast(o, timeout=30)
    w.broadcast(None, timeout=30)
    for p in procs:
        out, _ = p.communicate(timeout=60)
        assert p.returncode == 0 and out.strip() == f"SUM {want}", out
    w.close()


@pytest.mark.parametrize("shm", ["1", "0"])
def test_tp2_server_mode_background_loop(shm):
    """request fan-out to the peer rank over the shm ring (and over Gloo when disabled) gives the single-process results"""
    port = _free_port()
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   PYTHONPATH=ROOT, OMP_NUM_THREADS="2", SGLB_SHM_BROADCAST=shm, PEER_LIFETIME="25")
        procs.append(subprocess.Popen([sys.executable, os.path.join(HERE, "_mp_server_worker.py")], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.PIPE, text=True, cwd=ROOT))
    o, e = procs[0].communicate(timeout=200)
    assert procs[0].returncode == 0, e[-3000:]
    got = json.loads([l for l in o.splitlines() if l.startswith("RESULT ")][-1][7:])
    procs[1].kill()
    import sglang_b200 as sgl

    eng = sgl.Engine(model_path="dummy:tiny-llama", page_size=4, log_level="error")
    sp = {"max_new_tokens": 6, "temperature": 0, "ignore_eos": True}
    ref = [eng.generate(input_ids=p, sampling_params=sp)["output_ids"] for p in ([5, 6, 7, 8, 9], [3, 4], list(range(20, 50)), [9, 8, 7], [1, 2, 3, 4])]
    eng.shutdown()
    assert got == ref


def test_deepseek_mla_tp2_ep2():
    kw = dict(model_path="dummy:tiny-deepseek-v3", page_size=4)
    ref = _run(kw, 1, LP)
    _check(_run({**kw, "tp_size": 2}, 2, LP), ref, atol=5e-3)
    _check(_run({**kw, "tp_size": 2, "ep_size": 2}, 2, LP), ref, atol=5e-3)


def test_dp_attention_tp2_dp2_and_tp4_dp2():
    """DP-attention: requests are split between attention groups, MLP runs on the gathered tokens over the full TP group; results must
    equal the single-process engine (more requests than groups, uneven lengths -> idle batches at the tail)"""
    kw = dict(model_path="dummy:tiny-llama", page_size=4, chunked_prefill_size=32)
    ref = _run(kw, 1, LP)
    _check(_run({**kw, "tp_size": 2, "dp_size": 2, "enable_dp_attention": True}, 2, LP), ref)
    _check(_run({**kw, "tp_size": 4, "dp_size": 2, "enable_dp_attention": True}, 4, LP), ref)
    kw = dict(model_path="dummy:tiny-mixtral", page_size=4)
    ref = _run(kw, 1, LP)
    dpk = {**kw, "tp_size": 2, "dp_size": 2, "ep_size": 2, "enable_dp_attention": True}
    _check(_run(dpk, 2, LP, {"SGLB_MOE_A2A": "1", "SGLB_TEST_EXPECT_A2A": "1"}), ref)      # all-to-all expert dispatch / combine
    _check(_run(dpk, 2, LP, {"SGLB_MOE_A2A": "0"}), ref)                                   # gather-all-tokens dispatcher


def test_classic_family_tp2():
    for preset in ("tiny-gpt2", "tiny-neox"):
        kw = dict(model_path=f"dummy:{preset}", page_size=4)
        _check(_run({**kw, "tp_size": 2}, 2, LP), _run(kw, 1, LP))


def test_dp_attention_with_prefill_delayer():
    """the delayer's per-pass gather must stay matched across DP groups (uneven load -> 'mixed' passes are delayed, bounded)"""
    kw = dict(model_path="dummy:tiny-llama", page_size=4)
    ref = _run(kw, 1, LP)
    _check(_run({**kw, "tp_size": 2, "dp_size": 2, "enable_dp_attention": True, "enable_prefill_delayer": True,
                 "prefill_delayer_max_delay_passes": 3}, 2, LP), ref)


def test_decode_context_parallel_matches_single():
    """DCP: the two ranks that would hold copies of the single KV head split the tokens instead (token p on rank p % 2); prefill runs as
    per-residue virtual requests, decode on the local shard, partials merged by LSE -> must equal the single-process engine, with chunked
    prefill, prefix reuse (second wave shares a prefix) and sequences shorter than the group"""
    over = json.dumps({"num_key_value_heads": 1})
    kw = dict(model_path="dummy:tiny-llama", page_size=4, json_model_override_args=over, chunked_prefill_size=16)
    ref = _run(kw, 1, LP)
    _check(_run({**kw, "tp_size": 2, "dcp_size": 2}, 2, LP, {"SGLB_TEST_EXPECT_DCP": "1"}), ref)
    _check(_run({**kw, "tp_size": 4, "dcp_size": 2, "page_size": 8}, 4, LP), _run({**kw, "page_size": 8}, 1, LP))
    _check(_run({**kw, "tp_size": 4, "dcp_size": 4}, 4, LP), ref)


def test_prefill_context_parallel_matches_single():
    """prefill CP: every request's new rows are zig-zag split between the attention groups (K/V all-gathered per layer), decode runs the
    plain TP path on the replicated KV; chunked prefill + uneven tiny requests exercise empty blocks"""
    kw = dict(model_path="dummy:tiny-llama", page_size=4, chunked_prefill_size=32)
    ref = _run(kw, 1, LP)
    env = {"SGLB_CP_MIN_TOKENS": "1", "SGLB_TEST_EXPECT_CP": "1"}
    _check(_run({**kw, "tp_size": 2, "cp_size": 2}, 2, LP, env), ref)
    _check(_run({**kw, "tp_size": 4, "cp_size": 2}, 4, LP, env), ref)
    kw = dict(model_path="dummy:tiny-mixtral", page_size=4)
    _check(_run({**kw, "tp_size": 2, "cp_size": 2, "ep_size": 2}, 2, LP, env), _run(kw, 1, LP))


def test_hybrid_mamba_tp2():
    """Mamba-2 heads / B-C groups / conv channels sharded over TP (fused in_proj and conv tensors sliced per block), attention layers of the
    hybrid on the paged pool; single replicated group + full-width gated norm (all-reduced variance) for the pure Mamba-2 preset"""
    for preset in ("tiny-nemotron-h", "tiny-mamba2"):
        kw = dict(model_path=f"dummy:{preset}", page_size=4, chunked_prefill_size=16)
        _check(_run({**kw, "tp_size": 2}, 2, LP), _run(kw, 1, LP))


def test_gpt_oss_tp2():
    kw = dict(model_path="dummy:tiny-gpt-oss", page_size=4, chunked_prefill_size=16)
    _check(_run({**kw, "tp_size": 2}, 2, LP), _run(kw, 1, LP))


def test_rank_consensus_checker_passes_on_spmd_scheduler():
    """SGLB_RANK_CONSENSUS=crash: every batch decision of the two SPMD schedulers is hashed and min/max-reduced; a healthy run never diverges"""
    kw = dict(model_path="dummy:tiny-llama", page_size=4, chunked_prefill_size=16)
    _check(_run({**kw, "tp_size": 2}, 2, LP, {"SGLB_RANK_CONSENSUS": "crash"}), _run(kw, 1, LP))


def test_two_batch_overlap_with_dp_attention_a2a():
    """TBO under DP-attention + expert-parallel all-to-all: the two micro-batches of every group interleave their gathers / all-to-alls"""
    kw = dict(model_path="dummy:tiny-mixtral", page_size=4)
    ref = _run(kw, 1, LP)
    dpk = {**kw, "tp_size": 2, "dp_size": 2, "ep_size": 2, "enable_dp_attention": True, "enable_two_batch_overlap": True}
    _check(_run(dpk, 2, LP, {"SGLB_MOE_A2A": "1"}), ref)
    _check(_run({**kw, "tp_size": 2, "enable_two_batch_overlap": True}, 2, LP), ref)
    # uneven load: 3 requests over 2 groups -> one group prefills / decodes a single request (empty second micro-batch) while the other
    # splits two; idle forwards at the tail run two empty micro-batches
    odd = PROMPTS[:3]
    _check(_run(dpk, 2, LP, {"SGLB_MOE_A2A": "1"}, prompts=odd), _run(kw, 1, LP, prompts=odd))


def test_dwdp_prefetched_expert_weights_match_all_to_all_dispatch():
    """DWDP: prefills above the token threshold all-gather the peers' experts (layer-ahead) and run every expert locally; decode
    steps stay on the all-to-all path.  Both must agree with the single-process engine."""
    kw = dict(model_path="dummy:tiny-mixtral", page_size=4)
    ref = _run(kw, 1, LP)
    dpk = {**kw, "tp_size": 2, "dp_size": 2, "ep_size": 2, "enable_dp_attention": True, "enable_dwdp": True, "dwdp_min_tokens": 8}
    out = _run(dpk, 2, LP, {"SGLB_MOE_A2A": "1"})
    _check(out, ref)
    st = out["dwdp"]
    assert st["forwards"] > 0 and st["prefetches"] >= st["forwards"] // 2 and st["a2a_calls"] > 0      # prefill used DWDP, decode used a2a


def test_eplb_redundant_experts_and_elastic_ep_rank_loss():
    """EP2: (1) re-plan with one redundant slot per rank from a skewed load — weights move between ranks, hot experts get replicas, tokens
    are unchanged; (2) rank 1 'dies' (its expert memory is wiped): the survivor restores rank 1's experts from its host backup and serves
    every expert alone."""
    kw = dict(model_path="dummy:tiny-mixtral", page_size=4, disable_radix_cache=True)
    ref = _run(kw, 1)
    load = [[100.0, 1.0, 1.0, 50.0]] * 2
    out = _run({**kw, "tp_size": 2, "ep_size": 2}, 2, rpc=[{"method": "eplb_rebalance", "kwargs": {"redundant_slots_per_rank": 1, "load": load}}])
    assert out["ids"] == ref["ids"] and out["ids2"] == ref["ids"]
    r = out["rpc"][0]
    assert r["slots"] == 3 and max(r["imbalance_after"]) < max(r["imbalance_before"]) and r["moves"]["peer"] > 0
    out = _run({**kw, "tp_size": 2, "ep_size": 2}, 2, rpc=[{"method": "build_expert_backup"}, {"zero_rank": 1},
                                                           {"method": "elastic_ep_recover", "kwargs": {"dead_ranks": [1]}}])
    assert out["ids"] == ref["ids"] and out["ids2"] == ref["ids"]
    assert out["rpc"][0]["experts"] == 4 and out["rpc"][1]["live_ranks"] == [0] and out["rpc"][1]["moves"]["backup"] == 4    # 2 experts x 2 layers


def test_mamba1_jamba_tp2(tmp_path):
    """Mamba-1 channels sharded over TP (row-parallel x_proj with an all-reduce for dt / B / C), Jamba = Mamba-1 + attention + MoE"""
    import transformers as tf

    import torch

    torch.manual_seed(0)
    cfg = tf.JambaConfig(hidden_size=64, intermediate_size=128, num_hidden_layers=4, num_attention_heads=4, num_key_value_heads=2, vocab_size=512,
                         max_position_embeddings=512, attn_layer_period=2, attn_layer_offset=1, expert_layer_period=2, expert_layer_offset=1,
                         num_experts=4, num_experts_per_tok=2, mamba_d_state=8, mamba_d_conv=4, mamba_expand=2, mamba_dt_rank=4, pad_token_id=0,
                         use_mamba_kernels=False)
    hf = tf.JambaForCausalLM(cfg).float()
    with torch.no_grad():
        for p in hf.parameters():
            if p.dim() > 1:
                p.normal_(0, 0.08)
    d = str(tmp_path / "jamba")
    hf.save_pretrained(d, safe_serialization=True)
    kw = dict(model_path=d, page_size=4, chunked_prefill_size=16, skip_tokenizer_init=True)
    _check(_run({**kw, "tp_size": 2}, 2, LP), _run(kw, 1, LP), atol=5e-3)


def test_qwen3_next_tp2(tmp_path):
    """gated DeltaNet heads (k-head groups), conv channels and the gated softmax attention sharded over TP"""
    import torch
    import transformers as tf

    torch.manual_seed(0)
    cfg = tf.Qwen3NextConfig(hidden_size=64, intermediate_size=128, num_hidden_layers=4, num_attention_heads=4, num_key_value_heads=2, vocab_size=512,
                             max_position_embeddings=512, head_dim=16, partial_rotary_factor=0.25, linear_num_key_heads=2, linear_num_value_heads=4,
                             linear_key_head_dim=8, linear_value_head_dim=8, linear_conv_kernel_dim=4, num_experts=4, num_experts_per_tok=2,
                             moe_intermediate_size=32, shared_expert_intermediate_size=48, decoder_sparse_step=1, mlp_only_layers=[],
                             layer_types=["linear_attention", "linear_attention", "linear_attention", "full_attention"], tie_word_embeddings=False)
    hf = tf.Qwen3NextForCausalLM(cfg).float()
    with torch.no_grad():
        for p in hf.parameters():
            if p.dim() > 1:
                p.normal_(0, 0.08)
    d = str(tmp_path / "q3n")
    hf.save_pretrained(d, safe_serialization=True)
    kw = dict(model_path=d, page_size=4, chunked_prefill_size=16, skip_tokenizer_init=True)
    _check(_run({**kw, "tp_size": 2}, 2, LP), _run(kw, 1, LP), atol=5e-3)


def test_multi_node_launch_two_nodes_tp4():
    """``--nnodes 2``: two launch_server invocations (one per "node", both on this host) each spawn their local ranks and rendezvous at
    --dist-init-addr; node 0 serves HTTP for the whole TP=4 job"""
    import signal
    import time
    import urllib.request

    http, rdv = _free_port(), _free_port()
    procs = []
    try:
        for node in (0, 1):
            env = dict(os.environ, PYTHONPATH=ROOT, OMP_NUM_THREADS="2")
            for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT"):
                env.pop(k, None)
            procs.append(subprocess.Popen([sys.executable, "-m", "sglang_b200.launch_server", "--model-path", "dummy:tiny-llama", "--device", "cpu",
                                           "--page-size", "4", "--tp", "4", "--nnodes", "2", "--node-rank", str(node), "--dist-init-addr",
                                           f"127.0.0.1:{rdv}", "--port", str(http), "--log-level", "error", "--skip-tokenizer-init"],
                                          env=env, cwd=ROOT, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True, start_new_session=True))
        body = json.dumps({"input_ids": PROMPTS[0], "sampling_params": SAMPLING}).encode()
        out, t0 = None, time.time()
        while time.time() - t0 < 180:
            if any(p.poll() is not None for p in procs):
                raise AssertionError("a node exited: " + "".join((p.stderr.read() or "")[-1500:] for p in procs if p.poll() is not None))
            try:
                req = urllib.request.Request(f"http://127.0.0.1:{http}/generate", data=body, headers={"Content-Type": "application/json"})
                out = json.loads(urllib.request.urlopen(req, timeout=60).read())
                break
            except OSError:
                time.sleep(1.0)
        assert out is not None, "server did not come up"
        ref = _run(dict(model_path="dummy:tiny-llama", page_size=4), 1)
        assert out["output_ids"] == ref["ids"][0]
    finally:
        for p in procs:
            try:
                os.killpg(p.pid, signal.SIGKILL)       # the node's own process group: launcher + the ranks it spawned
            except ProcessLookupError:
                pass


def test_sequence_parallel_prefill():
    """--enable-sp: residual stream scattered over the token rows between the TP blocks (reduce-scatter after o_proj / down_proj, all-gather
    before qkv / gate_up); token counts not divisible by tp exercise the padding; decode keeps the plain path"""
    env = {"SGLB_SP_MIN_TOKENS": "1", "SGLB_TEST_EXPECT_SP": "1"}
    for preset in ("tiny-llama", "tiny-phi3"):
        kw = dict(model_path=f"dummy:{preset}", page_size=4, chunked_prefill_size=32)
        ref = _run(kw, 1, LP)
        _check(_run({**kw, "tp_size": 2, "enable_sp": True}, 2, LP, env), ref)
    kw = dict(model_path="dummy:tiny-llama", page_size=4, chunked_prefill_size=32)
    _check(_run({**kw, "tp_size": 4, "enable_sp": True}, 4, LP, env), _run(kw, 1, LP))


def test_lora_tp2_matches_single():
    """adapters sharded like their layers (B rows of column-parallel, A columns of row-parallel projections); mixed batch of two adapters + base"""
    kw = dict(model_path="dummy:tiny-llama", page_size=4, enable_lora=True, lora_paths=["a=dummy:rank=4", "b=dummy:rank=8"])
    gen = {**LP, "lora_path": ["a", None, "b", "a"]}
    ref = _run(kw, 1, gen)
    _check(_run({**kw, "tp_size": 2}, 2, gen), ref)
    assert ref["ids"] != _run(dict(model_path="dummy:tiny-llama", page_size=4), 1, LP)["ids"]


def test_vocab_parallel_greedy_matches_gathered_argmax():
    """TP decode: per-shard argmax + 8-byte exchange + merge == argmax over the all-gathered logits (layers/linear.py vocab_parallel_greedy)"""
    _run({"model_path": "dummy:tiny-llama", "tp_size": 2, "page_size": 4}, 2, extra_env={"SGLB_TEST_GREEDY_SHARD": "1"})


def test_minimax_m2_full_width_qk_norm_tp(tmp_path):
    """q/k RMSNorm over all heads of the layer: the mean square crosses TP ranks (one [T, 1] all-reduce); tp=4 > 2 KV heads exercises the
    replicated-KV discount"""
    import torch
    import transformers as tf

    torch.manual_seed(0)
    cfg = tf.MiniMaxM2Config(hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2, vocab_size=512,
                             max_position_embeddings=512, head_dim=16, num_local_experts=4, num_experts_per_tok=2)
    hf = tf.MiniMaxM2ForCausalLM(cfg).float()
    with torch.no_grad():
        for n, p in hf.named_parameters():
            if p.dim() > 1:
                p.normal_(0, 0.08)
            elif "norm" in n:
                p.add_(torch.randn_like(p) * 0.1)
    d = str(tmp_path / "mm2")
    hf.save_pretrained(d, safe_serialization=True)
    kw = dict(model_path=d, page_size=4, chunked_prefill_size=16, skip_tokenizer_init=True)
    ref = _run(kw, 1, LP)
    _check(_run({**kw, "tp_size": 2}, 2, LP), ref, atol=5e-3)
    _check(_run({**kw, "tp_size": 4}, 4, LP), ref, atol=5e-3)


def test_qwen3_5_tp2(tmp_path):
    """Qwen3.5 checkpoints (flat in_proj_qkv / z / b / a) regrouped per key head at load, then sharded like Qwen3-Next"""
    import torch
    import transformers as tf

    torch.manual_seed(0)
    cfg = tf.Qwen3_5TextConfig(hidden_size=64, intermediate_size=128, num_hidden_layers=4, num_attention_heads=4, num_key_value_heads=2, vocab_size=512,
                               max_position_embeddings=512, head_dim=16, linear_num_key_heads=2, linear_num_value_heads=4, linear_key_head_dim=8,
                               linear_value_head_dim=8, linear_conv_kernel_dim=4,
                               layer_types=["linear_attention", "linear_attention", "linear_attention", "full_attention"], tie_word_embeddings=False)
    hf = tf.Qwen3_5ForCausalLM(cfg).float()
    with torch.no_grad():
        for p in hf.parameters():
            if p.dim() > 1:
                p.normal_(0, 0.08)
    d = str(tmp_path / "q35")
    hf.save_pretrained(d, safe_serialization=True)
    kw = dict(model_path=d, page_size=4, chunked_prefill_size=16, skip_tokenizer_init=True)
    _check(_run({**kw, "tp_size": 2}, 2, LP), _run(kw, 1, LP), atol=5e-3)
