"""Decode attention A/B on identical shapes (Llama-3-8B heads, page 32): our query-major kernel (v2, attention_tc.cu), our keys-as-M kernel
(v3, attention_decode.cu) and FlashInfer's trtllm-gen decode (the reference's default on sm_100: trtllm_mha_backend.py:1470) — CUDA events,
L2 flushed between iterations, median of 7; roofline = KV bytes / measured HBM copy bandwidth (MEASURED_PEAKS.json).
Runs in TWO processes for v2 / v3 (the choice is read once at import): python tools/attn_decode_bench.py [v2|v3|fi] -> gpurun_out/attn_decode_bench_<tag>.json"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
tag = sys.argv[1] if len(sys.argv) > 1 else "v3"
if tag in ("v2", "v3"):
    os.environ["SGLB_ATTN_DECODE"] = tag
import numpy as np
import torch
from sglang_b200 import ops
from gpu_cases import _attn_setup
ops.load_native()
dev = "cuda"
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650}
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

def timeit(fn, iters=7, warm=3):
    for _ in range(warm): fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort(); return ts[len(ts) // 2]

rng = np.random.default_rng(0)
CASES = {
    "bs256_ctx512": (8, [512] * 256, 1), "bs256_ragged32-1500": (8, rng.integers(32, 1500, 256).tolist(), 1),
    "bs1024_ragged(trace-like)": (8, np.clip(rng.lognormal(5.9, 0.7, 1024), 8, 3000).astype(int).tolist(), 1),
    "bs512_ctx400": (8, [400] * 512, 1), "bs64_ctx2048": (8, [2048] * 64, 1), "bs16_ctx1024_split": (8, [1024] * 16, 4),
    "bs8_ctx4096_split8": (8, [4096] * 8, 8), "tp4shard_bs512_ragged": (2, rng.integers(32, 1500, 512).tolist(), 1),
    "tp8shard_bs1024_ragged": (1, rng.integers(32, 1500, 1024).tolist(), 1),
    # layout probes: Hkv = 1 makes every (page, head) one contiguous 8 KB run; same bytes as the Hkv = 8 cases above
    "probe_hkv1_bs2048_ctx512": (1, [512] * 2048, 1),
    "probe_hkv1_bs8192_ragged": (1, np.clip(rng.lognormal(5.9, 0.7, 8192), 8, 3000).astype(int).tolist(), 1),
}
res = {}
only = sys.argv[2] if len(sys.argv) > 2 else None          # optional: one case only (ncu captures)
for name, (Hkv, lens, splits) in CASES.items():
    if only and only not in name:
        continue
    B, Hq, ps = len(lens), Hkv * 4, 32
    s = _attn_setup(B, [1] * B, lens, Hq, Hkv, 128, ps)
    out = torch.empty_like(s["q"])
    kv_bytes = sum(((l + ps - 1) // ps) * ps for l in lens) * Hkv * 128 * 2 * 2        # page-granular (what any paged kernel must move)
    exact = sum(lens) * Hkv * 128 * 2 * 2
    if tag == "fi":
        import flashinfer
        # trtllm-gen wants HND pages [pages, 2, Hkv, ps, D]; build it from the same cache contents
        n_pages = s["kc"].shape[0] // ps
        kvc = torch.stack([s["kc"].view(n_pages, ps, Hkv, 128).transpose(1, 2), s["vc"].view(n_pages, ps, Hkv, 128).transpose(1, 2)], 1).contiguous()
        bt = s["pt"][s["slots"].long()].contiguous()
        wsb = torch.zeros(256 << 20, dtype=torch.uint8, device=dev)
        seq = s["seq"].clone()
        f = lambda: flashinfer.decode.trtllm_batch_decode_with_kv_cache(s["q"], kvc, wsb, bt, seq, max(lens), 128 ** -0.5, 1.0, out=out)
    else:
        ws = None
        if splits > 1:
            ws = (torch.empty(splits, B, Hq, 128, dtype=torch.float32, device=dev), torch.empty(splits, B, Hq, dtype=torch.float32, device=dev))
        f = lambda: ops.paged_attention(s["q"], s["kc"], s["vc"], s["pt"], s["slots"], s["seq"], None, None, 1, ps, 128 ** -0.5,
                                        num_splits=splits, workspace=ws, impl=2, out=out)
    try:
        t = timeit(f)
        res[name] = {"us": round(t * 1e3, 1), "kv_gbs_page_granular": round(kv_bytes / t / 1e6, 1), "frac_of_measured_hbm": round(kv_bytes / t / 1e6 / peaks["hbm_gbs"], 3),
                     "frac_exact_bytes": round(exact / t / 1e6 / peaks["hbm_gbs"], 3)}
    except Exception as e:
        res[name] = {"error": f"{type(e).__name__}: {str(e)[:200]}"}
    print(tag, name, res[name], flush=True)
    if len(sys.argv) > 3 and sys.argv[3] == "prof" and tag == "v3" and Hkv * 4 == Hq:
        # per-role cycle accounting (instrumented build of the same kernel): where does a tile's time go?
        buf = torch.zeros(148 * 40, dtype=torch.int64, device=dev)
        torch.ops.sglb.attn_decode_prof(buf)
        flush.zero_(); f(); torch.cuda.synchronize()
        torch.ops.sglb.attn_decode_prof(torch.empty(0, dtype=torch.int64, device=dev))
        a = buf.view(148, 4, 10).double().cpu().numpy()
        tiles = a[:, 2, 8].sum()
        names = [["item_empty", "q_empty", "k_empty", "v_empty", "(tiles)"],
                 ["item_full", "k_full", "p_full", "v_full", "o_free", "q_full"],
                 ["item_full", "s_full", "tmem_ld", "max_exchange", "rescale", "p_store+fence", "arrive", "epilogue", "(tiles)"],
                 ["item_full", "-", "-", "v_empty", "(tiles)"]]
        print(f"  prof: {int(tiles)} tiles over 148 CTAs; cycles per tile, summed waits of each role's elected thread")
        for r, role in enumerate(["producer waits", "mma waits", "softmax"]):
            print("   ", role + ":", ", ".join(f"{n} {a[:, r, i].sum() / tiles:.0f}" for i, n in enumerate(names[r]) if not n.startswith("(") and n != "-"))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
if not only:
    json.dump({"impl": tag, "hbm_gbs_measured": peaks["hbm_gbs"], "cases": res}, open(os.path.join(ROOT, "gpurun_out", f"attn_decode_bench_{tag}.json"), "w"), indent=1)
