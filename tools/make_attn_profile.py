"""profiles/attn_decode_r2.md from the A/B JSONs (tools/attn_decode_bench.py) and the two ncu reports of run E."""
import csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
def cases(tag):
    try: return json.load(open(os.path.join(G, f"attn_decode_bench_{tag}.json")))["cases"]
    except Exception: return {}
tabs = {"v2 (query-major, r1)": cases("v2"), "v3 one softmax group": cases("v3_1group"), "v3 two softmax groups": cases("v3"),
        "FlashInfer 0.6.11 trtllm-gen": cases("fi")}
names = list(tabs["v2 (query-major, r1)"].keys())
out = ["# Decode attention A/B on identical shapes (round 2; 1 x B200; tools/attn_decode_bench.py)", "",
       "CUDA events, 256 MB L2 flush between iterations, median of 7.  Llama-3-8B heads (G = 4, D = 128), bf16 KV, page 32.  `frac` = page-granular KV",
       "bytes / time / MEASURED_PEAKS.json hbm_gbs (6569.6 GB/s) — \"of measured\".  v2 = `attention_tc.cu` persistent KT=64 (round 1); v3 =",
       "`attention_decode.cu` (keys-as-M); FlashInfer = `trtllm_batch_decode_with_kv_cache` (the reference's default backend on sm_100).", "",
       "| case | " + " | ".join(f"{k} us / frac" for k in tabs) + " |", "|---|" + "---|" * len(tabs)]
for n in names:
    row = [n]
    for k, t in tabs.items():
        c = t.get(n, {})
        row.append(f"{c.get('us', '-')} / {c.get('frac_of_measured_hbm', '-')}" if "us" in c else c.get("error", "-")[:40])
    out.append("| " + " | ".join(row) + " |")
def ncu(rep):
    p = os.path.join(G, rep)
    if not os.path.exists(p): return {}
    txt = subprocess.run(["ncu", "-i", p, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    if len(rows) < 3: return {}
    return dict(zip(rows[0], rows[2]))
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]
a, b = ncu("ncu_decode_v3.ncu-rep"), ncu("ncu_decode_fi.ncu-rep")
c = ncu("ncu_decode_v3_2g.ncu-rep")
out += ["", "## ncu --set full --clock-control none, case bs1024_ragged(trace-like) (8192 work items, 1.94 GB of KV)", "",
        "| metric | v3 one group (run E) | v3 two groups | FlashInfer trtllm-gen (run E) |", "|---|---|---|---|"]
for w in want:
    out.append(f"| {w} | {a.get(w, '-')} | {c.get(w, '-')} | {b.get(w, '-')} |")
out += ["", "## Reading",
        "* Both kernels move the same 1.944 GB (no over-fetch on our side: page-granular tail fetch).  The one-group kernel reached 60 % of DRAM peak, the",
        "  library 83 %: tensor pipe 4 %, 6 warps/SM, top stall = long scoreboard — the serial chain `s_full -> tcgen05.ld -> row max (shuffles + named",
        "  barrier) -> exp2 -> P to smem -> fence.proxy.async -> p_full` costs ~3.0 k cycles per 128-key tile and ONE group of softmax warps runs it for",
        "  every tile, while HBM needs a 64 KB tile every 2.8 k cycles per SM.",
        "* Fix (this round): two softmax groups, tile j -> group j & 1, private S^T / P / O^T, partial softmax states merged in the CTA at the end of",
        "  the item (`attention_decode.cu`).  Double-buffered Q and a prefetched queue atomic alone changed nothing (run D), which is what pointed at the",
        "  softmax chain rather than per-item latency."]
open(os.path.join(ROOT, "profiles", "attn_decode_r2.md"), "w").write("\n".join(out) + "\n")
print("\n".join(out[:30]))
