"""profiles/attn_decode_r2.md from the A/B JSONs (tools/attn_decode_bench.py) and the two ncu reports of run E."""
import csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
def cases(tag):
    try: return json.load(open(os.path.join(G, f"attn_decode_bench_{tag}.json")))["cases"]
    except Exception: return {}
tabs = {"v2 query-major (r1)": cases("v2"), "v3 final": cases("v3_final"), "v3 two softmax groups": cases("v3_2groups"),
        "v3 K/V producer warps + 5-D V box": cases("v3_pf0"), "v3 + L2 prefetcher (222 items)": cases("v3_pf222"),
        "FlashInfer 0.6.11 trtllm-gen": cases("fi")}
names = list(tabs["v3 final"].keys())
out = ["# Decode attention A/B on identical shapes (round 2; 1 x B200; tools/attn_decode_bench.py)", "",
       "CUDA events, 256 MB L2 flush between iterations, median of 7.  Llama-3-8B heads (G = 4, D = 128), bf16 KV, page 32.  `frac` = page-granular KV",
       "bytes / time / MEASURED_PEAKS.json hbm_gbs (6569.6 GB/s) — \"of measured\".  v2 = `attention_tc.cu` persistent KT=64 (round 1); v3 =",
       "`attention_decode.cu` (keys-as-M; \"final\" = what ships, the other v3 columns are experiments that were measured and removed);", "FlashInfer = `trtllm_batch_decode_with_kv_cache` (the reference's default backend on sm_100).", "",
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
        "| metric | v3 (run E) | FlashInfer trtllm-gen (run E) |", "|---|---|---|"]
for w in want:
    out.append(f"| {w} | {a.get(w, '-')} | {b.get(w, '-')} |")
out += ["", "## Per-role cycle accounting (instrumented instantiation of the same template, `tools/attn_decode_bench.py v3 <case> prof`, run I)", "",
        "Cycles per 128-key tile (64 KB of KV), summed waits of each role's elected thread, 148 CTAs; a tile takes ~4 300 cycles end to end.", "",
        "| case | producer waits (k_empty / v_empty) | MMA waits (k_full / p_full / v_full) | softmax: wait s_full | softmax: work (ld + max + rescale + P + arrive) | epilogue per tile |",
        "|---|---|---|---|---|---|",
        "| bs256_ctx512 | 126 / 188 | 1065 / 119 / 124 | 1480 | 36 + 233 + 250 + 224 + 37 = 780 | 815 |",
        "| tp8shard_bs1024_ragged | 121 / 159 | 989 / 117 / 118 | 1734 | 786 | 522 |",
        "| bs64_ctx2048 | 104 / 393 | 1102 / 112 / 114 | 2029 | 783 | 208 |",
        "", "## Reading",
        "* Same bytes as the library (1.944 GB on the trace-like case, page-granular tail fetch), 60 % of DRAM peak in ncu vs 83 % for trtllm-gen on that one",
        "  shape; on uniform shapes and on the `Hkv = 1` probes (contiguous 8 KB pages) the library sits at 0.61–0.74 as well.  Pool layout does not matter",
        "  (`probe_hkv1_*` = same bytes, contiguous pages: same fraction).",
        "* The softmax warps do ~780 cycles of work per tile and WAIT 1 500–2 000 cycles for `s_full`; the MMA thread waits ~1 050 cycles per tile for",
        "  `k_full`; the producers wait on full rings.  Nobody is busy: the kernel sits on the loaded-latency curve of the paged gather —",
        "  `tools/tma_gather_bench.cu` (profiles/tma_gather_r2.md): 96 KB in flight per SM -> 4.5 TB/s, 192 KB -> 5.7 TB/s (5 us loaded latency), and a",
        "  consumer that holds stages for 2 x 1 500 cycles drops 5.7 to 5.1.  The rings already take 192 of the 227 KB of shared memory.",
        "* Experiments that followed from wrong or partial readings of the above, measured and removed: two softmax groups (no change: the consumer was",
        "  not the limiter, and with two tiles in the consumer fewer stages are left in flight); separate K / V producer warps + one 5-D TMA box per V page",
        "  (the micro-benchmark's issue-rate limit is real at 5.7+ TB/s but is not what binds at 4.5; slower on ragged batches); an L2 prefetcher warp",
        "  walking the work queue ahead of its head (+30 % time at every look-ahead: the prefetches are extra TMA / L2 traffic, not extra bandwidth).",
        "* What is left on the table: more bytes in flight per byte of shared memory — fp8 KV (half the bytes per key), 2-CTA clusters sharing one",
        "  K/V ring across two query groups — and the epilogue (800 cycles per tile on 4-tile items: the softmax warps also do the O^T read-out)."]
open(os.path.join(ROOT, "profiles", "attn_decode_r2.md"), "w").write("\n".join(out) + "\n")
print("\n".join(out[:30]))
