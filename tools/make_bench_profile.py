"""profiles/bench_r2.md from the bench JSON lines of the round-2 GPU runs (gpurun_out/bench_r2*.json, vllm_probe.json)."""
import glob, json, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
def line(name):
    p = os.path.join(G, name)
    if not os.path.exists(p): return None
    for l in open(p):
        l = l.strip()
        if l.startswith("{") and ('"metric"' in l or '"unavailable"' in l):
            try: return json.loads(l)
            except Exception: pass
    return None
rows = [("reference (unmodified, `baseline/_ref`, mode in config)", line("bench_r2o_ref.json") or line("bench_r2f_ref.json")),
        ("**ours**, TP = 1", line("bench_r2o_ours.json") or line("bench_r2f_ours.json")),
        ("vLLM 0.22 (`--impl vllm`, run C, steps 1 / warmup 1)", line("vllm_probe.json")),
        ("ours, TP = 4 (run H, 4 x B200)", line("bench_r2h_tp4.json")), ("ours, TP = 8 (run N, 8 x B200)", line("bench_r2n_tp8.json"))]
out = ["# Headline benchmark, round 2 (same box per section, same synthetic ShareGPT-shaped trace, Llama-3-8B bf16, random-init weights)", "",
       "`python bench.py [--impl reference|vllm] --gpus N --steps K --warmup W`; 1 024 requests per trace (mean 271 in / 208 out tokens), request rate",
       "infinite, `ignore_eos`, greedy.  `value` = output tokens / s over the K timed traces (ours: CUDA events on the engine stream, max over ranks;",
       "reference / vLLM: wall clock around their own `generate`, their GPU work lives in child processes).", "",
       "| arm | output tok/s | e2e tok/s | ms / trace | steps / warmup | SM clock (median / max), throttle reasons | notes |", "|---|---|---|---|---|---|---|"]
for name, d in rows:
    if d is None:
        out.append(f"| {name} | not run | | | | | |"); continue
    if "unavailable" in d:
        out.append(f"| {name} | unavailable | | | | | {d['unavailable'][:200]} |"); continue
    c = d.get("clocks") or {}
    cfg = d.get("config", {})
    note = []
    if cfg.get("reference_mode"): note.append(f"mode `{cfg['reference_mode']}`: " + ", ".join(f"{k}={v}" for k, v in (cfg.get('reference_flags') or {}).items()))
    if d.get("step_split"): note.append(f"host {d['step_split']['host_ms_per_iteration']:.2f} ms / iteration, device {d['step_split']['device_ms_per_forward']:.2f} ms / forward, {d.get('gpu_launches')} kernel launches")
    if cfg.get("modes_that_failed"): note.append("modes that failed first: " + "; ".join(m[:90] for m in cfg["modes_that_failed"]))
    e2e = (d.get("e2e") or {}).get("value")
    out.append(f"| {name} | {d['value']:.0f} | {e2e:.0f} | {d['ms_per_step']:.0f} | {d['steps']} / {d['warmup']} | {c.get('sm_mhz')} / {c.get('sm_max_mhz')}, {c.get('reasons')} | {' ; '.join(note)} |")
ours, ref = rows[1][1], rows[0][1]
if ours and ref and "value" in ours and "value" in ref:
    out += ["", f"Ratio ours / reference on the same box, back to back: **{ours['value'] / ref['value']:.3f}** on the device-timed number, "
                f"**{ours['e2e']['value'] / ref['value']:.3f}** end to end (text in -> text out vs the reference's ids in -> ids out)."]
out += ["", "History of the own arm this round (all TP = 1, steps 3 / warmup 3): 16 257 (round 1) -> 17 147 (keys-as-M decode kernel, run D) -> 17 084 (run F) ->",
        "16 349 with the L2 prefetcher warp (run L, reverted).  TP scaling and its limits: profiles/mgpu_tp4_r2.md, profiles/mgpu_tp8_r2.md.",
        "", "Online serving (run K, `tools/online_bench.sh --enable-multi-process-frontend --prefill-cuda-graph-max-tokens 2048`, random 512–1024-token prompts,",
        "64–128 output tokens): 1 req/s: TTFT mean 18.0 / median 17.7 / p99 26.9 ms, ITL mean 3.95 ms; 8 req/s: TTFT mean 20.3 / p99 42.3 ms, ITL mean 4.58 ms."]
open(os.path.join(ROOT, "profiles", "bench_r2.md"), "w").write("\n".join(out) + "\n")
print("\n".join(out))
