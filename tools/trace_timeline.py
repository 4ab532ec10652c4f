"""Where does a bench trace spend its device time?  Wraps ModelRunner.forward_batch_generation with CUDA events (on the runner's
stream, resolved after the trace) and buckets the steps: prefill/extend vs decode by batch size.  -> gpurun_out/timeline_<tag>.json/.txt"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench
import sglang_b200 as sgl

ap = argparse.ArgumentParser()
ap.add_argument("--num-prompts", type=int, default=1024); ap.add_argument("--max-running", type=int, default=1024)
ap.add_argument("--tag", default="t"); ap.add_argument("--chunk", type=int, default=None)
a = ap.parse_args()
kw = dict(chunked_prefill_size=a.chunk) if a.chunk else {}
eng = sgl.Engine(model_path="dummy:llama-3-8b", page_size=32, max_running_requests=a.max_running, cuda_graph_max_bs=a.max_running,
                 log_level="warning", **kw)
pin, pout = bench.sharegpt_shaped_trace(a.num_prompts, 0)
V = eng.model_config.vocab_size
def ids(seed):
    r = np.random.default_rng(seed); return [r.integers(300, V - 1000, size=n).tolist() for n in pin]
sps = [{"max_new_tokens": m, "temperature": 0, "ignore_eos": True} for m in pout]
for w in range(2):
    eng.generate(input_ids=ids(w), sampling_params=sps)
runner = eng.runner
orig = runner.forward_batch_generation
recs = []
def wrapped(sb):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(runner.stream)
    r = orig(sb)
    e.record(runner.stream)
    dec = sb.forward_mode.is_decode()
    recs.append((s, e, dec, len(sb.reqs), int(sb.q_lens.sum()) if not dec else len(sb.reqs)))
    return r
runner.forward_batch_generation = wrapped
torch.cuda.synchronize()
t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
t0.record(runner.stream)
eng.generate(input_ids=ids(7), sampling_params=sps)
t1.record(runner.stream); torch.cuda.synchronize()
total = t0.elapsed_time(t1)
pre_ms = pre_tok = 0.0; n_pre = 0
buckets = {}
for s, e, dec, bs, toks in recs:
    ms = s.elapsed_time(e)
    if dec:
        b = 1 << max(0, (bs - 1).bit_length())
        d = buckets.setdefault(b, [0, 0.0, 0]); d[0] += 1; d[1] += ms; d[2] += bs
    else:
        pre_ms += ms; pre_tok += toks; n_pre += 1
busy = pre_ms + sum(v[1] for v in buckets.values())
out = {"trace_ms": total, "device_busy_ms": busy, "idle_ms": total - busy, "prefill": {"steps": n_pre, "ms": pre_ms, "tokens": pre_tok,
       "tok_per_s": pre_tok / pre_ms * 1e3 if pre_ms else 0}, "decode": {str(k): {"steps": v[0], "ms": v[1], "tokens": v[2], "ms_per_step": v[1] / v[0]} for k, v in sorted(buckets.items())},
       "output_tok_per_s": sum(pout) / total * 1e3}
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"timeline_{a.tag}.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
eng.shutdown(); sys.stdout.flush(); os._exit(0)
