"""Profile one bench trace: torch.profiler kernel table (top kernels by device time, split prefill/decode) + host-side
scheduler phase timing.  Output -> gpurun_out/profile_<tag>.txt"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from torch.profiler import ProfilerActivity, profile

import bench
import sglang_b200 as sgl

ap = argparse.ArgumentParser()
ap.add_argument("--num-prompts", type=int, default=64)
ap.add_argument("--attn", default="tc")
ap.add_argument("--gemm", default="lib")
ap.add_argument("--tag", default="p")
ap.add_argument("--max-out", type=int, default=64)
ap.add_argument("--no-graph", action="store_true")
a = ap.parse_args()

eng = sgl.Engine(model_path="dummy:llama-3-8b", page_size=32, attention_backend=a.attn, gemm_backend=a.gemm, max_running_requests=256,
                 cuda_graph_max_bs=256, log_level="warning", disable_cuda_graph=a.no_graph)
pin, pout = bench.sharegpt_shaped_trace(a.num_prompts, 0)
pout = [min(p, a.max_out) for p in pout]
rng = np.random.default_rng(1)
ids = lambda: [rng.integers(300, 100000, size=n).tolist() for n in pin]
sps = [{"max_new_tokens": m, "temperature": 0, "ignore_eos": True} for m in pout]
eng.generate(input_ids=ids(), sampling_params=sps)
torch.cuda.synchronize()

# host-side phase timing: wrap scheduler methods
sch = eng.scheduler
phases = {}
def wrap(obj, name):
    fn = getattr(obj, name)
    def w(*x, **k):
        t = time.perf_counter(); r = fn(*x, **k); phases[name] = phases.get(name, 0) + time.perf_counter() - t
        phases[name + "#"] = phases.get(name + "#", 0) + 1
        return r
    setattr(obj, name, w)
for n in ("get_next_batch_to_run", "run_batch", "process_batch_result", "_prepare_for_extend", "_update_running_batch"):
    wrap(sch, n)
t0 = time.perf_counter()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    eng.generate(input_ids=ids(), sampling_params=sps)
    torch.cuda.synchronize()
wall = time.perf_counter() - t0
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", f"profile_{a.tag}.txt"), "w") as f:
    f.write(f"wall {wall:.3f}s  in_tokens {sum(pin)} out_tokens {sum(pout)} forward_ct {eng.runner.forward_ct}\n")
    f.write(json.dumps({k: round(v, 4) for k, v in phases.items()}) + "\n")
    f.write(prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=70))
print(open(os.path.join(ROOT, "gpurun_out", f"profile_{a.tag}.txt")).read()[:6000])
