"""Kernel-level breakdown of steady-state decode: prompts are prefilled once (filling the radix cache), then the same
prompts are decoded for N tokens under torch.profiler — the second pass re-uses the cached prefix, so nearly all device time
is decode steps (CUDA-graph replays).   python tools/decode_breakdown.py --bs 256 --ctx 512 --out 48 [--tp N via torchrun]"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np
import torch
from torch.profiler import ProfilerActivity, profile
import sglang_b200 as sgl

ap = argparse.ArgumentParser()
ap.add_argument("--bs", type=int, default=256); ap.add_argument("--ctx", type=int, default=512); ap.add_argument("--out", type=int, default=48)
ap.add_argument("--attn", default="tc"); ap.add_argument("--gemm", default="lib"); ap.add_argument("--tag", default="decode")
ap.add_argument("--model", default="llama-3-8b")
a = ap.parse_args()
world = int(os.environ.get("WORLD_SIZE", 1)); rank = int(os.environ.get("RANK", 0))
eng = sgl.Engine(model_path=f"dummy:{a.model}", tp_size=world, page_size=32, attention_backend=a.attn, gemm_backend=a.gemm,
                 max_running_requests=a.bs, cuda_graph_max_bs=a.bs, log_level="warning")
rng = np.random.default_rng(0)
prompts = [rng.integers(300, 100000, size=a.ctx).tolist() for _ in range(a.bs)]
sp1 = {"max_new_tokens": 1, "temperature": 0, "ignore_eos": True}
spn = {"max_new_tokens": a.out, "temperature": 0, "ignore_eos": True}
eng.generate(input_ids=prompts, sampling_params=sp1)
eng.generate(input_ids=prompts, sampling_params=spn)      # warm graphs for this bs
torch.cuda.synchronize()
f0 = eng.runner.forward_ct
t0 = time.perf_counter()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    eng.generate(input_ids=prompts, sampling_params=spn)
    torch.cuda.synchronize()
wall = time.perf_counter() - t0
steps = eng.runner.forward_ct - f0
if rank == 0:
    ka = prof.key_averages()
    tot = sum(k.device_time_total for k in ka) / 1000.0
    rows = sorted(((k.device_time_total / 1000.0, k.count, k.key) for k in ka if k.device_time_total > 0), reverse=True)
    path = os.path.join(ROOT, "gpurun_out", f"breakdown_{a.tag}.txt")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "w") as f:
        f.write(f"bs={a.bs} ctx={a.ctx} out={a.out} tp={world} attn={a.attn} gemm={a.gemm}: wall {wall*1000:.1f} ms, forward steps {steps}, "
                f"device busy {tot:.1f} ms, wall/step {wall*1000/steps:.3f} ms, busy/step {tot/steps:.3f} ms, tok/s {a.bs*a.out/wall:.0f}\n")
        for ms, cnt, key in rows[:30]:
            f.write(f"{ms:9.2f} ms {100*ms/tot:5.1f}%  n={cnt:6d}  avg {1000*ms/cnt:8.1f} us  {key[:110]}\n")
    print(open(path).read())
eng.shutdown()
sys.stdout.flush()
os._exit(0)
