#!/usr/bin/env python
"""Headline benchmark (BASELINE.json): output tokens/s (whole box, device-timed, max over ranks) + p50 TTFT for
Llama-3-8B bf16 at TP = --gpus, synthetic ShareGPT-shaped request trace, random-init weights.

A *step* = serving one complete trace (``--num-prompts`` requests, request-rate = inf, continuous batching with
chunked prefill + radix cache) through the engine.  ``value``   : trace replayed from pre-tokenised ids, timed with
CUDA events on the engine stream (max over ranks).  ``e2e.value``: the same trace through the public API
``Engine.generate(text prompts) -> text``: tokenisation, per-step pinned-host->device metadata copies, device->host
token reads and detokenisation all inside a barrier+synchronize-bracketed wall-clock region.

    python bench.py --gpus 1 --steps 3 --warmup 3
    python -m torch.distributed.run --nproc-per-node 8 ... bench.py --gpus 8
    python bench.py --impl reference ...     # the unmodified reference from baseline/_ref (if it can run)
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import faulthandler
if os.environ.get("SGLB_HANG_DUMP_S"):
    faulthandler.dump_traceback_later(int(os.environ["SGLB_HANG_DUMP_S"]), repeat=False, exit=True)

import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "vllm"])
    ap.add_argument("--model", default="llama-3-8b")
    ap.add_argument("--num-prompts", type=int, default=1024)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--attention-backend", default=os.environ.get("SGLB_BENCH_ATTN", "auto"))
    ap.add_argument("--gemm-backend", default=os.environ.get("SGLB_BENCH_GEMM", "auto"))
    ap.add_argument("--page-size", type=int, default=32)
    ap.add_argument("--max-running", type=int, default=1024)
    ap.add_argument("--disable-cuda-graph", action="store_true")
    ap.add_argument("--disable-overlap", action="store_true")
    ap.add_argument("--comm-backend", default="auto")
    ap.add_argument("--quantization", default=None)
    return ap.parse_args()


def sharegpt_shaped_trace(n: int, seed: int, max_in: int = 2048, max_out: int = 1024):
    """(prompt_len, output_len) pairs with ShareGPT-like heavy-tailed (log-normal) lengths; the dataset itself is not
    available offline (reference sampler: python/sglang/benchmark/datasets/sharegpt.py:82-129)."""
    import numpy as np

    rng = np.random.default_rng(seed)
    pin = np.clip(rng.lognormal(mean=5.2, sigma=1.0, size=n), 4, max_in).astype(int)     # median ~180, mean ~300
    pout = np.clip(rng.lognormal(mean=4.95, sigma=0.9, size=n), 4, max_out).astype(int)  # median ~140, mean ~210
    return pin.tolist(), pout.tolist()


class ClockSampler(threading.Thread):
    """samples SM clock / throttle reasons through NVML while the timed region runs"""

    def __init__(self, index: int = 0, period: float = 0.2):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.stop_evt = threading.Event()
        self.sm, self.reasons, self.max_mhz, self.power = [], set(), None, []

    def run(self):
        try:
            import pynvml as nv

            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap", nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                     nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                     nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                     nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake"}
            while not self.stop_evt.is_set():
                self.sm.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                try:
                    self.power.append(nv.nvmlDeviceGetPowerUsage(h) / 1000.0)
                except Exception:
                    pass
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
                time.sleep(self.period)
        except Exception as e:  # NVML missing: report that rather than fake numbers
            self.reasons.add(f"nvml_unavailable:{type(e).__name__}")

    def summary(self):
        import statistics

        return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.sm), "power_w_max": max(self.power) if self.power else None}


def _run_arm_subprocess(script: str, extra: list, args, timeout_s: float):
    """runs one baseline arm in a clean child interpreter (its own sys.path / env) and returns (json line | None, failure text)"""
    import subprocess

    cmd = [sys.executable, os.path.join(ROOT, "baseline", script), "--gpus", str(args.gpus), "--steps", str(args.steps), "--warmup",
           str(args.warmup), "--model", args.model, "--num-prompts", str(args.num_prompts), "--seed", str(args.seed), "--max-running",
           str(args.max_running)] + extra
    import signal
    import types

    proc = subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, cwd=ROOT, start_new_session=True)
    try:
        out, err = proc.communicate(timeout=timeout_s)
    except subprocess.TimeoutExpired:
        try:
            os.killpg(proc.pid, signal.SIGKILL)      # the arm's own process group (engine + its TP workers), nothing else
        except Exception:
            pass
        out, err = proc.communicate()
        return None, f"timeout after {timeout_s:.0f}s"
    finally:
        try:
            os.killpg(proc.pid, signal.SIGKILL)      # stragglers of a crashed engine must not hold the GPU for the next arm
        except Exception:
            pass
    p = types.SimpleNamespace(stdout=out or "", stderr=err or "", returncode=proc.returncode)
    for ln in reversed(p.stdout.splitlines()):
        ln = ln.strip()
        if ln.startswith("{") and '"metric"' in ln:
            try:
                return json.loads(ln), ""
            except Exception:
                pass
    tail = (p.stderr or p.stdout or "").strip().splitlines()
    why = next((t for t in reversed(tail) if "Error" in t or "error" in t or "Exception" in t), tail[-1] if tail else f"rc={p.returncode}")
    log = os.path.join(ROOT, "gpurun_out", f"baseline_{script.replace('.py', '')}_{'_'.join(extra).replace('--', '') or 'run'}.log")
    try:
        os.makedirs(os.path.dirname(log), exist_ok=True)
        with open(log, "w") as f:
            f.write(p.stdout[-20000:] + "\n=== stderr ===\n" + p.stderr[-60000:])
    except Exception:
        pass
    return None, f"rc={p.returncode}: {why[:240]}"


def run_reference(args) -> None:
    """The unmodified reference from baseline/_ref through its own public API (sglang.Engine): baseline/run_reference.py.  Under torchrun only
    rank 0 drives it (the reference spawns and wires its own TP workers); the other ranks leave."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    ref = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref, "sglang")):
        print(json.dumps({"impl": "reference", "unavailable": "baseline/_ref not installed (pip --no-index --no-deps install of /root/reference/python)"}))
        return
    modes = [m for m in os.environ.get("SGLB_REF_MODES", "stock,fi,fi_nopg,degraded,degraded_nopg").split(",") if m]
    fails = []
    budget = 900.0 + 60.0 * (args.steps + args.warmup)
    deadline = time.time() + 2.0 * budget          # the whole ladder: a box where every mode hangs must not eat the round
    for m in modes:
        left = deadline - time.time()
        if left < 120.0:
            fails.append(f"{m}: skipped (ladder budget spent)")
            continue
        line, why = _run_arm_subprocess("run_reference.py", ["--mode", m], args, min(budget, left))
        if line is not None:
            if fails:
                line["config"]["modes_that_failed"] = fails
            print(json.dumps(line), flush=True)
            return
        fails.append(f"{m}: {why}")
    print(json.dumps({"impl": "reference", "unavailable": "; ".join(fails)[:900]}), flush=True)


def run_vllm(args) -> None:
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    line, why = _run_arm_subprocess("run_vllm.py", [], args, 900.0 + 60.0 * (args.steps + args.warmup))
    print(json.dumps(line if line is not None else {"impl": "vllm", "unavailable": why}), flush=True)


def main() -> None:
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    if args.impl == "vllm":
        run_vllm(args)
        return
    import numpy as np
    import torch
    import torch.distributed as dist

    import sglang_b200 as sgl
    from sglang_b200 import ops

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", rank))
    if args.gpus > 1 and world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} needs torchrun with {args.gpus} ranks (WORLD_SIZE={world})")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device")
    torch.cuda.set_device(local_rank)
    eng = sgl.Engine(model_path=f"dummy:{args.model}", tp_size=args.gpus, page_size=args.page_size,
                     attention_backend=args.attention_backend, gemm_backend=args.gemm_backend, max_running_requests=args.max_running,
                     cuda_graph_max_bs=args.max_running, disable_cuda_graph=args.disable_cuda_graph,
                     disable_overlap_schedule=args.disable_overlap, comm_backend=args.comm_backend, quantization=args.quantization,
                     log_level="warning", base_gpu_id=local_rank, random_seed=args.seed)
    pin, pout = sharegpt_shaped_trace(args.num_prompts, args.seed)
    rng = np.random.default_rng(args.seed + 1)
    V = eng.model_config.vocab_size
    tok = eng.tokenizer
    sps = [{"max_new_tokens": m, "temperature": 0, "ignore_eos": True} for m in pout]
    total_out = int(sum(pout))
    total_in = int(sum(pin))

    def fresh_ids(step: int):
        r = np.random.default_rng(args.seed + 100 + step)   # new content every step: no cross-step radix hits
        return [r.integers(300, V - 1000, size=n).tolist() for n in pin]

    def fresh_text(step: int):
        r = np.random.default_rng(args.seed + 5000 + step)
        return ["".join(chr(c) for c in r.integers(97, 123, size=max(1, n - 1))) for n in pin]   # n-1 bytes + BOS = n tokens

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    stream = eng.runner.stream
    # ------------------------------------------------------------------ device-timed region (pre-tokenised ids)
    for w in range(args.warmup):
        eng.generate(input_ids=fresh_ids(-1 - w), sampling_params=sps)
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    c0, g0 = ops.launch_count(), eng.runner.graph_launches
    h0, d0 = eng.runner.h2d_bytes, eng.runner.d2h_bytes
    f0 = eng.runner.forward_ct
    from sglang_b200.model_executor.model_runner import BatchResult

    w0, sw0, sc0 = BatchResult.wait_s, eng.scheduler.step_wall_s, eng.scheduler.step_calls
    t_wall0 = time.perf_counter()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for k in range(args.steps):
        eng.generate(input_ids=fresh_ids(k), sampling_params=sps)
    ev1.record(stream)
    barrier()
    dev_ms = ev0.elapsed_time(ev1)
    wall_ms = 1000.0 * (time.perf_counter() - t_wall0)
    n_sched = max(1, eng.scheduler.step_calls - sc0)
    # where a scheduler iteration goes: host = Python time inside Scheduler.step() that was NOT spent blocked on a CUDA event; the device side is
    # the whole timed region / forward steps (graph replays back to back: a gap shows up as device_ms_per_step > the kernels' own time)
    host_split = {"scheduler_iterations": n_sched, "host_ms_per_iteration": 1000.0 * ((eng.scheduler.step_wall_s - sw0) - (BatchResult.wait_s - w0)) / n_sched,
                  "blocked_on_device_ms_per_iteration": 1000.0 * (BatchResult.wait_s - w0) / n_sched,
                  "outside_scheduler_ms_per_iteration": (wall_ms - 1000.0 * (eng.scheduler.step_wall_s - sw0)) / n_sched}
    launches = (ops.launch_count() - c0) + (eng.runner.graph_launches - g0)
    fwd = eng.runner.forward_ct - f0
    h2d = (eng.runner.h2d_bytes - h0) / max(1, fwd)
    d2h = (eng.runner.d2h_bytes - d0) / max(1, fwd)
    # ------------------------------------------------------------------ end-to-end region (public text API)
    ttfts = []
    for w in range(min(1, args.warmup)):
        eng.generate(fresh_text(-1 - w), sps)
    barrier()
    t0 = time.perf_counter()
    for k in range(args.steps):
        outs = eng.generate(fresh_text(k), sps)
        if rank == 0:
            ttfts.extend(o["meta_info"].get("ttft", 0.0) for o in outs)
            assert all(isinstance(o["text"], str) for o in outs)
    barrier()
    e2e_s = time.perf_counter() - t0
    sampler.stop_evt.set()
    sampler.join(timeout=2)

    t = torch.tensor([dev_ms, e2e_s * 1000.0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms = float(t[0]), float(t[1])
    if rank == 0:
        value = total_out * args.steps / (dev_ms / 1000.0)
        e2e_value = total_out * args.steps / (e2e_ms / 1000.0)
        import statistics

        print(json.dumps({
            "metric": "output_tokens_per_s", "value": value, "unit": "tok/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "bf16" if not args.quantization else args.quantization, "data": "synthetic",
            "config": {"model": args.model, "global_batch": args.num_prompts, "seq_len": f"sharegpt-shaped lognormal: mean in {total_in // args.num_prompts} / out {total_out // args.num_prompts}",
                       "parallelism": f"tp{args.gpus}", "step": "one full trace (request-rate inf, continuous batching)",
                       "l2": "inputs_larger_than_L2 (16 GB weights + KV streamed every decode step)",
                       "attention_backend": eng.server_args.attention_backend, "gemm_backend": os.environ.get("SGLB_GEMM_IMPL", "auto"),
                       "cuda_graph": not args.disable_cuda_graph, "overlap_schedule": eng.scheduler.enable_overlap,
                       "page_size": args.page_size, "max_running": args.max_running, "input_tokens_per_step": total_in,
                       "output_tokens_per_step": total_out, "total_tokens_per_s": (total_in + total_out) * args.steps / (dev_ms / 1000.0)},
            "p50_ttft_ms": 1000.0 * statistics.median(ttfts) if ttfts else None,
            "e2e": {"value": e2e_value, "unit": "tok/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "note": "text prompts -> tokenizer -> engine -> detokenizer -> text; bytes are per engine forward step"},
            "gpu_launches": int(launches), "step_split": {**host_split, "device_ms_per_forward": dev_ms / max(1, fwd), "forward_steps": int(fwd)},
            "clocks": sampler.summary(), "impl": "ours"}))
    eng.shutdown()
    sys.stdout.flush()
    if world > 1:
        dist.barrier()
        # NCCL communicator teardown can block for minutes when CUDA graphs still reference the communicator's streams;
        # the result line is out and every rank has passed the barrier, so leave without running the destructors.
        sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
