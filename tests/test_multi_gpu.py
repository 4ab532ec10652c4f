"""Multi-GPU gate (``pytest -m gpu`` on a box with >= 2 GPUs; skipped otherwise).

Round 1 shipped with a manual multi-GPU script only, which is how the TP>=4 mbarrier deadlock went unseen until the driver's scaling run.  This
gate launches exactly what the driver launches — ``python -m torch.distributed.run ... bench.py --gpus N`` for several full traces — plus the NVL
collective / fused-kernel numerics and the TP-vs-TP=1 prompt-logprob parity of ``tools/mgpu_test.py``.

Reference strategy: test/registered/ (2/4/8-GPU suites, ``run_bench_serving`` floors) — SURVEY.md §4.
"""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus() -> int:
    try:
        import torch

        return torch.cuda.device_count() if torch.cuda.is_available() else 0
    except Exception:
        return 0


def _torchrun(n: int, port: int, script: str, *args: str, timeout: int = 900, env=None):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), script, *args]
    e = dict(os.environ, PYTHONPATH=ROOT, SGLB_HANG_DUMP_S=str(timeout - 30), **(env or {}))
    return subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=timeout, env=e, start_new_session=True)


@pytest.mark.parametrize("n", [2, 4, 8])
def test_bench_survives_full_traces(n):
    """the driver's command, >= 3 timed traces after 3 warm-up traces: rc 0 and one well-formed JSON line"""
    if _n_gpus() < n:
        pytest.skip(f"needs {n} GPUs")
    p = _torchrun(n, 29600 + n, "bench.py", "--gpus", str(n), "--steps", "3", "--warmup", "3", timeout=1200)
    tail = (p.stdout + p.stderr)[-3000:]
    assert p.returncode == 0, f"bench.py --gpus {n} failed (rc {p.returncode}):\n{tail}"
    assert "mbarrier wait timeout" not in p.stdout + p.stderr, tail
    lines = [ln for ln in p.stdout.splitlines() if ln.startswith("{") and '"metric"' in ln]
    assert len(lines) == 1, tail
    rec = json.loads(lines[0])
    assert rec["n_gpus"] == n and rec["steps"] == 3 and rec["value"] > 0 and rec["config"]["parallelism"] == f"tp{n}"
    assert rec["gpu_launches"] > 0 and rec["e2e"]["value"] > 0


@pytest.mark.parametrize("n", [2, 4, 8])
def test_tp_matches_tp1_and_collectives(n):
    """NVL collectives / fused GEMM+AR vs NCCL and fp32 references, then TP=n engine vs TP=1 engine (teacher-forced prompt logprobs)"""
    if _n_gpus() < n:
        pytest.skip(f"needs {n} GPUs")
    out = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    ref = os.path.join(out, f"tp1_ref_gate{n}.json")
    env = dict(os.environ, PYTHONPATH=ROOT, MGPU_MODEL="tiny-llama-gpu8")
    r = subprocess.run([sys.executable, "tools/mgpu_ref.py", ref], cwd=ROOT, capture_output=True, text=True, timeout=400, env=env)
    assert r.returncode == 0, (r.stdout + r.stderr)[-2000:]
    p = _torchrun(n, 29700 + n, "tools/mgpu_test.py", ref, timeout=900, env={"MGPU_MODEL": "tiny-llama-gpu8"})
    tail = (p.stdout + p.stderr)[-3000:]
    assert p.returncode == 0 and "MGPU OK" in p.stdout, tail
