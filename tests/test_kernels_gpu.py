"""GPU tier: kernel numerics vs the fp32 torch reference + GPU engine end-to-end.  Each case runs in its own process
(see tests/gpu_cases.py) under a timeout so an experimental-kernel trap cannot take the other cases down."""
import json
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CASES = ["rmsnorm", "act", "rope", "misc", "sampling", "attn_simt", "gemm_tc", "gemm_split_silu", "gemm_skinny", "moe", "fp8", "attn_tc_decode", "attn_tc_prefill",
         "attn_tc_split", "attn_tc_persist", "attn_tc_window"]


def run_case(name: str, timeout: int = 300):
    env = dict(os.environ, PYTHONPATH=ROOT)
    p = subprocess.run([sys.executable, os.path.join(HERE, "gpu_cases.py"), name], capture_output=True, text=True, timeout=timeout,
                       env=env, cwd=ROOT)
    return p.returncode, p.stdout[-3000:], p.stderr[-4000:]


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_kernel_case(name):
    rc, out, err = run_case(name)
    assert rc == 0, f"case {name} failed\nSTDOUT:\n{out}\nSTDERR:\n{err}"
    assert json.loads(out.strip().splitlines()[-1])["ok"]


@pytest.mark.gpu
def test_native_extension_is_loaded():
    import torch

    from sglang_b200 import ops

    assert ops.load_native(), "the CUDA extension must load on a GPU box (no silent torch fallback)"
    x = torch.randn(4, 1024, device="cuda", dtype=torch.bfloat16)
    c0 = ops.launch_count()
    ops.rmsnorm(x, torch.ones(1024, device="cuda", dtype=torch.bfloat16), 1e-5)
    assert ops.launch_count() == c0 + 1


@pytest.mark.gpu
@pytest.mark.parametrize("attn,gemm,graph", [("simt", "lib", False), ("tc", "lib", True), ("tc", "tc", True)])
def test_engine_gpu_matches_reference_path(attn, gemm, graph):
    """greedy generations of the tcgen05 path == the reference-grade SIMT/cuBLAS path on the same weights"""
    code = f'''
import json, sys
import sglang_b200 as sgl
import numpy as np
rng = np.random.default_rng(0)
prompts = [rng.integers(3, 2000, size=int(n)).tolist() for n in rng.integers(5, 300, size=12)]
eng = sgl.Engine(model_path="dummy:tiny-llama-gpu", attention_backend="{attn}", gemm_backend="{gemm}", disable_cuda_graph={not graph},
                 page_size=32, chunked_prefill_size=256, cuda_graph_max_bs=16, mem_fraction_static=0.2, max_running_requests=16)
out = eng.generate(input_ids=prompts, sampling_params={{"max_new_tokens": 24, "temperature": 0, "ignore_eos": True}})
eng.scheduler.check_invariants()
print(json.dumps([o["output_ids"] for o in out]))
'''
    outs = {}
    for tag, c in (("test", code), ("ref", code.replace(f'attention_backend="{attn}"', 'attention_backend="simt"').replace(
            f'gemm_backend="{gemm}"', 'gemm_backend="lib"').replace(f"disable_cuda_graph={not graph}", "disable_cuda_graph=True"))):
        p = subprocess.run([sys.executable, "-c", c], capture_output=True, text=True, timeout=600, env=dict(os.environ, PYTHONPATH=ROOT), cwd=ROOT)
        assert p.returncode == 0, p.stderr[-4000:]
        outs[tag] = json.loads(p.stdout.strip().splitlines()[-1])
    same = sum(a == b for a, b in zip(outs["test"], outs["ref"]))
    # bf16 reductions in a different order may flip a near-tie late in a random-weight model: demand the large majority
    first = sum(a[:4] == b[:4] for a, b in zip(outs["test"], outs["ref"]))
    assert first == len(outs["ref"]) and same >= len(outs["ref"]) - 3, (same, first)
