#!/bin/bash
# round 2, run G: one-group decode v3 + layout probes, sampler / kvio cases, full GPU tier, smoke
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONPATH=$PWD
echo "== decode v3 numerics"; timeout 400 python tests/gpu_cases.py attn_decode_v3 2>&1 | tail -2
echo "== attn decode bench v3"; timeout 400 python tools/attn_decode_bench.py v3 2>&1 | grep -v Warning | tail -12
echo "== attn decode bench fi probes"; timeout 400 python tools/attn_decode_bench.py fi probe 2>&1 | grep -v Warning | tail -3
echo "== sampling case"; timeout 400 python tests/gpu_cases.py sampling 2>&1 | tail -3
echo "== kvio case"; timeout 400 python tests/gpu_cases.py kvio 2>&1 | tail -3
echo "== full gpu tier"; timeout 2400 python -m pytest tests -m gpu -x -q --deselect tests/test_multi_gpu.py 2>&1 | tail -8
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
