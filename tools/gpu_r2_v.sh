#!/bin/bash
# round 2, run V (1 GPU): last sanity pass on the final tree — smoke(), the feature cases that create two engines, a short bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONPATH=$PWD
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
echo "== feature cases"; timeout 400 python -m pytest tests/test_features_gpu.py -m gpu -q -k "tree or multi_process" 2>&1 | tail -4
echo "== bench (1 step)"; timeout 600 python bench.py --steps 1 --warmup 3 > gpurun_out/bench_r2v_ours.json 2> gpurun_out/bench_r2v_ours.err; echo "rc=$?"; cut -c1-400 gpurun_out/bench_r2v_ours.json; tail -2 gpurun_out/bench_r2v_ours.err | cut -c1-300
