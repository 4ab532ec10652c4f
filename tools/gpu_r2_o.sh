#!/bin/bash
# round 2, run O (1 GPU): the whole GPU tier on the final tree, smoke(), then both bench arms back to back (reference first, as the driver does)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONPATH=$PWD
echo "== full gpu tier"; timeout 2400 python -m pytest tests -m gpu -q --deselect tests/test_multi_gpu.py 2>&1 | tail -15
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
echo "== reference arm"
SGLB_REF_MODES=fi,fi_nopg,degraded_nopg timeout 1500 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/bench_r2o_ref.json 2> gpurun_out/bench_r2o_ref.err; echo "rc=$?"; cut -c1-1500 gpurun_out/bench_r2o_ref.json
echo "== ours tp1"
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r2o_ours.json 2> gpurun_out/bench_r2o_ours.err; echo "rc=$?"; cut -c1-2600 gpurun_out/bench_r2o_ours.json
