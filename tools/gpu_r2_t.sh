#!/bin/bash
# round 2, run T (1 GPU): the complete GPU tier on the final tree (kernel cases, engine parity, prefill graphs, feature tier), smoke()
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_multi_gpu.py 2>&1 | grep -v "Warning\|warnings.warn" > gpurun_out/r2t_pytest.log; tail -25 gpurun_out/r2t_pytest.log | cut -c1-300
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
