#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu7.log
for cfg in "256 512" "64 1024" "512 400"; do set -- $cfg
  timeout 300 python tools/decode_breakdown.py --bs $1 --ctx $2 --out 40 --tag bs$1_ctx$2 2>&1 | grep -v "^W0\|Warn" | tail -34
done
echo "== bench 1024 prompts, max-running 1024" | tee gpurun_out/bench7.log
timeout 900 python bench.py --steps 2 --warmup 3 --num-prompts 1024 --max-running 1024 2>&1 | tail -1 | tee -a gpurun_out/bench7.log
