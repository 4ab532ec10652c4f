#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -25 | tee gpurun_out/pytest_gpu8.log
timeout 300 python tools/gemm_decode_bench.py 2>&1 | grep -v "^W0" | tee gpurun_out/gemm_decode_bench.log
for cfg in "256 512" "64 1024"; do set -- $cfg
  timeout 300 python tools/decode_breakdown.py --bs $1 --ctx $2 --out 40 --gemm auto --tag sk_bs$1_ctx$2 2>&1 | grep -v "^W0\|Warn" | tail -24
done
echo "== bench 1024 prompts, max-running 1024 (skinny on)" | tee gpurun_out/bench8.log
timeout 900 python bench.py --steps 2 --warmup 3 --num-prompts 1024 --max-running 1024 2>&1 | tail -1 | tee -a gpurun_out/bench8.log
