#!/bin/bash
# round 2, run A: attention hand-off fix — stress (fixed + legacy protocol), GPU tests, smoke
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONPATH=$PWD
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv | tee gpurun_out/r2a_gpu.txt
timeout 400 python tools/attn_stress.py --iters 60000 --out gpurun_out/attn_stress_fixed.json > gpurun_out/attn_stress_fixed.log 2>&1
echo "stress fixed rc=$?" | tee -a gpurun_out/attn_stress_fixed.log
SGLB_ATTN_LEGACY_EPILOGUE=1 timeout 400 python tools/attn_stress.py --iters 60000 --out gpurun_out/attn_stress_legacy.json > gpurun_out/attn_stress_legacy.log 2>&1
echo "stress legacy rc=$?" | tee -a gpurun_out/attn_stress_legacy.log
tail -8 gpurun_out/attn_stress_fixed.log; tail -12 gpurun_out/attn_stress_legacy.log
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/pytest_gpu_r2a.log
timeout 120 python __graft_entry__.py smoke 2>&1 | grep -v "^W0\|sglb build" | tail -3 | tee gpurun_out/smoke_r2a.log
