#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/pytest_gpu_final.log
timeout 120 python __graft_entry__.py smoke 2>&1 | grep -v "^W0\|sglb build" | tail -4 | tee gpurun_out/smoke_final.log
