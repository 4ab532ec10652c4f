#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 200 python tools/debug/swa_gpu.py 2>&1 | grep -v "Warning\|warnings.warn" | tail -70 | cut -c1-220
