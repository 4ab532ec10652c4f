#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 500 python -m pytest tests/test_features_gpu.py -m gpu -q -k "mamba" 2>&1 | grep -v "Warning\|warnings.warn" | tail -30 | cut -c1-300
