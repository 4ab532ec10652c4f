#!/bin/bash
# round 2, run Q (1 GPU): the three feature cases that failed in run P, after the fixes (full failure output kept)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 900 python -m pytest tests/test_features_gpu.py -m gpu -q -k "mamba2 or sliding" 2>&1 | grep -v "Warning\|warnings.warn" | tail -120 > gpurun_out/r2q_pytest.log; tail -60 gpurun_out/r2q_pytest.log
