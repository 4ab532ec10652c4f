#!/bin/bash
# round 2, run P (1 GPU): the GPU feature tier (tree speculation, Mamba kernels + snapshot cache, sliding pool, multi-process front end)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 1500 python -m pytest tests/test_features_gpu.py -m gpu -q 2>&1 | tail -40
