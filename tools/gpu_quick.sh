#!/bin/bash
# short GPU validation of the serving hot path: engine parity (4 configs) + extension-loaded check + smoke()
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 110 python -m pytest tests -m gpu -x -q -k "engine_gpu_matches_reference_path or native_extension" 2>&1 | tail -4 | tee gpurun_out/pytest_gpu_quick.log
timeout 40 python __graft_entry__.py smoke 2>&1 | grep -v "^W0\|sglb build" | tail -3 | tee gpurun_out/smoke_quick.log
