#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 420 python -m pytest tests -m gpu -q 2>&1 | tail -12 | tee gpurun_out/pytest_gpu13.log
timeout 200 python tools/fp8_moe_bench.py 2>&1 | grep -v "^W0" | tee gpurun_out/fp8_moe_bench.log | tail -30
for k in fp8 moe skinny; do
  timeout 100 ncu --set full --clock-control none --import-source on -k regex:"skinny|fp8_block|gemm_tc" -c 1 -o gpurun_out/ncu_$k -f python tools/ncu_target.py $k 2 > gpurun_out/ncu_$k.log 2>&1
  echo "ncu $k rc=$?"
done
ls -la gpurun_out/*.ncu-rep | tail -5
