#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 600 python tools/fp8_moe_bench.py 2>&1 | grep -v "^W0" | tee gpurun_out/fp8_moe_bench.log | tail -40
for k in skinny skinny_silu fp8 moe attn_decode_persist; do
  timeout 240 ncu --set full --clock-control none --import-source on -k regex:"skinny|fp8_block|gemm_tc|persist" -c 2 -o gpurun_out/ncu_$k -f python tools/ncu_target.py $k 3 > gpurun_out/ncu_$k.log 2>&1
  echo "ncu $k rc=$?"
done
ls -la gpurun_out/*.ncu-rep | tail -8
timeout 400 python tools/trace_timeline.py --tag r13 2>&1 | grep -v "^W0" | tail -40
