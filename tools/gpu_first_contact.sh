#!/bin/bash
# first GPU contact: kernel cases, tiny engine in 3 configs, then the flagship bench with the safe and the tcgen05 paths
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONPATH=$PWD
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
python tools/gpu_diag.py 2>&1 | tee gpurun_out/diag.log
for cfg in "simt lib" "tc lib" "tc tc"; do
  set -- $cfg
  echo "== smoke attn=$1 gemm=$2" | tee -a gpurun_out/smoke.log
  SGLB_SMOKE_ATTN=$1 SGLB_SMOKE_GEMM=$2 timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee -a gpurun_out/smoke.log
done
echo "== bench safe path" | tee gpurun_out/bench.log
timeout 900 python bench.py --steps 1 --warmup 1 --num-prompts 128 --attention-backend simt --gemm-backend lib --disable-cuda-graph 2>&1 | tail -3 | tee -a gpurun_out/bench.log
echo "== bench tc attention + graphs" | tee -a gpurun_out/bench.log
timeout 900 python bench.py --steps 2 --warmup 3 --attention-backend tc --gemm-backend lib 2>&1 | tail -3 | tee -a gpurun_out/bench.log
echo "== bench tc attention + tc gemm + graphs" | tee -a gpurun_out/bench.log
timeout 900 python bench.py --steps 2 --warmup 3 --attention-backend tc --gemm-backend tc 2>&1 | tail -3 | tee -a gpurun_out/bench.log
