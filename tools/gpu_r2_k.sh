#!/bin/bash
# round 2, run K (1 GPU): two-producer decode kernel numerics + A/B + cycle accounting, prefill-graph parity, online serving latency
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONPATH=$PWD
echo "== decode v3 numerics"; timeout 400 python tests/gpu_cases.py attn_decode_v3 2>&1 | tail -3
echo "== attn decode bench v3"; timeout 400 python tools/attn_decode_bench.py v3 2>&1 | grep -v Warning | tail -12
echo "== decode attention prof"; for c in bs256_ctx512 bs1024_ragged; do timeout 200 python tools/attn_decode_bench.py v3 $c prof 2>&1 | grep -v Warning | tail -6; done
echo "== prefill graph test"; timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "prefill_cuda_graphs" 2>&1 | tail -5
echo "== online serving (multi-process frontend, prefill graphs)"; timeout 900 bash tools/online_bench.sh --enable-multi-process-frontend --prefill-cuda-graph-max-tokens 2048 2>&1 | grep -E "server up|ttft|itl|tpot|throughput|completed|failed|died|Error" | head -40
