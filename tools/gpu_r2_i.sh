#!/bin/bash
# round 2, run I (1 GPU): decode-attention per-role cycle accounting, prefill-graph parity test, online serving latency
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONPATH=$PWD
echo "== decode attention prof"; for c in bs256_ctx512 bs1024_ragged bs64_ctx2048; do timeout 200 python tools/attn_decode_bench.py v3 $c prof 2>&1 | grep -v Warning | tail -5; done
echo "== prefill graph test"; timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "prefill_cuda_graphs" 2>&1 | tail -5
echo "== online serving (multi-process frontend, prefill graphs)"; timeout 900 bash tools/online_bench.sh --enable-multi-process-frontend --prefill-cuda-graph-max-tokens 2048 2>&1 | grep -E "server up|ttft|itl|tpot|throughput|completed|failed|died|Error" | head -40
