#!/bin/bash
# 8-GPU validation: collectives (numerics + graph-timed latency), TP=8 engine vs TP=1, comm microbench, TP=8 headline bench
cd "$(dirname "$0")/.."
N=${1:-8}
mkdir -p gpurun_out
export PYTHONPATH=$PWD NCCL_DEBUG=WARN MGPU_MODEL=tiny-llama-gpu8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
SGLB_HANG_DUMP_S=130 timeout 150 python tools/mgpu_ref.py gpurun_out/tp1_ref8.json > gpurun_out/mgpu_ref8.log 2>&1; echo "ref rc=$?"
SGLB_HANG_DUMP_S=280 timeout 300 $TR --master-port 29511 tools/mgpu_test.py gpurun_out/tp1_ref8.json > gpurun_out/mgpu_$N.log 2>&1; echo "mgpu rc=$?"
grep -E "SECTION|MGPU|TP engine" gpurun_out/mgpu_$N.log | cut -c1-260 | head -12
SGLB_HANG_DUMP_S=180 timeout 200 $TR --master-port 29521 tools/comm_bench.py > gpurun_out/comm_bench_$N.log 2>&1; echo "comm rc=$?"
grep -E "^allreduce|^fused|COMM" gpurun_out/comm_bench_$N.log | cut -c1-300
SGLB_HANG_DUMP_S=380 timeout 400 $TR --master-port 29512 bench.py --gpus $N --steps 1 --warmup 3 > gpurun_out/bench_tp$N.log 2>&1; echo "bench rc=$?"
tail -1 gpurun_out/bench_tp$N.log | cut -c1-600
