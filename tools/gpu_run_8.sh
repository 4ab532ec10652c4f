#!/bin/bash
# 8-GPU validation: collectives (numerics + graph-timed latency), TP=8 engine vs TP=1, TP=8 headline bench.  Tight stage timeouts:
# an 8-GPU call is charged 8x, a hang must not eat the round's budget.
cd "$(dirname "$0")/.."
N=${1:-8}
mkdir -p gpurun_out
export PYTHONPATH=$PWD NCCL_DEBUG=WARN MGPU_MODEL=tiny-llama-gpu8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
SGLB_HANG_DUMP_S=80 timeout 100 python tools/mgpu_ref.py gpurun_out/tp1_ref8.json > gpurun_out/mgpu_ref8.log 2>&1; echo "ref rc=$?"
SGLB_HANG_DUMP_S=170 timeout 190 $TR --master-port 29511 tools/mgpu_test.py gpurun_out/tp1_ref8.json > gpurun_out/mgpu_$N.log 2>&1; echo "mgpu rc=$?"
grep -E "SECTION|MGPU|TP engine" gpurun_out/mgpu_$N.log | cut -c1-260 | head -12
SGLB_HANG_DUMP_S=230 timeout 250 $TR --master-port 29512 bench.py --gpus $N --steps 1 --warmup 3 > gpurun_out/bench_tp$N.log 2>&1; echo "bench rc=$?"
tail -1 gpurun_out/bench_tp$N.log | cut -c1-600
