#!/bin/bash
# usage: tools/gpu_run_mgpu.sh N   — every stage has a tight timeout and writes its FULL log to gpurun_out/
cd "$(dirname "$0")/.."
N=${1:-2}
mkdir -p gpurun_out
export PYTHONPATH=$PWD
export SGLB_HANG_DUMP_S=240 NCCL_DEBUG=WARN TORCH_NCCL_ASYNC_ERROR_HANDLING=1
timeout 200 python tools/mgpu_ref.py gpurun_out/tp1_ref.json > gpurun_out/mgpu_ref.log 2>&1; echo "ref rc=$?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/mgpu_test.py gpurun_out/tp1_ref.json > gpurun_out/mgpu_$N.log 2>&1; echo "mgpu rc=$?"
grep -E "ok|OK|Error|error|assert|Traceback|rank" gpurun_out/mgpu_$N.log | head -40
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 1 --warmup 3 --num-prompts 256 > gpurun_out/bench_tp$N.log 2>&1; echo "bench rc=$?"
tail -5 gpurun_out/bench_tp$N.log | cut -c1-1200
