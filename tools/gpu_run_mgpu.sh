#!/bin/bash
# usage: tools/gpu_run_mgpu.sh N
cd "$(dirname "$0")/.."
N=${1:-2}
mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 300 python tools/mgpu_ref.py gpurun_out/tp1_ref.json 2>&1 | tail -2
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/mgpu_test.py gpurun_out/tp1_ref.json 2>&1 | tail -30 | tee gpurun_out/mgpu_$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 1 --warmup 3 --num-prompts 256 2>&1 | tail -3 | tee gpurun_out/bench_tp$N.log
