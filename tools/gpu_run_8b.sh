#!/bin/bash
# 8-GPU headline bench, pull then push fused-AR mode; each stage bounded (an 8-GPU call is charged 8x)
cd "$(dirname "$0")/.."
N=${1:-8}
mkdir -p gpurun_out
export PYTHONPATH=$PWD NCCL_DEBUG=WARN
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
for mode in pull push; do
SGLB_FUSED_AR=$mode SGLB_HANG_DUMP_S=150 timeout 170 $TR --master-port 2951$( [ $mode = pull ] && echo 2 || echo 3 ) bench.py --gpus $N --steps 1 --warmup 3 > gpurun_out/bench_tp${N}_$mode.log 2>&1; echo "bench $mode rc=$?"
grep -E "^\{|timeout|Timeout \(" gpurun_out/bench_tp${N}_$mode.log | head -4 | cut -c1-700
done
