#!/bin/bash
cd "$(dirname "$0")/.."
N=${1:-2}
mkdir -p gpurun_out
export PYTHONPATH=$PWD NCCL_DEBUG=WARN
SGLB_HANG_DUMP_S=200 timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 tools/comm_bench.py > gpurun_out/comm_bench_$N.log 2>&1; echo "comm rc=$?"
grep -v "^W0\|^\s*$\|Warn\|warn" gpurun_out/comm_bench_$N.log | tail -20 | cut -c1-300
