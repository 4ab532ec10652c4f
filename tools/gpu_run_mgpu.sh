#!/bin/bash
# usage: tools/gpu_run_mgpu.sh N   — every stage has a tight timeout and writes its FULL log to gpurun_out/
cd "$(dirname "$0")/.."
N=${1:-2}
mkdir -p gpurun_out
export PYTHONPATH=$PWD
export NCCL_DEBUG=WARN TORCH_NCCL_ASYNC_ERROR_HANDLING=1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
SGLB_HANG_DUMP_S=180 timeout 200 python tools/mgpu_ref.py gpurun_out/tp1_ref.json > gpurun_out/mgpu_ref.log 2>&1; echo "ref rc=$?"
SGLB_HANG_DUMP_S=300 timeout 360 $TR --master-port 29511 tools/mgpu_test.py gpurun_out/tp1_ref.json > gpurun_out/mgpu_$N.log 2>&1; echo "mgpu rc=$?"
grep -E "SECTION|MGPU|TP engine" gpurun_out/mgpu_$N.log | cut -c1-300 | head -20
SGLB_HANG_DUMP_S=200 timeout 240 $TR --master-port 29513 tools/decode_breakdown.py --bs 256 --ctx 512 --out 40 --gemm auto --tag tp${N}_bs256 > gpurun_out/breakdown_tp$N.log 2>&1; echo "breakdown rc=$?"
head -14 gpurun_out/breakdown_tp${N}_bs256.txt 2>/dev/null | cut -c1-170
for mode in push pull; do
SGLB_FUSED_AR=$mode SGLB_HANG_DUMP_S=400 timeout 460 $TR --master-port 29512 bench.py --gpus $N --steps 1 --warmup 3 --num-prompts 1024 --max-running 1024 > gpurun_out/bench_tp${N}_$mode.log 2>&1; echo "bench $mode rc=$?"
tail -1 gpurun_out/bench_tp${N}_$mode.log | cut -c1-400
done
