#!/bin/bash
# round 2, run N (8 GPUs): TP=8 under the driver's command (size-based AR dispatch, vocab-parallel greedy graphs), TP=8 parity gate
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONPATH=$PWD
echo "== bench tp8"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29638 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/bench_r2n_tp8.json 2> gpurun_out/bench_r2n_tp8.err; echo "rc=$?"
grep '"metric"' gpurun_out/bench_r2n_tp8.json | cut -c1-2600; grep -i "mbarrier\|error\|Traceback\|timeout" gpurun_out/bench_r2n_tp8.err | head -10; tail -3 gpurun_out/bench_r2n_tp8.err | cut -c1-300
echo "== parity gate tp8"
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -x -q -k "test_tp_matches_tp1_and_collectives and 8" 2>&1 | tail -6
grep -i "^allreduce\|^fused ar\|rs+norm+ag {\|SECTION\|FAILED" gpurun_out/mgpu_rank0.log | cut -c1-1500 | head -20
