#!/bin/bash
# round 2, run H (4 GPUs): TMA gather micro-benchmark, vocab-parallel greedy case, TP=4 bench under the driver's command, TP=4 parity gate
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONPATH=$PWD
echo "== tma gather bench Hkv=8"; timeout 120 tools/bin/tma_gather_bench 8 2>&1 | tail -22
echo "== tma gather bench Hkv=1"; timeout 120 tools/bin/tma_gather_bench 1 2>&1 | tail -22
echo "== sampling case (greedy shard)"; timeout 300 python tests/gpu_cases.py sampling 2>&1 | tail -2
echo "== bench tp4"
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29614 bench.py --gpus 4 --steps 6 --warmup 3 > gpurun_out/bench_r2h_tp4.json 2> gpurun_out/bench_r2h_tp4.err; echo "rc=$?"
grep '"metric"' gpurun_out/bench_r2h_tp4.json | cut -c1-2600; grep -i "mbarrier\|error\|Traceback" gpurun_out/bench_r2h_tp4.err | head -10; tail -5 gpurun_out/bench_r2h_tp4.err
echo "== parity gate tp4"
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -x -q -k "test_tp_matches_tp1_and_collectives and 4" 2>&1 | tail -15
grep -i "us\b\|latency\|allreduce" gpurun_out/mgpu_rank0.log | head -30
