#!/bin/bash
# round 2, run M (4 GPUs): reduce-scatter + norm + all-gather kernel (numerics + timing in the parity gate), TP=4 bench with it
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONPATH=$PWD
echo "== parity gate tp4 (incl. rs_norm_ag section)"
timeout 700 python -m pytest tests/test_multi_gpu.py -m gpu -x -q -k "test_tp_matches_tp1_and_collectives and 4" 2>&1 | tail -12
grep -i "rs+norm\|SECTION\|FAILED" gpurun_out/mgpu_rank0.log | head -30
echo "== bench tp4 (row-sharded residual on)"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29624 bench.py --gpus 4 --steps 3 --warmup 3 > gpurun_out/bench_r2m_tp4.json 2> gpurun_out/bench_r2m_tp4.err; echo "rc=$?"
grep '"metric"' gpurun_out/bench_r2m_tp4.json | cut -c1-2600; grep -i "mbarrier\|error\|Traceback\|timeout" gpurun_out/bench_r2m_tp4.err | head -10; tail -3 gpurun_out/bench_r2m_tp4.err | cut -c1-300
