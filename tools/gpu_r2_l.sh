#!/bin/bash
# round 2, run L (1 GPU): decode attention with the L2 prefetcher warp (A/B over the look-ahead), then the TP=1 headline bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONPATH=$PWD
echo "== decode v3 numerics (prefetch on)"; timeout 400 python tests/gpu_cases.py attn_decode_v3 2>&1 | tail -2
for pf in 0 148 222 600; do
  echo "== attn decode bench v3, SGLB_ATTN_PREFETCH=$pf"; SGLB_ATTN_PREFETCH=$pf timeout 300 python tools/attn_decode_bench.py v3 2>&1 | grep -v Warning | tail -11 | cut -c1-120
  cp gpurun_out/attn_decode_bench_v3.json gpurun_out/attn_decode_bench_v3_pf$pf.json
done
echo "== prof (prefetch 222)"; timeout 200 python tools/attn_decode_bench.py v3 bs256_ctx512 prof 2>&1 | grep -v Warning | tail -6
echo "== ours tp1"
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r2l_ours.json 2> gpurun_out/bench_r2l_ours.err; echo "rc=$?"; cat gpurun_out/bench_r2l_ours.json | cut -c1-2600
