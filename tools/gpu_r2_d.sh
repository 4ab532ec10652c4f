#!/bin/bash
# round 2, run D: decode v3 (double-buffered Q, prefetched queue atomic) numerics + A/B, stream-K fixup MLP, reference arm (full model), ours
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONPATH=$PWD
echo "== decode v3 numerics"; timeout 400 python tests/gpu_cases.py attn_decode_v3 2>&1 | tail -3
echo "== attn decode bench v3"; timeout 300 python tools/attn_decode_bench.py v3 2>&1 | grep -v Warning | tail -10
echo "== streamk numerics + bench"; timeout 300 python tests/gpu_cases.py gemm_streamk 2>&1 | tail -2
timeout 300 python tools/gemm_streamk_bench.py 1,16,32,64 2>&1 | grep "^|" | tee gpurun_out/streamk_bench2.log
echo "== ours tp1 (v3 decode + cuBLAS)"
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r2d_ours.json 2> gpurun_out/bench_r2d_ours.err; echo "rc=$?"; cat gpurun_out/bench_r2d_ours.json | cut -c1-2500
echo "== reference arm (full model)"
SGLB_REF_MODES=fi,degraded timeout 1700 python bench.py --impl reference --steps 2 --warmup 2 > gpurun_out/bench_r2d_ref.json 2> gpurun_out/bench_r2d_ref.err; echo "rc=$?"; cat gpurun_out/bench_r2d_ref.json | cut -c1-2500
echo "== timeline v3+cublas"; timeout 400 python tools/trace_timeline.py --tag r2d_v3_lib 2>&1 | tail -42
mkdir -p gpurun_out/caches
tar czf gpurun_out/caches/ref_jit_cache.tgz -C baseline _ref_jit_cache 2>/dev/null
find baseline/_fi_ws -name "*.o" -delete 2>/dev/null; sz=$(du -sm baseline/_fi_ws | cut -f1); echo "fi_ws MB: $sz"
if [ "$sz" -lt 45 ]; then tar czf gpurun_out/caches/fi_ws.tgz -C baseline _fi_ws; fi
