#!/bin/bash
# round 2, run F: two-softmax-group decode kernel (numerics + A/B), LoRA SGMV case, reference arm with bounded prefill graphs, ours
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONPATH=$PWD
echo "== decode v3 (2 groups) numerics"; timeout 400 python tests/gpu_cases.py attn_decode_v3 2>&1 | tail -3
echo "== attn decode bench v3"; timeout 300 python tools/attn_decode_bench.py v3 2>&1 | grep -v Warning | tail -10
echo "== lora case"; timeout 400 python tests/gpu_cases.py lora 2>&1 | tail -3
echo "== reference arm (full model)"
SGLB_REF_MODES=fi,fi_nopg,degraded_nopg timeout 2400 python bench.py --impl reference --steps 2 --warmup 2 > gpurun_out/bench_r2f_ref.json 2> gpurun_out/bench_r2f_ref.err; echo "rc=$?"; cat gpurun_out/bench_r2f_ref.json | cut -c1-3000
echo "== ours tp1"
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r2f_ours.json 2> gpurun_out/bench_r2f_ours.err; echo "rc=$?"; cat gpurun_out/bench_r2f_ours.json | cut -c1-2600
mkdir -p gpurun_out/caches
tar czf gpurun_out/caches/ref_jit_cache.tgz -C baseline _ref_jit_cache 2>/dev/null
find baseline/_fi_ws -name "*.o" -delete 2>/dev/null; sz=$(du -sm baseline/_fi_ws | cut -f1); echo "fi_ws MB: $sz"
if [ "$sz" -lt 45 ]; then tar czf gpurun_out/caches/fi_ws.tgz -C baseline _fi_ws; fi
