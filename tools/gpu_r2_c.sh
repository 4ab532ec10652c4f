#!/bin/bash
# round 2, run C: stream-K GEMM numerics + bench, legacy-protocol hang repro on TP-shard shapes, reference arm probes, trace timeline
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONPATH=$PWD
echo "== streamk numerics"; timeout 300 python tests/gpu_cases.py gemm_streamk 2>&1 | tail -5 | tee gpurun_out/streamk_case.log
echo "== streamk bench"; timeout 300 python tools/gemm_streamk_bench.py 1,16,64,128,256 2>&1 | tail -40 | tee gpurun_out/streamk_bench.log
echo "== decode v3 numerics"; timeout 400 python tests/gpu_cases.py attn_decode_v3 2>&1 | tail -5 | tee gpurun_out/decode_v3_case.log
for t in v2 v3 fi; do echo "== attn decode bench $t"; timeout 300 python tools/attn_decode_bench.py $t 2>&1 | grep -v Warning | tail -12; done
echo "== legacy repro tp8shard"
SGLB_STREAMK_MAX_M=0 SGLB_ATTN_DECODE=v2 SGLB_ATTN_LEGACY_EPILOGUE=1 timeout 400 python bench.py --model llama-3-8b-tp8shard --steps 10 --warmup 5 > gpurun_out/legacy_tp8shard.json 2> gpurun_out/legacy_tp8shard.err
echo "rc=$?"; tail -c 600 gpurun_out/legacy_tp8shard.json; grep -a "sglb:" gpurun_out/legacy_tp8shard.err | head -20
echo "== fixed (v3 decode) tp8shard"
SGLB_STREAMK_MAX_M=0 timeout 400 python bench.py --model llama-3-8b-tp8shard --steps 10 --warmup 5 > gpurun_out/fixed_tp8shard.json 2> gpurun_out/fixed_tp8shard.err
echo "rc=$?"; tail -c 600 gpurun_out/fixed_tp8shard.json
echo "== reference tiny probes"
for m in stock degraded; do
  timeout 600 python baseline/run_reference.py --mode $m --model llama-tiny --num-prompts 64 --steps 1 --warmup 1 > gpurun_out/ref_tiny_$m.out 2> gpurun_out/ref_tiny_$m.err
  echo "ref tiny $m rc=$?"; tail -c 700 gpurun_out/ref_tiny_$m.out; grep -a "Error\|error" gpurun_out/ref_tiny_$m.err | tail -5 | cut -c1-300
done
echo "== timeline v2+cublas"; SGLB_STREAMK_MAX_M=0 SGLB_ATTN_DECODE=v2 timeout 400 python tools/trace_timeline.py --tag r2c_v2_lib 2>&1 | tail -42
echo "== timeline v3+cublas"; SGLB_STREAMK_MAX_M=0 timeout 400 python tools/trace_timeline.py --tag r2c_v3_lib 2>&1 | tail -42
echo "== timeline v3+streamk"; timeout 400 python tools/trace_timeline.py --tag r2c_v3_streamk 2>&1 | tail -42
mkdir -p gpurun_out/caches
tar czf gpurun_out/caches/ref_jit_cache.tgz -C baseline _ref_jit_cache 2>/dev/null
find baseline/_fi_ws -name "*.o" -delete 2>/dev/null; sz=$(du -sm baseline/_fi_ws | cut -f1); echo "fi_ws MB: $sz"
if [ "$sz" -lt 45 ]; then tar czf gpurun_out/caches/fi_ws.tgz -C baseline _fi_ws; fi
