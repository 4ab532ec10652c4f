#!/bin/bash
# round 2, run E: reference arm (full model, mem fraction 0.8), ncu captures of the decode attention kernels (ours v3 vs FlashInfer trtllm-gen)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONPATH=$PWD
echo "== reference arm (full model)"
SGLB_REF_MODES=fi,degraded timeout 1700 python bench.py --impl reference --steps 2 --warmup 2 > gpurun_out/bench_r2e_ref.json 2> gpurun_out/bench_r2e_ref.err; echo "rc=$?"; cat gpurun_out/bench_r2e_ref.json | cut -c1-3000
echo "== ncu v3"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:attn_decode_kernel -s 3 -c 1 -f -o gpurun_out/ncu_decode_v3 python tools/attn_decode_bench.py v3 bs1024 > gpurun_out/ncu_v3.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/ncu_v3.log
echo "== ncu fi"
timeout 400 ncu --set full --clock-control none -k regex:fmha -s 3 -c 1 -f -o gpurun_out/ncu_decode_fi python tools/attn_decode_bench.py fi bs1024 > gpurun_out/ncu_fi.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/ncu_fi.log
ls -la gpurun_out/*.ncu-rep
mkdir -p gpurun_out/caches
tar czf gpurun_out/caches/ref_jit_cache.tgz -C baseline _ref_jit_cache 2>/dev/null
find baseline/_fi_ws -name "*.o" -delete 2>/dev/null; sz=$(du -sm baseline/_fi_ws | cut -f1); echo "fi_ws MB: $sz"
if [ "$sz" -lt 45 ]; then tar czf gpurun_out/caches/fi_ws.tgz -C baseline _fi_ws; fi
