#!/bin/bash
# round 2, run B: large attention stress (legacy protocol should reproduce the r1 deadlock, fixed must not), then the baseline arms (probe)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONPATH=$PWD
SGLB_ATTN_LEGACY_EPILOGUE=1 timeout 300 python tools/attn_stress.py --iters 60000 --out gpurun_out/attn_stress_legacy.json > gpurun_out/attn_stress_legacy.log 2>&1
echo "stress legacy rc=$?" | tee -a gpurun_out/attn_stress_legacy.log
timeout 300 python tools/attn_stress.py --iters 60000 --out gpurun_out/attn_stress_fixed.json > gpurun_out/attn_stress_fixed.log 2>&1
echo "stress fixed rc=$?" | tee -a gpurun_out/attn_stress_fixed.log
tail -12 gpurun_out/attn_stress_legacy.log | cut -c1-400; tail -4 gpurun_out/attn_stress_fixed.log | cut -c1-400
# baseline arms, short: does each one run at all on this box?
SGLB_REF_MODES=stock,degraded timeout 1500 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/ref_probe.json 2> gpurun_out/ref_probe.err
echo "reference arm rc=$?"; cat gpurun_out/ref_probe.json | cut -c1-1500
timeout 900 python bench.py --impl vllm --steps 1 --warmup 1 > gpurun_out/vllm_probe.json 2> gpurun_out/vllm_probe.err
echo "vllm arm rc=$?"; cat gpurun_out/vllm_probe.json | cut -c1-1500
du -sh baseline/_ref_jit_cache baseline/_fi_ws 2>/dev/null
mkdir -p gpurun_out/caches
tar czf gpurun_out/caches/ref_jit_cache.tgz -C baseline _ref_jit_cache 2>/dev/null
if [ -d baseline/_fi_ws ]; then
  find baseline/_fi_ws -name "*.o" -delete 2>/dev/null
  sz=$(du -sm baseline/_fi_ws | cut -f1); echo "fi_ws MB (no .o): $sz"
  if [ "$sz" -lt 45 ]; then tar czf gpurun_out/caches/fi_ws.tgz -C baseline _fi_ws; fi
fi
ls -la gpurun_out/caches
