#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONPATH=$PWD
python tools/gpu_diag.py attn_tc_decode attn_tc_prefill attn_tc_split attn_tc_window 2>&1 | tee gpurun_out/diag3.log
for t in attn_decode attn_prefill gemm; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"attn_fwd_tc|gemm_tc" -s 2 -c 1 -f -o gpurun_out/ncu_$t python tools/ncu_target.py $t 4 > gpurun_out/ncu_$t.log 2>&1
done
python - <<'PY' 2>&1 | tee gpurun_out/attn_micro3.log
import sys, os
sys.path.insert(0, "tools"); sys.argv=["x"]
import json, torch
sys.path.insert(0, "tests")
from sglang_b200 import ops
from gpu_cases import _attn_setup
ops.load_native()
def t(fn, it=5):
    for _ in range(3): fn()
    ts=[]
    for _ in range(it):
        s,e=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); s.record(); fn(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
    return sorted(ts)[len(ts)//2]
for tag,B,ql,sl,dec in [("decode_bs256_ctx512",256,[1]*256,[512]*256,True),("decode_bs64_ctx2048",64,[1]*64,[2048]*64,True),("prefill_4x2048",4,[2048]*4,[2048]*4,False)]:
    s=_attn_setup(B,ql,sl,32,8,128,32); out=torch.empty_like(s["q"])
    ms=t(lambda: ops.paged_attention(s["q"],s["kc"],s["vc"],s["pt"],s["slots"],s["seq"],None if dec else s["cu"],None if dec else s["ql"],max(ql),32,128**-0.5,impl=0,out=out))
    print(tag, round(ms,4), "ms", round(sum(sl)*8*128*4/ms/1e6,1), "GB/s kv")
PY
