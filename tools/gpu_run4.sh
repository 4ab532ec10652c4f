#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONPATH=$PWD
python tools/gpu_diag.py attn_tc_decode attn_tc_prefill attn_tc_split attn_tc_window 2>&1 | tee gpurun_out/diag4.log
python - <<'PY' 2>&1 | tee gpurun_out/attn_micro4.log
import sys, os, json, torch
sys.path.insert(0, "tests")
from sglang_b200 import ops
from gpu_cases import _attn_setup
ops.load_native()
peaks=json.load(open("MEASURED_PEAKS.json")) if os.path.exists("MEASURED_PEAKS.json") else {"hbm_gbs":6650,"bf16_tflops":1590}
flush=torch.empty(256<<20,dtype=torch.uint8,device="cuda")
def t(fn, it=7):
    for _ in range(3): fn()
    ts=[]
    for _ in range(it):
        flush.zero_()
        s,e=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); s.record(); fn(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
    return sorted(ts)[len(ts)//2]
import numpy as np
rng=np.random.default_rng(0)
cases=[("decode_bs256_ctx512",256,[1]*256,[512]*256,True,1),("decode_bs256_ragged",256,[1]*256,rng.integers(32,1500,256).tolist(),True,1),
("decode_bs64_ctx2048",64,[1]*64,[2048]*64,True,1),("decode_bs8_ctx4096_s8",8,[1]*8,[4096]*8,True,8),("decode_bs512_ctx512",512,[1]*512,[512]*512,True,1),
("prefill_28seq",28,rng.integers(50,600,28).tolist(),None,False,1),("prefill_4x2048",4,[2048]*4,[2048]*4,False,1),("prefill_1x8192",1,[8192],[8192],False,1)]
res={}
for tag,B,ql,sl,dec,sp in cases:
    sl = sl or ql
    s=_attn_setup(B,ql,sl,32,8,128,32); out=torch.empty_like(s["q"])
    ws=(torch.empty(sp,s["T"],32,128,dtype=torch.float32,device="cuda"),torch.empty(sp,s["T"],32,dtype=torch.float32,device="cuda")) if sp>1 else None
    ms=t(lambda: ops.paged_attention(s["q"],s["kc"],s["vc"],s["pt"],s["slots"],s["seq"],None if dec else s["cu"],None if dec else s["ql"],max(ql),32,128**-0.5,num_splits=sp,workspace=ws,impl=0,out=out))
    kvb=sum(sl)*8*128*4; fl=sum(4.0*32*128*q*(l-q/2) for q,l in zip(ql,sl))
    roof=max(kvb/(peaks["hbm_gbs"]*1e6), fl/(peaks["bf16_tflops"]*1e9))
    res[tag]={"ms":ms,"kv_gbs":kvb/ms/1e6,"tflops":fl/ms/1e9,"roofline_ms":roof,"frac_of_measured_roofline":roof/ms}
    print(tag, {k:round(v,4) for k,v in res[tag].items()}, flush=True)
json.dump(res, open("gpurun_out/attn_micro4.json","w"), indent=1)
PY
echo "== bench 1024 prompts, max-running 512" | tee gpurun_out/bench4.log
timeout 1200 python bench.py --steps 2 --warmup 3 --num-prompts 1024 --max-running 512 --attention-backend tc --gemm-backend lib 2>&1 | tail -2 | tee -a gpurun_out/bench4.log
