#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -25 | tee gpurun_out/pytest_gpu9.log
for persist in 0 1; do
SGLB_ATTN_PERSIST=$persist python - <<'PY' 2>&1 | grep -v "^W0" | tee gpurun_out/attn_micro9_p$persist.log
import sys, os, json, torch
sys.path.insert(0, "tests")
from sglang_b200 import ops
from sglang_b200.environ import envs
from gpu_cases import _attn_setup
ops.load_native()
peaks=json.load(open("MEASURED_PEAKS.json")) if os.path.exists("MEASURED_PEAKS.json") else {}
hbm=float(peaks.get("hbm_gbs", 6570)); tf=float(peaks.get("bf16_tflops", 1689))
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
("decode_bs64_ctx2048",64,[1]*64,[2048]*64,True,1),("decode_bs512_ctx512",512,[1]*512,[512]*512,True,1),("decode_bs1024_ragged",1024,[1]*1024,rng.integers(16,900,1024).tolist(),True,1),
("prefill_28seq",28,rng.integers(50,600,28).tolist(),None,False,1),("prefill_64seq_ragged",64,rng.integers(20,1200,64).tolist(),None,False,1),("prefill_4x2048",4,[2048]*4,[2048]*4,False,1),("prefill_1x8192",1,[8192],[8192],False,1)]
res={}
for tag,B,ql,sl,dec,sp in cases:
    sl = sl or ql
    s=_attn_setup(B,ql,sl,32,8,128,32); out=torch.empty_like(s["q"])
    ms=t(lambda: ops.paged_attention(s["q"],s["kc"],s["vc"],s["pt"],s["slots"],s["seq"],None if dec else s["cu"],None if dec else s["ql"],max(ql),32,128**-0.5,num_splits=sp,impl=0,out=out))
    kvb=sum(sl)*8*128*4; fl=sum(4.0*32*128*q*(l-q/2) for q,l in zip(ql,sl))
    roof=max(kvb/(hbm*1e6), fl/(tf*1e9))
    res[tag]={"ms":round(ms,4),"kv_gbs":round(kvb/ms/1e6),"tflops":round(fl/ms/1e9,1),"roofline_ms":round(roof,4),"frac_of_measured_roofline":round(roof/ms,3)}
    print(("persist" if envs.SGLB_ATTN_PERSIST.get() else "classic"), tag, res[tag], flush=True)
json.dump(res, open(f"gpurun_out/attn_micro9_p{int(envs.SGLB_ATTN_PERSIST.get())}.json","w"), indent=1)
PY
done
timeout 300 python tools/gemm_decode_bench.py 2>&1 | grep -v "^W0" | tee gpurun_out/gemm_decode_bench9.log
echo "== bench 1024 prompts, max-running 1024 (persistent attention, skinny=tp)" | tee gpurun_out/bench9.log
timeout 900 python bench.py --steps 2 --warmup 3 --num-prompts 1024 --max-running 1024 2>&1 | tail -1 | tee -a gpurun_out/bench9.log
