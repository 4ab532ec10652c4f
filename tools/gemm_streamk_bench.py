"""Decode-shape GEMMs measured the way the engine runs them: the four projections of a Llama-3-8B layer x 8 distinct weight sets
(8 x 436 MB >> 126 MB L2) captured in ONE CUDA graph and replayed; per-op time = graph time / count.
  lib     : cuBLAS (torch.mm) [+ act_and_mul + fused add-norm in the layer chain]
  streamk : csrc/kernels/gemm_streamk.cu (persistent stream-K, SiLU*mul fused into gate_up) [+ fused add-norm]
-> gpurun_out/gemm_streamk_bench.json (+ a markdown table on stdout)"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from sglang_b200 import ops
ops.load_native()
dev = "cuda"
H, I, QKV = 4096, 14336, 6144
L = 8
torch.manual_seed(0)
W = [dict(qkv=(torch.randn(QKV, H, device=dev) * 0.02).bfloat16(), o=(torch.randn(H, H, device=dev) * 0.02).bfloat16(),
          gu=(torch.randn(2 * I, H, device=dev) * 0.02).bfloat16(), dn=(torch.randn(H, I, device=dev) * 0.02).bfloat16(),
          n1=torch.ones(H, device=dev).bfloat16(), n2=torch.ones(H, device=dev).bfloat16()) for _ in range(L)]
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
hbm = float(peaks.get("hbm_gbs", 6570))
res = {}

def run_graph(fn, iters=10):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn(); fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        fn()
    for _ in range(3): g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): g.replay()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters

Ms = [int(x) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else "1,16,32,64,128,256".split(","))]
for M in Ms:
    x0 = (torch.randn(M, H, device=dev) * 0.5).bfloat16()
    att = (torch.randn(M, H, device=dev) * 0.5).bfloat16()
    resid = (torch.randn(M, H, device=dev) * 0.5).bfloat16()
    hI = (torch.randn(M, I, device=dev) * 0.5).bfloat16()
    def per_op(kind, key):
        def f():
            for w in W:
                xin = hI if key == "dn" else x0
                if kind == "lib": torch.mm(xin, w[key].t())
                else: ops.gemm_streamk(xin, w[key], silu_split=(key == "gu"))
        return f
    row = {}
    for key, nbytes in (("qkv", QKV * H * 2), ("o", H * H * 2), ("gu", 2 * I * H * 2), ("dn", H * I * 2)):
        t_lib = run_graph(per_op("lib", key)) / L * 1000
        t_sk = run_graph(per_op("streamk", key)) / L * 1000
        fl = nbytes / hbm / 1e3
        row[key] = {"lib_us": round(t_lib, 1), "streamk_us": round(t_sk, 1), "hbm_floor_us": round(fl, 1),
                    "lib_frac": round(fl / t_lib, 3), "streamk_frac": round(fl / t_sk, 3)}
    def layer_lib():
        r = resid.clone()
        for w in W:
            qkv = torch.mm(x0, w["qkv"].t())
            o = torch.mm(att, w["o"].t())
            ops.fused_add_rmsnorm(o, r, w["n1"], 1e-5)
            gu = torch.mm(o, w["gu"].t())
            h = ops.silu_and_mul(gu)
            d = torch.mm(h, w["dn"].t())
            ops.fused_add_rmsnorm(d, r, w["n2"], 1e-5)
    def layer_sk():
        r = resid.clone()
        for w in W:
            qkv = ops.gemm_streamk(x0, w["qkv"])
            o = ops.gemm_streamk(att, w["o"])
            ops.fused_add_rmsnorm(o, r, w["n1"], 1e-5)
            h = ops.gemm_streamk(o, w["gu"], silu_split=True)
            d = ops.gemm_streamk(h, w["dn"])
            ops.fused_add_rmsnorm(d, r, w["n2"], 1e-5)
    fl = 436207616 / hbm / 1e3
    tl, ts = run_graph(layer_lib) / L * 1000, run_graph(layer_sk) / L * 1000
    row["layer_chain"] = {"lib_us": round(tl, 1), "streamk_us": round(ts, 1), "hbm_floor_us": round(fl, 1), "lib_frac": round(fl / tl, 3),
                          "streamk_frac": round(fl / ts, 3)}
    res[f"M{M}"] = row
    print(f"M={M}", json.dumps(row), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump({"hbm_gbs_measured": hbm, "results": res}, open(os.path.join(ROOT, "gpurun_out", "gemm_streamk_bench.json"), "w"), indent=1)
print("| M | op | cuBLAS us | stream-K us | HBM floor us | cuBLAS frac of measured HBM roofline | stream-K frac |\n|---|---|---|---|---|---|---|")
for m, row in res.items():
    for k, v in row.items():
        print(f"| {m[1:]} | {k} | {v['lib_us']} | {v['streamk_us']} | {v['hbm_floor_us']} | {v['lib_frac']} | {v['streamk_frac']} |")
