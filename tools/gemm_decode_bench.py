"""Decode-shape GEMM comparison measured the way the engine runs them: all four projections of a Llama-3-8B layer (x 8 distinct
weight sets so nothing stays in L2: 8 x 436 MB >> 126 MB) captured in ONE CUDA graph and replayed; per-op time = graph / count.
  lib    : cuBLAS (torch.mm) + act_and_mul + fused add-norm     (the unfused baseline)
  skinny : our swap-AB split-K tcgen05 kernels with fused consumers (rope/norm fold the slabs, silu fused in the epilogue)"""
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
hbm = float(peaks.get("hbm_gbs", peaks.get("hbm_copy_gbs", 6570)))
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

for M in (1, 16, 64, 128, 256):
    x0 = (torch.randn(M, H, device=dev) * 0.5).bfloat16()
    att = (torch.randn(M, H, device=dev) * 0.5).bfloat16()
    resid = (torch.randn(M, H, device=dev) * 0.5).bfloat16()
    def per_op(kind, key, n_out):
        def f():
            for w in W:
                if kind == "lib": torch.mm(x0 if key != "dn" else hI, w[key].t())
                elif key == "gu": ops.gemm_skinny(x0, w[key], silu_split=True)
                elif key == "dn": ops.gemm_skinny_slabs(hI, w[key])
                else: ops.gemm_skinny_slabs(x0, w[key])
        return f
    hI = (torch.randn(M, I, device=dev) * 0.5).bfloat16()
    row = {}
    for key, nbytes in (("qkv", QKV * H * 2), ("o", H * H * 2), ("gu", 2 * I * H * 2), ("dn", H * I * 2)):
        t_lib = run_graph(per_op("lib", key, 0)) / L * 1000
        t_sk = run_graph(per_op("skinny", key, 0)) / L * 1000
        row[key] = {"lib_us": round(t_lib, 1), "skinny_us": round(t_sk, 1), "hbm_floor_us": round(nbytes / hbm / 1e3, 1),
                    "skinny_frac_of_hbm_roofline": round(nbytes / hbm / 1e3 / t_sk, 3)}
    # whole MLP+attn-projection chain of a layer incl. the elementwise consumers
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
    def layer_skinny():
        r = resid.clone()
        for w in W:
            qkv = ops.gemm_skinny_slabs(x0, w["qkv"])
            o = ops.gemm_skinny_slabs(att, w["o"])
            x1 = ops.slab_reduce_add_rmsnorm(o, r, w["n1"], 1e-5)
            h = ops.gemm_skinny(x1, w["gu"], silu_split=True)
            d = ops.gemm_skinny_slabs(h, w["dn"])
            x2 = ops.slab_reduce_add_rmsnorm(d, r, w["n2"], 1e-5)
    row["layer_chain"] = {"lib_us": round(run_graph(layer_lib) / L * 1000, 1), "skinny_us": round(run_graph(layer_skinny) / L * 1000, 1),
                          "hbm_floor_us": round(436207616 / hbm / 1e3, 1)}
    res[f"M{M}"] = row
    print(f"M={M}", json.dumps(row), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "gemm_decode_bench.json"), "w"), indent=1)
