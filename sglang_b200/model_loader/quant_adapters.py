"""Quantised *checkpoint formats* -> the engine's own weight formats.

Parity: reference ``python/sglang/srt/layers/quantization/`` (28 methods: awq, gptq, compressed-tensors w4a16 / w8a8-int8 / fp8, fp8
per-tensor / per-channel / block, modelopt fp8 / NVFP4, mxfp4, bitsandbytes nf4 / fp4 …).  The reference keeps one kernel family per format
(marlin, machete, cutlass w4a8, triton int8, …).  This engine has exactly three runtime weight formats, all with sm_100a kernels —
bf16 (tcgen05 GEMM), fp8 e4m3 with 128x128 block scales (``kernels/gemm_fp8.cu``) and int4 group-128 (W4A16) — so a foreign format is an
*input* problem: ``adapt_quantized_checkpoint`` wraps the checkpoint iterator, collects the tensors that belong to one quantised linear
(``qweight / qzeros / scales / g_idx``, ``weight_packed / weight_scale / weight_zero_point``, ``weight + weight_scale[_inv|_2]``,
``*_blocks / *_scales``, bitsandbytes ``absmax / quant_map / quant_state``), de-quantises them exactly and yields ``<prefix>.weight``;
sharding, stacking (qkv / gate_up) and optional re-quantisation (``--quantization fp8_block|w4a16``) then run unchanged.

Block-fp8 checkpoints (``weight_scale_inv``) pass through untouched when the engine itself runs ``fp8_block`` (no precision round trip).
"""
from __future__ import annotations

import json
from typing import Dict, Iterable, Iterator, List, Optional, Tuple

import torch

_AWQ_ORDER = [0, 4, 1, 5, 2, 6, 3, 7]          # nibble i of an int32 holds column 8 j + _AWQ_ORDER[i]
# e2m1 values (sign bit = 8)
_FP4 = torch.tensor([0.0, 0.5, 1.0, 1.5, 2.0, 3.0, 4.0, 6.0, -0.0, -0.5, -1.0, -1.5, -2.0, -3.0, -4.0, -6.0])
_SUFFIXES = ("qweight", "qzeros", "scales", "g_idx", "weight_packed", "weight_scale", "weight_zero_point", "weight_shape", "weight_g_idx",
             "weight_scale_inv", "weight_scale_2", "input_scale", "input_scale_inv", "k_scale", "v_scale", "weight.absmax", "weight.quant_map",
             "weight.nested_absmax", "weight.nested_quant_map", "weight.quant_state.bitsandbytes__nf4", "weight.quant_state.bitsandbytes__fp4",
             "weight")


def _unpack_int32(x: torch.Tensor, bits: int, dim: int, order: Optional[List[int]] = None) -> torch.Tensor:
    """unpack ``32 / bits`` fields per int32 along ``dim`` (field i = bits [i*bits, (i+1)*bits))"""
    n = 32 // bits
    shifts = torch.arange(n, dtype=torch.int32) * bits
    x = x.to(torch.int32)
    parts = ((x.unsqueeze(-1) >> shifts) & ((1 << bits) - 1)).to(torch.int32)         # [..., n]
    if order is not None:
        inv = [order.index(i) for i in range(n)]
        parts = parts[..., inv]
    if dim in (-1, x.dim() - 1):
        return parts.flatten(-2)
    assert dim == 0 and x.dim() == 2
    return parts.permute(0, 2, 1).reshape(x.shape[0] * n, x.shape[1])


# ------------------------------------------------------------------------------------------------ de-quantisers (exact)
def dequant_awq(qweight, qzeros, scales, bits: int = 4) -> torch.Tensor:
    """AutoAWQ GEMM layout: qweight i32 [K, N/8], qzeros i32 [K/g, N/8], scales [K/g, N]  ->  [N, K]"""
    q = _unpack_int32(qweight, bits, -1, _AWQ_ORDER)                      # [K, N]
    z = _unpack_int32(qzeros, bits, -1, _AWQ_ORDER)                       # [K/g, N]
    g = q.shape[0] // scales.shape[0]
    w = (q - z.repeat_interleave(g, 0)).float() * scales.float().repeat_interleave(g, 0)
    return w.t().contiguous()


def dequant_gptq(qweight, qzeros, scales, g_idx=None, bits: int = 4, zero_offset: int = 1) -> torch.Tensor:
    """AutoGPTQ: qweight i32 [K*bits/32, N] packed along K, qzeros i32 [K/g, N*bits/32], scales [K/g, N], g_idx [K] -> [N, K];
    v1 checkpoints store ``zero - 1`` (``zero_offset=1``), gptq_v2 stores the zero itself"""
    q = _unpack_int32(qweight, bits, 0)                                   # [K, N]
    z = _unpack_int32(qzeros, bits, -1) + zero_offset                     # [K/g, N]
    K = q.shape[0]
    if g_idx is None or g_idx.numel() == 0:
        g_idx = torch.arange(K) // (K // scales.shape[0])
    gi = g_idx.long()
    w = (q - z[gi]).float() * scales.float()[gi]
    return w.t().contiguous()


def dequant_compressed_int(weight_packed, weight_scale, zero_point=None, shape=None, bits: int = 4, g_idx=None) -> torch.Tensor:
    """compressed-tensors pack-quantized: weight_packed i32 [N, K*bits/32] (unsigned = signed + 2^(bits-1)), weight_scale [N, K/g] or [N, 1],
    optional packed zero point [N*bits/32, K/g]"""
    q = _unpack_int32(weight_packed, bits, -1)
    if shape is not None:
        q = q[: int(shape[0]), : int(shape[1])]
    N, K = q.shape
    s = weight_scale.float()
    ng = s.shape[1] if s.dim() == 2 else 1
    s = s.reshape(N, ng)
    gi = (g_idx.long() if g_idx is not None and g_idx.numel() and int(g_idx.min()) >= 0 else torch.arange(K) // (K // ng))
    if zero_point is not None:
        z = _unpack_int32(zero_point, bits, 0)[:N] if zero_point.dtype == torch.int32 else zero_point.to(torch.int32)
        w = (q - z[:, gi]).float()
    else:
        w = (q - (1 << (bits - 1))).float()
    return w * s[:, gi]


def dequant_scaled(weight, scale) -> torch.Tensor:
    """int8 / fp8 weight [N, K] with a scalar, per-channel [N] / [N, 1] or per-128x128-block [N/128, K/128] scale"""
    w = weight.float()
    s = scale.float()
    if s.numel() == 1:
        return w * s.reshape(())
    if s.dim() == 2 and s.shape[0] != w.shape[0]:
        bn, bk = -(-w.shape[0] // s.shape[0]), -(-w.shape[1] // s.shape[1])
        return w * s.repeat_interleave(bn, 0)[: w.shape[0]].repeat_interleave(bk, 1)[:, : w.shape[1]]
    if s.dim() == 2 and s.shape[1] not in (1, w.shape[1]):
        return w * s.repeat_interleave(w.shape[1] // s.shape[1], 1)          # per-group along K
    return w * s.reshape(w.shape[0], -1)


def _fp4_unpack(packed: torch.Tensor) -> torch.Tensor:
    """uint8 [..., K/2] (low nibble first) -> e2m1 values [..., K]"""
    p = packed.view(torch.uint8)
    idx = torch.stack([p & 0xF, p >> 4], dim=-1).flatten(-2).long()
    return _FP4[idx]


def dequant_nvfp4(weight, weight_scale, weight_scale_2) -> torch.Tensor:
    """ModelOpt NVFP4: weight u8 [N, K/2] e2m1 pairs, weight_scale fp8-e4m3 [N, K/16], weight_scale_2 fp32 scalar"""
    w = _fp4_unpack(weight)
    s = weight_scale.float() * weight_scale_2.float().reshape(())
    return w * s.repeat_interleave(16, dim=-1)[..., : w.shape[-1]]


def dequant_mxfp4(blocks, scales) -> torch.Tensor:
    """OCP MXFP4 (gpt-oss): blocks u8 [..., K/32, 16], scales u8 e8m0 [..., K/32]  ->  [..., K]"""
    w = _fp4_unpack(blocks)                                                  # [..., K/32, 32]
    e = scales.view(torch.uint8).to(torch.int32) - 127
    return (w * torch.exp2(e.float()).unsqueeze(-1)).flatten(-2)


def dequant_bnb_4bit(weight, absmax, quant_map, state: Dict, nested_absmax=None, nested_quant_map=None) -> torch.Tensor:
    """bitsandbytes nf4 / fp4: weight u8 [N*K/2, 1] (HIGH nibble first), absmax per ``blocksize`` elements (optionally double-quantised)"""
    if nested_absmax is not None:
        nb = int(state.get("nested_blocksize", 256))
        am = nested_quant_map.float()[absmax.view(torch.uint8).long()] * nested_absmax.float().repeat_interleave(nb)[: absmax.numel()]
        am = am + float(state.get("nested_offset", 0.0))
    else:
        am = absmax.float()
    p = weight.view(torch.uint8).reshape(-1)
    idx = torch.stack([p >> 4, p & 0xF], dim=-1).flatten().long()
    bs = int(state.get("blocksize", 64))
    w = quant_map.float()[idx] * am.repeat_interleave(bs)[: idx.numel()]
    shape = state.get("shape")
    return w.reshape(tuple(shape)) if shape else w


# ------------------------------------------------------------------------------------------------ the adapter
def read_quant_config(model_path: str) -> Dict:
    import os

    for name in ("config.json", "hf_quant_config.json", "quantize_config.json"):
        f = os.path.join(model_path, name)
        if os.path.exists(f):
            with open(f) as fh:
                j = json.load(fh)
            q = j.get("quantization_config") or j.get("quantization") or (j if name != "config.json" else None)
            if q:
                return q
    return {}


def _split(name: str) -> Tuple[str, str]:
    for suf in _SUFFIXES:
        if name.endswith("." + suf):
            return name[: -len(suf) - 1], suf
    return name, ""


def _bits_of(qcfg: Dict, default: int = 4) -> int:
    if "bits" in qcfg:
        return int(qcfg["bits"])
    if "w_bit" in qcfg:
        return int(qcfg["w_bit"])
    for grp in (qcfg.get("config_groups") or {}).values():
        w = grp.get("weights") or {}
        if "num_bits" in w:
            return int(w["num_bits"])
    return default


def dequant_module(parts: Dict[str, torch.Tensor], qcfg: Dict) -> Optional[torch.Tensor]:
    """one quantised linear's tensors (by suffix) -> fp32 weight [N, K]; None when the module is not quantised"""
    method = (qcfg.get("quant_method") or qcfg.get("quant_algo") or "").lower()
    if "qweight" in parts:
        bits = _bits_of(qcfg)
        if method == "awq" or ("g_idx" not in parts and parts["qweight"].shape[0] == parts["scales"].shape[0] * max(1, int(qcfg.get("group_size", 128)))
                               and method != "gptq"):
            return dequant_awq(parts["qweight"], parts["qzeros"], parts["scales"], bits)
        v2 = qcfg.get("checkpoint_format") == "gptq_v2"
        return dequant_gptq(parts["qweight"], parts["qzeros"], parts["scales"], parts.get("g_idx"), bits, 0 if v2 else 1)
    if "weight_packed" in parts:
        return dequant_compressed_int(parts["weight_packed"], parts["weight_scale"], parts.get("weight_zero_point"), parts.get("weight_shape"),
                                      _bits_of(qcfg), parts.get("weight_g_idx"))
    w = parts.get("weight")
    if w is None:
        return None
    if any(k.startswith("weight.quant_state") for k in parts):
        key = next(k for k in parts if k.startswith("weight.quant_state"))
        state = json.loads(bytes(parts[key].view(torch.uint8).tolist()).decode("utf-8"))
        return dequant_bnb_4bit(w, parts["weight.absmax"], parts["weight.quant_map"], state, parts.get("weight.nested_absmax"),
                                parts.get("weight.nested_quant_map"))
    if "weight_scale_2" in parts:
        return dequant_nvfp4(w, parts["weight_scale"], parts["weight_scale_2"])
    if "weight_scale_inv" in parts:
        return dequant_scaled(w, parts["weight_scale_inv"])
    if "weight_scale" in parts and w.dtype in (torch.int8, torch.float8_e4m3fn, torch.float8_e5m2, torch.uint8):
        return dequant_scaled(w, parts["weight_scale"])
    return None


def adapt_quantized_checkpoint(it: Iterable[Tuple[str, torch.Tensor]], qcfg: Optional[Dict] = None, dtype: torch.dtype = torch.bfloat16,
                               keep_block_fp8: bool = False) -> Iterator[Tuple[str, torch.Tensor]]:
    """wrap a checkpoint iterator: quantised linears come out as ``<prefix>.weight`` in ``dtype``; everything else passes through.
    Tensors of one module may arrive in any order and across shards, so quantised modules are emitted when the stream ends or as soon as
    their set is complete (``qweight+qzeros+scales``; the rest is flushed at the end)."""
    qcfg = qcfg or {}
    pending: Dict[str, Dict[str, torch.Tensor]] = {}
    quantised_marks = {"qweight", "weight_packed", "weight_scale", "weight_scale_inv", "weight_scale_2", "weight.absmax"}

    def flush(prefix: str) -> Iterator[Tuple[str, torch.Tensor]]:
        parts = pending.pop(prefix)
        if keep_block_fp8 and "weight_scale_inv" in parts and parts.get("weight") is not None and parts["weight"].dtype == torch.float8_e4m3fn:
            yield prefix + ".weight", parts["weight"]
            yield prefix + ".weight_scale_inv", parts["weight_scale_inv"]
            return
        w = dequant_module(parts, qcfg)
        if w is None:
            for suf, t in parts.items():
                yield f"{prefix}.{suf}", t
            return
        yield prefix + ".weight", w.to(dtype)
        for suf in ("k_scale", "v_scale"):
            if suf in parts:
                yield f"{prefix}.{suf}", parts[suf]

    for name, t in it:
        if name.endswith("_blocks") and t.dtype == torch.uint8:            # mxfp4 expert tensors: <x>_blocks + <x>_scales
            pending.setdefault(name[:-7], {})["blocks"] = t
        elif name.endswith("_scales") and t.dtype == torch.uint8 and name[:-7] in pending:
            pending[name[:-7]]["scales"] = t
        else:
            prefix, suf = _split(name)
            if suf in ("", "bias") or (suf == "weight" and t.dtype in (torch.bfloat16, torch.float16, torch.float32)):
                yield name, t
                continue
            pending.setdefault(prefix, {})[suf] = t
        # early emission keeps peak host memory at one module
        for pre in list(pending):
            p = pending[pre]
            if "blocks" in p and "scales" in p:
                pending.pop(pre)
                w = dequant_mxfp4(p["blocks"], p["scales"]).to(dtype)
                # stacked expert tensors are stored [E, out, in] in MXFP4 checkpoints but [E, in, out] in the bf16 ones: emit the latter
                yield pre, (w.transpose(1, 2).contiguous() if w.dim() == 3 else w)
            elif {"qweight", "qzeros", "scales"} <= set(p) and ("g_idx" in p or (qcfg.get("quant_method") or "").lower() == "awq"):
                yield from flush(pre)
    for pre in list(pending):
        p = pending[pre]
        if "blocks" in p:        # scales never came: malformed
            raise ValueError(f"mxfp4 tensor {pre}_blocks has no matching _scales")
        if set(p) == {"weight"} or not (set(p) & quantised_marks):
            for suf, t in pending.pop(pre).items():
                yield f"{pre}.{suf}", t
        else:
            yield from flush(pre)
