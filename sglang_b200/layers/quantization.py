"""Quantisation methods for linear layers.

Parity: reference ``python/sglang/srt/layers/quantization/`` (registry ``__init__.py``; ``unquant.py:72-169``;
``fp8.py`` + ``fp8_utils.py:552-808`` block-128x128 fp8 with per-token-group activation quant; W4A16 families
awq/gptq).  Methods here:

* ``UnquantizedMethod``  bf16 weights -> ``ops.linear`` (own tcgen05 GEMM or cuBLASLt)
* ``Fp8BlockMethod``     e4m3 weights with 128x128 block scales, activations quantised per token per 128-group
                         (UE8M0-compatible power-of-two scales optional) -> block-scaled GEMM
* ``Fp8PerTensorMethod`` e4m3 weights with one scale (static/dynamic activation scale)
* ``W4A16Method``        int4 group-quantised weights (group 128), dequantised into the bf16 GEMM
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import nn

from .. import ops
from ..ops import torch_ref


class QuantMethod:
    name = "base"

    def create_weights(self, layer: nn.Module, in_features: int, out_features: int, dtype: torch.dtype) -> None:
        raise NotImplementedError

    def process_weights_after_loading(self, layer: nn.Module) -> None:
        pass

    def apply(self, layer: nn.Module, x: torch.Tensor, bias: Optional[torch.Tensor]) -> torch.Tensor:
        raise NotImplementedError


class UnquantizedMethod(QuantMethod):
    name = "none"

    def create_weights(self, layer, in_features, out_features, dtype):
        layer.weight = nn.Parameter(torch.zeros(out_features, in_features, dtype=dtype), requires_grad=False)

    def apply(self, layer, x, bias):
        return ops.linear(x, layer.weight, bias)


class Fp8BlockMethod(QuantMethod):
    """weights e4m3 [N,K] + fp32 scales [ceil(N/128), ceil(K/128)]; checkpoint may already be block-fp8
    (``weight_scale_inv``) or bf16 (quantised online in ``process_weights_after_loading``)."""
    name = "fp8_block"
    block = 128

    def __init__(self, ue8m0: bool = False):
        self.ue8m0 = ue8m0

    def create_weights(self, layer, in_features, out_features, dtype):
        layer.weight = nn.Parameter(torch.zeros(out_features, in_features, dtype=dtype), requires_grad=False)
        layer.weight_scale_inv = None
        layer.out_dtype = dtype

    def process_weights_after_loading(self, layer):
        if layer.weight.dtype == torch.float8_e4m3fn:
            return
        q, s = torch_ref.block_quant_weight_fp8(layer.weight.data, self.block, self.ue8m0)
        layer.weight = nn.Parameter(q, requires_grad=False)
        layer.weight_scale_inv = nn.Parameter(s, requires_grad=False)

    def apply(self, layer, x, bias):
        from ..ops import quant_ops

        y = quant_ops.fp8_block_linear(x, layer.weight, layer.weight_scale_inv, self.block, self.ue8m0, layer.out_dtype)
        if bias is not None:
            y = y + bias
        return y


class Fp8PerTensorMethod(QuantMethod):
    name = "fp8"

    def create_weights(self, layer, in_features, out_features, dtype):
        layer.weight = nn.Parameter(torch.zeros(out_features, in_features, dtype=dtype), requires_grad=False)
        layer.weight_scale = None
        layer.out_dtype = dtype

    def process_weights_after_loading(self, layer):
        if layer.weight.dtype == torch.float8_e4m3fn:
            return
        w = layer.weight.data.float()
        s = w.abs().amax().clamp(min=1e-10) / torch_ref.FP8_MAX
        layer.weight = nn.Parameter((w / s).clamp(-448, 448).to(torch.float8_e4m3fn), requires_grad=False)
        layer.weight_scale = nn.Parameter(s.reshape(1), requires_grad=False)

    def apply(self, layer, x, bias):
        xs = x.float().abs().amax().clamp(min=1e-10) / torch_ref.FP8_MAX
        xq = (x.float() / xs).clamp(-448, 448).to(torch.float8_e4m3fn)
        if x.is_cuda and hasattr(torch, "_scaled_mm"):
            y = torch._scaled_mm(xq, layer.weight.t(), scale_a=xs.reshape(1).float(), scale_b=layer.weight_scale.float(),
                                 out_dtype=layer.out_dtype)
        else:
            y = ((xq.float() * xs) @ (layer.weight.float() * layer.weight_scale).t()).to(layer.out_dtype)
        return y if bias is None else y + bias


class W4A16Method(QuantMethod):
    """symmetric int4, group size 128 along K; packed two nibbles per byte [N, K/2] + fp16-ish scales [N, K/128]"""
    name = "w4a16"
    group = 128

    def create_weights(self, layer, in_features, out_features, dtype):
        layer.weight = nn.Parameter(torch.zeros(out_features, in_features, dtype=dtype), requires_grad=False)
        layer.qweight = None
        layer.scales = None
        layer.out_dtype = dtype

    def process_weights_after_loading(self, layer):
        w = layer.weight.data.float()
        N, K = w.shape
        g = self.group
        wg = w.view(N, K // g, g)
        s = wg.abs().amax(-1).clamp(min=1e-8) / 7.0
        q = torch.round(wg / s[..., None]).clamp(-8, 7).to(torch.int8).view(N, K)
        packed = ((q[:, 0::2] & 0xF) | ((q[:, 1::2] & 0xF) << 4)).to(torch.uint8)
        layer.qweight = nn.Parameter(packed, requires_grad=False)
        layer.scales = nn.Parameter(s.to(layer.out_dtype), requires_grad=False)
        layer.weight = None

    @staticmethod
    def dequant(qweight: torch.Tensor, scales: torch.Tensor, group: int, dtype) -> torch.Tensor:
        lo = (qweight & 0xF).to(torch.int8)
        hi = (qweight >> 4).to(torch.int8)
        lo = torch.where(lo > 7, lo - 16, lo)
        hi = torch.where(hi > 7, hi - 16, hi)
        q = torch.stack([lo, hi], dim=-1).flatten(-2).float()
        N, K = q.shape
        return (q.view(N, K // group, group) * scales.float()[..., None]).view(N, K).to(dtype)

    def apply(self, layer, x, bias):
        w = self.dequant(layer.qweight, layer.scales, self.group, x.dtype)
        return ops.linear(x, w, bias, impl="lib")


def get_quant_method(name: Optional[str]) -> Optional[QuantMethod]:
    if name in (None, "none", "bf16"):
        return None
    if name in ("fp8_block", "fp8-block", "block_fp8"):
        return Fp8BlockMethod()
    if name in ("mxfp8", "fp8_block_ue8m0"):
        return Fp8BlockMethod(ue8m0=True)
    if name == "fp8":
        return Fp8PerTensorMethod()
    if name in ("w4a16", "int4", "awq", "gptq"):
        return W4A16Method()
    raise ValueError(f"unsupported quantization {name!r}")
