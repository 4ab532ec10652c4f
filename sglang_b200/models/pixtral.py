"""Mistral-Small-3.x multimodal (``Mistral3ForConditionalGeneration``): Pixtral vision transformer + learned 2x2 patch merger + Mistral decoder.

Parity: reference ``python/sglang/srt/models/pixtral.py`` (``PixtralVisionModel``: RMSNorm pre-norm blocks, gated SiLU MLP, 2-D rotary table
indexed by the patch's (row, column)) and ``models/mistral.py`` (``Mistral3ForConditionalGeneration``); processor layout
``multimodal/processors/pixtral.py`` (one ``[IMG]`` per merged patch, ``[IMG_BREAK]`` after every row, ``[IMG_END]`` after the last).

Images keep their aspect ratio, so every image has its own patch grid: the tower consumes the flattened-patch + grid representation that the
Qwen-VL path already uses (``pixel_patches`` [N, 3 p p] in raster order, ``grid_thw`` = (1, rows, cols) per image) and attends inside one image
only.  Only the ``[IMG]`` rows are replaced by vision features; break / end tokens are ordinary vocabulary entries, so one image contributes
one contiguous run per merged row to the embed splice."""
from __future__ import annotations

from typing import Any, Dict, Iterable, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

from ..configs.model_config import ModelConfig
from .llama import _DT, LlamaForCausalLM


class _RMS(nn.Module):
    def __init__(self, n: int, eps: float, dtype):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(n, dtype=dtype), requires_grad=False)
        self.eps = eps

    def forward(self, x):
        y = x.float()
        return (y * torch.rsqrt(y.pow(2).mean(-1, keepdim=True) + self.eps)).to(x.dtype) * self.weight


def _rot_half(x):
    a, b = x[..., : x.shape[-1] // 2], x[..., x.shape[-1] // 2:]
    return torch.cat([-b, a], -1)


class PixtralBlock(nn.Module):
    def __init__(self, h: int, inter: int, heads: int, head_dim: int, dtype, act: str = "gelu"):
        super().__init__()
        self.heads, self.hd = heads, head_dim
        self.act = F.silu if act == "silu" else (lambda t: F.gelu(t, approximate="tanh")) if "tanh" in act else F.gelu
        self.attention_norm, self.ffn_norm = _RMS(h, 1e-5, dtype), _RMS(h, 1e-5, dtype)
        self.q_proj = nn.Linear(h, heads * head_dim, bias=False, dtype=dtype)
        self.k_proj = nn.Linear(h, heads * head_dim, bias=False, dtype=dtype)
        self.v_proj = nn.Linear(h, heads * head_dim, bias=False, dtype=dtype)
        self.o_proj = nn.Linear(heads * head_dim, h, bias=False, dtype=dtype)
        self.gate_proj = nn.Linear(h, inter, bias=False, dtype=dtype)
        self.up_proj = nn.Linear(h, inter, bias=False, dtype=dtype)
        self.down_proj = nn.Linear(inter, h, bias=False, dtype=dtype)

    def forward(self, x, cos, sin, seg: List[int]):
        N = x.shape[0]
        y = self.attention_norm(x)
        q, k, v = (p(y).view(N, self.heads, self.hd) for p in (self.q_proj, self.k_proj, self.v_proj))
        q = (q.float() * cos[:, None] + _rot_half(q.float()) * sin[:, None]).to(x.dtype)
        k = (k.float() * cos[:, None] + _rot_half(k.float()) * sin[:, None]).to(x.dtype)
        outs, s = [], 0
        for n in seg:          # block-diagonal: full attention inside one image
            o = F.scaled_dot_product_attention(q[s:s + n].transpose(0, 1), k[s:s + n].transpose(0, 1), v[s:s + n].transpose(0, 1))
            outs.append(o.transpose(0, 1).reshape(n, -1))
            s += n
        x = x + self.o_proj(torch.cat(outs, 0))
        y = self.ffn_norm(x)
        return x + self.down_proj(self.act(self.gate_proj(y)) * self.up_proj(y))


class PixtralVisionTower(nn.Module):
    def __init__(self, v: Dict[str, Any], dtype):
        super().__init__()
        h, self.patch = v["hidden_size"], v["patch_size"]
        heads = v["num_attention_heads"]
        hd = v.get("head_dim") or h // heads
        self.side = v["image_size"] // self.patch
        self.patch_conv = nn.Linear(3 * self.patch * self.patch, h, bias=False, dtype=dtype)          # Conv2d with stride = kernel
        self.ln_pre = _RMS(h, 1e-5, dtype)
        self.layers = nn.ModuleList([PixtralBlock(h, v["intermediate_size"], heads, hd, dtype, v.get("hidden_act", "gelu")) for _ in range(v["num_hidden_layers"])])
        f = 1.0 / (float(v.get("rope_theta", 10000.0)) ** (torch.arange(0, hd, 2, dtype=torch.float32) / hd))
        self.register_buffer("freq_h", f[::2].clone(), persistent=False)          # even pairs rotate with the row, odd pairs with the column
        self.register_buffer("freq_w", f[1::2].clone(), persistent=False)

    def forward(self, patches: torch.Tensor, grid_thw: np.ndarray) -> torch.Tensor:
        x = self.ln_pre(self.patch_conv(patches))
        rows, cols, seg = [], [], []
        for _, gh, gw in grid_thw.tolist():
            rows.append(torch.arange(gh).repeat_interleave(gw))
            cols.append(torch.arange(gw).repeat(gh))
            seg.append(gh * gw)
        r = torch.cat(rows).to(x.device).float()
        c = torch.cat(cols).to(x.device).float()
        ang = torch.cat([r[:, None] * self.freq_h[None], c[:, None] * self.freq_w[None]], -1)
        ang = torch.cat([ang, ang], -1)
        cos, sin = ang.cos(), ang.sin()
        for blk in self.layers:
            x = blk(x, cos, sin, seg)
        return x


class Mistral3Projector(nn.Module):
    def __init__(self, vh: int, th: int, merge: int, eps: float, bias: bool, act: str, dtype):
        super().__init__()
        self.merge = merge
        self.norm = _RMS(vh, eps, dtype)
        self.merging_layer = nn.Linear(vh * merge * merge, vh, bias=False, dtype=dtype)
        self.linear_1 = nn.Linear(vh, th, bias=bias, dtype=dtype)
        self.linear_2 = nn.Linear(th, th, bias=bias, dtype=dtype)
        self.act = act

    def forward(self, x: torch.Tensor, grid_thw: np.ndarray) -> torch.Tensor:
        x = self.norm(x)
        d, m = x.shape[-1], self.merge
        outs, s = [], 0
        for _, gh, gw in grid_thw.tolist():
            g = x[s:s + gh * gw].view(gh, gw, d).permute(2, 0, 1).unsqueeze(0)
            outs.append(F.unfold(g.float(), kernel_size=m, stride=m).view(d * m * m, -1).t().to(x.dtype))       # [tokens, d*m*m], channel-major
            s += gh * gw
        y = self.linear_1(self.merging_layer(torch.cat(outs, 0)))
        y = F.gelu(y) if self.act == "gelu" else F.silu(y)
        return self.linear_2(y)


class Mistral3ForConditionalGeneration(LlamaForCausalLM):
    def __init__(self, cfg: ModelConfig, quant_name: Optional[str] = None, device="cpu"):
        super().__init__(cfg, quant_name, device)
        v = cfg.vision
        dt = _DT[cfg.dtype]
        if v.get("skip_tower"):
            self.vision_tower = self.projector = None
        else:
            self.vision_tower = PixtralVisionTower(v, dt)
            self.projector = Mistral3Projector(v["hidden_size"], cfg.hidden_size, v["spatial_merge_size"], cfg.rms_norm_eps,
                                               bool(v.get("projector_bias", False)), v.get("projector_hidden_act", "gelu"), dt)

    @torch.no_grad()
    def encode_images(self, patches: torch.Tensor, grid_thw: Any = None) -> torch.Tensor:
        g = np.asarray(grid_thw, dtype=np.int64).reshape(-1, 3)
        return self.projector(self.vision_tower(patches, g), g)

    def load_weights(self, weights: Iterable[Tuple[str, torch.Tensor]]) -> List[str]:
        own = dict(self.named_parameters())
        rest, loaded = [], []
        for name, w in weights:
            n = name[len("model."):] if name.startswith(("model.vision_tower.", "model.multi_modal_projector.", "model.language_model.")) else name
            if n.startswith("language_model."):
                n = n[len("language_model."):]
                rest.append((n if n.startswith(("model.", "lm_head.")) else "model." + n, w))
                continue
            if n.startswith(("vision_tower.", "multi_modal_projector.")):
                if self.vision_tower is None:
                    continue
                n = n.replace("vision_tower.transformer.layers.", "vision_tower.layers.").replace(".attention.", ".").replace(".feed_forward.", ".") \
                    .replace("multi_modal_projector.patch_merger.", "projector.").replace("multi_modal_projector.", "projector.")
                if n in own:
                    own[n].data.copy_(w.reshape(own[n].shape))
                    loaded.append(name)
                continue
            rest.append((name, w))
        return loaded + super().load_weights(rest)


class _LlavaProjector(nn.Module):
    def __init__(self, vh: int, th: int, bias: bool, act: str, dtype):
        super().__init__()
        self.linear_1 = nn.Linear(vh, th, bias=bias, dtype=dtype)
        self.linear_2 = nn.Linear(th, th, bias=bias, dtype=dtype)
        self.act = act

    def forward(self, x: torch.Tensor, grid_thw=None) -> torch.Tensor:
        y = self.linear_1(x)
        return self.linear_2(F.gelu(y) if self.act == "gelu" else F.silu(y))


class PixtralForConditionalGeneration(Mistral3ForConditionalGeneration):
    """Pixtral-12B in the Hugging Face LLaVA layout (``LlavaForConditionalGeneration`` whose ``vision_config.model_type`` is ``pixtral``; reference
    ``python/sglang/srt/models/pixtral.py``): the same tower, no patch merging — every patch is one image token — and the two-layer LLaVA
    projector.  The config loader routes such checkpoints here."""

    def __init__(self, cfg: ModelConfig, quant_name: Optional[str] = None, device="cpu"):
        super().__init__(cfg, quant_name, device)
        if self.vision_tower is not None:
            v = cfg.vision
            self.projector = _LlavaProjector(v["hidden_size"], cfg.hidden_size, bool(v.get("projector_bias", True)), v.get("projector_hidden_act", "gelu"),
                                             _DT[cfg.dtype])


class LightOnOCRForConditionalGeneration(Mistral3ForConditionalGeneration):
    """LightOnOCR (reference ``python/sglang/srt/models/lightonocr.py``): the Mistral3 recipe — Pixtral tower, RMSNorm + learned 2x2 merger, two
    bias-free linears — in front of a Qwen3 decoder (per-head q/k norm comes from the config); the checkpoint calls the two vision parts
    ``vision_encoder`` / ``vision_projection``."""

    def load_weights(self, weights):
        def rename():
            for name, w in weights:
                yield name.replace("vision_encoder.", "vision_tower.").replace("vision_projection.", "multi_modal_projector."), w
        return super().load_weights(rename())
