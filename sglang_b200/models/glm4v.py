"""GLM-4.1V (``Glm4vForConditionalGeneration``): dynamic-resolution vision transformer (RMSNorm blocks, gated SiLU MLP, 2-D rotary, a learned
position grid sampled bicubically at every patch centre, 2x2 convolutional down-sampling, gated merger) feeding the GLM-4 decoder (sandwich
norms, partial interleaved rotary) with multimodal rotary positions.

Parity: reference ``python/sglang/srt/models/glm4v.py`` + ``multimodal/processors/glm4v.py``.  Everything around the tower is shared with the
Qwen-VL path of this package: flattened patches + grids in merge-block order, hash pad tokens, the per-token cos/sin table for M-RoPE
(``models/qwen2_vl.py``), and the GLM-4 text stack (``models/families2.py``)."""
from __future__ import annotations

from typing import Any, Dict, List

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

from .families2 import Glm4ForCausalLM, Glm4MoeForCausalLM
from .qwen2_vl import _VRMS, Qwen2VLForConditionalGeneration, Qwen2VisionTransformer, VisionBlock25


class _GlmBlock(VisionBlock25):
    def __init__(self, dim: int, heads: int, inter: int, dtype, eps: float, attn_bias: bool):
        nn.Module.__init__(self)
        self.heads = heads
        self.norm1, self.norm2 = _VRMS(dim, dtype, eps), _VRMS(dim, dtype, eps)
        self.qkv, self.proj = nn.Linear(dim, 3 * dim, bias=attn_bias, dtype=dtype), nn.Linear(dim, dim, bias=False, dtype=dtype)
        self.gate_proj, self.up_proj = nn.Linear(dim, inter, bias=False, dtype=dtype), nn.Linear(dim, inter, bias=False, dtype=dtype)
        self.down_proj = nn.Linear(inter, dim, bias=False, dtype=dtype)


class _GlmMerger(nn.Module):
    def __init__(self, dim: int, ctx: int, dtype):
        super().__init__()
        self.proj = nn.Linear(dim, dim, bias=False, dtype=dtype)
        self.post_projection_norm = nn.LayerNorm(dim, dtype=dtype)
        self.gate_proj, self.up_proj = nn.Linear(dim, ctx, bias=False, dtype=dtype), nn.Linear(dim, ctx, bias=False, dtype=dtype)
        self.down_proj = nn.Linear(ctx, dim, bias=False, dtype=dtype)

    def forward(self, x):
        x = F.gelu(self.post_projection_norm(self.proj(x)))
        return self.down_proj(F.silu(self.gate_proj(x)) * self.up_proj(x))


class Glm4vVisionTransformer(Qwen2VisionTransformer):
    def __init__(self, v: Dict[str, Any], dtype):
        nn.Module.__init__(self)
        self.embed_dim, self.heads, self.patch, self.merge, self.tps = v["embed_dim"], v["num_heads"], v["patch_size"], v["spatial_merge_size"], \
            v.get("temporal_patch_size", 2)
        eps = v.get("rms_norm_eps", 1e-5)
        self.patch_embed = nn.Linear(3 * self.tps * self.patch * self.patch, self.embed_dim, bias=True, dtype=dtype)
        self.post_conv_layernorm = _VRMS(self.embed_dim, dtype, eps)
        self.side = v["image_size"] // self.patch
        self.position_embedding = nn.Embedding(self.side ** 2, self.embed_dim, dtype=dtype)
        self.blocks = nn.ModuleList([_GlmBlock(self.embed_dim, self.heads, v["out_hidden_size"], dtype, eps, bool(v.get("attention_bias", False)))
                                     for _ in range(v["depth"])])
        self.post_layernorm = _VRMS(self.embed_dim, dtype, eps)
        self.downsample = nn.Conv2d(self.embed_dim, v["out_hidden_size"], self.merge, self.merge, dtype=dtype)
        self.merger = _GlmMerger(v["out_hidden_size"], v["intermediate_size"], dtype)
        hd = self.embed_dim // self.heads
        self.register_buffer("inv_freq", 1.0 / (10000.0 ** (torch.arange(0, hd // 2, 2, dtype=torch.float32) / (hd // 2))), persistent=False)

    def _coords(self, grid_thw: np.ndarray):
        hs, ws, th, tw = [], [], [], []
        m = self.merge
        for t, h, w in grid_thw.tolist():
            hp = np.arange(h)[:, None].repeat(w, 1).reshape(h // m, m, w // m, m).transpose(0, 2, 1, 3).reshape(-1)
            wp = np.arange(w)[None, :].repeat(h, 0).reshape(h // m, m, w // m, m).transpose(0, 2, 1, 3).reshape(-1)
            hs.append(np.tile(hp, t)); ws.append(np.tile(wp, t))
            th.append(np.full(t * h * w, h)); tw.append(np.full(t * h * w, w))
        return (np.concatenate(x).astype(np.float32) for x in (hs, ws, th, tw))

    def pos_sample(self, grid_thw: np.ndarray) -> torch.Tensor:
        """the learned [side, side] grid sampled (bicubic, border padding) at every patch centre of every image"""
        W = self.position_embedding.weight
        hc, wc, th, tw = (torch.from_numpy(x).to(W.device) for x in self._coords(grid_thw))
        grid = torch.stack(((wc + 0.5) / tw * 2 - 1, (hc + 0.5) / th * 2 - 1), -1)[None, :, None, :]
        pe = W.view(self.side, self.side, -1).permute(2, 0, 1)[None].float()
        return F.grid_sample(pe, grid, mode="bicubic", align_corners=False, padding_mode="border")[0, :, :, 0].t().to(W.dtype)

    def forward(self, patches: torch.Tensor, grid_thw: np.ndarray) -> torch.Tensor:
        x = self.post_conv_layernorm(self.patch_embed(patches))
        cos, sin = self.rot_pos_emb(grid_thw)
        x = x + self.pos_sample(grid_thw)
        seg: List[int] = []
        for t, h, w in grid_thw.tolist():
            seg += [h * w] * t
        for blk in self.blocks:
            x = blk(x, cos, sin, seg)
        x = self.post_layernorm(x)
        m = self.merge
        x = self.downsample(x.view(-1, m, m, x.shape[-1]).permute(0, 3, 1, 2)).view(-1, self.downsample.out_channels)
        return self.merger(x)


class Glm4vForConditionalGeneration(Qwen2VLForConditionalGeneration, Glm4ForCausalLM):
    def load_weights(self, weights):
        def rename():
            for name, w in weights:
                yield name.replace("visual.embeddings.position_embedding.", "visual.position_embedding."), w
        return super().load_weights(rename())


class Glm4vMoeForConditionalGeneration(Qwen2VLForConditionalGeneration, Glm4MoeForCausalLM):
    """GLM-4.5V: the same tower in front of the GLM-4.5 MoE decoder (reference ``python/sglang/srt/models/glm4v_moe.py``)"""

    def load_weights(self, weights):
        def rename():
            for name, w in weights:
                yield name.replace("visual.embeddings.position_embedding.", "visual.position_embedding."), w
        return super().load_weights(rename())
