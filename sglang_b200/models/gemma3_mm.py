"""Gemma-3 multimodal (``Gemma3ForConditionalGeneration``): SigLIP vision tower, average-pool + RMSNorm + linear projector (256 soft tokens
per image), Gemma-3 decoder in which the tokens of one image attend to each other *bidirectionally*.

Parity: reference ``python/sglang/srt/models/gemma3_mm.py`` + ``models/siglip.py``; processor ``multimodal/processors/gemma3.py``.

The bidirectional image blocks are expressed as attention masks for the prefill steps that contain image tokens: ``causal OR same-image``
for the global layers, ``(causal AND window) OR same-image`` for the sliding layers (``ForwardBatch.custom_mask`` / ``custom_mask_sliding``).
Text-only steps and every decode step keep the ordinary kernels.  An image whose tokens are split over two prefill chunks can only see the
part inside its own chunk bidirectionally (the earlier chunk has already been computed), so the scheduler-side chunk size should exceed
``mm_tokens_per_image`` — it does by default (8192 vs 256)."""
from __future__ import annotations

from typing import Any, Dict, Iterable, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

from ..configs.model_config import ModelConfig
from ..model_executor.forward_batch import ForwardBatch
from .families2 import Gemma3ForCausalLM
from .llama import _DT


class SiglipLayer(nn.Module):
    def __init__(self, h: int, inter: int, heads: int, eps: float, dtype):
        super().__init__()
        self.heads = heads
        self.layer_norm1, self.layer_norm2 = nn.LayerNorm(h, eps=eps, dtype=dtype), nn.LayerNorm(h, eps=eps, dtype=dtype)
        self.q_proj, self.k_proj, self.v_proj, self.out_proj = (nn.Linear(h, h, dtype=dtype) for _ in range(4))
        self.fc1, self.fc2 = nn.Linear(h, inter, dtype=dtype), nn.Linear(inter, h, dtype=dtype)

    def forward(self, x):
        B, N, H = x.shape
        y = self.layer_norm1(x)
        q, k, v = (p(y).view(B, N, self.heads, -1).transpose(1, 2) for p in (self.q_proj, self.k_proj, self.v_proj))
        x = x + self.out_proj(F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B, N, H))
        return x + self.fc2(F.gelu(self.fc1(self.layer_norm2(x)), approximate="tanh"))


class SiglipVisionTower(nn.Module):
    def __init__(self, v: Dict[str, Any], dtype):
        super().__init__()
        h, p, s = v["hidden_size"], v["patch_size"], v["image_size"]
        eps = v.get("layer_norm_eps", 1e-6)
        self.patch_embedding = nn.Conv2d(3, h, p, p, dtype=dtype)
        self.position_embedding = nn.Embedding((s // p) ** 2, h, dtype=dtype)
        self.layers = nn.ModuleList([SiglipLayer(h, v["intermediate_size"], v["num_attention_heads"], eps, dtype) for _ in range(v["num_hidden_layers"])])
        self.post_layernorm = nn.LayerNorm(h, eps=eps, dtype=dtype)

    def forward(self, pixel_values: torch.Tensor) -> torch.Tensor:
        x = self.patch_embedding(pixel_values.to(self.patch_embedding.weight.dtype)).flatten(2).transpose(1, 2) + self.position_embedding.weight[None]
        for l in self.layers:
            x = l(x)
        return self.post_layernorm(x)


class Gemma3Projector(nn.Module):
    def __init__(self, v: Dict[str, Any], text_hidden: int, dtype):
        super().__init__()
        vh = v["hidden_size"]
        self.mm_input_projection_weight = nn.Parameter(torch.zeros(vh, text_hidden, dtype=dtype), requires_grad=False)
        self.mm_soft_emb_norm = nn.Parameter(torch.zeros(vh, dtype=dtype), requires_grad=False)          # Gemma norm: (1 + w)
        self.eps = v.get("layer_norm_eps", 1e-6)
        self.side = v["image_size"] // v["patch_size"]
        self.kernel = self.side // int(round(v["tokens_per_image"] ** 0.5))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        B, _, H = x.shape
        y = F.avg_pool2d(x.transpose(1, 2).reshape(B, H, self.side, self.side), self.kernel, self.kernel).flatten(2).transpose(1, 2)
        f = y.float()
        f = f * torch.rsqrt(f.pow(2).mean(-1, keepdim=True) + self.eps) * (1.0 + self.mm_soft_emb_norm.float())
        return (f.to(x.dtype) @ self.mm_input_projection_weight).flatten(0, 1)


def image_block_masks(fb: ForwardBatch, rows: torch.Tensor, window: int) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """flattened per-request [q_len, kv_len] masks for this extend step: (global, sliding).  ``rows``: flat indices of the image tokens in
    the step; consecutive rows inside one request form one image block."""
    T = int(fb.positions.shape[0])
    dev = fb.positions.device
    grp = torch.zeros(T, dtype=torch.long, device=dev)
    if rows.numel():
        r = rows.long().sort().values
        cu = torch.as_tensor(np.concatenate([[0], np.cumsum(np.asarray(fb.q_lens_cpu))]), device=dev)
        new = torch.ones_like(r, dtype=torch.bool)
        new[1:] = (r[1:] != r[:-1] + 1) | (torch.searchsorted(cu, r[1:], right=True) != torch.searchsorted(cu, r[:-1], right=True))
        grp[r] = torch.cumsum(new.long(), 0)
    g_parts, s_parts, off = [], [], 0
    for L, n in zip(np.asarray(fb.seq_lens_cpu).tolist(), np.asarray(fb.q_lens_cpu).tolist()):
        qp = torch.arange(L - n, L, device=dev)[:, None]
        kp = torch.arange(L, device=dev)[None, :]
        causal = kp <= qp
        gq = grp[off:off + n]
        same = torch.zeros(n, L, dtype=torch.bool, device=dev)
        same[:, L - n:] = (gq[:, None] == gq[None, :]) & (gq[:, None] > 0)
        g_parts.append((causal | same).reshape(-1))
        if window > 0:
            s_parts.append(((causal & (kp > qp - window)) | same).reshape(-1))
        off += n
    return torch.cat(g_parts), (torch.cat(s_parts) if s_parts else None)


class Gemma3ForConditionalGeneration(Gemma3ForCausalLM):
    def __init__(self, cfg: ModelConfig, quant_name: Optional[str] = None, device="cpu"):
        super().__init__(cfg, quant_name, device)
        v = cfg.vision or {}
        dt = _DT[cfg.dtype]
        if not v or v.get("skip_tower"):
            self.vision_tower = self.projector = None
        else:
            self.vision_tower = SiglipVisionTower(v, dt)
            self.projector = Gemma3Projector(v, cfg.hidden_size, dt)
        self.num_image_tokens = int(v.get("tokens_per_image", 256))
        self.window = int(cfg.sliding_window or 0)

    @torch.no_grad()
    def encode_images(self, pixel_values: torch.Tensor) -> torch.Tensor:
        return self.projector(self.vision_tower(pixel_values))

    def forward(self, fb: ForwardBatch, inputs_embeds=None):
        if fb.mm_embeds is not None and fb.cu_q is not None and fb.custom_mask is None:
            fb.custom_mask, fb.custom_mask_sliding = image_block_masks(fb, fb.mm_embeds[0], self.window)
        return super().forward(fb, inputs_embeds)

    def load_weights(self, weights: Iterable[Tuple[str, torch.Tensor]]) -> List[str]:
        own = dict(self.named_parameters())
        rest, loaded = [], []
        for name, w in weights:
            n = name[len("model."):] if name.startswith(("model.vision_tower.", "model.multi_modal_projector.", "model.language_model.")) else name
            if n.startswith("language_model."):
                n = n[len("language_model."):]
                rest.append((n if n.startswith(("model.", "lm_head.")) else "model." + n, w))
            elif n.startswith(("vision_tower.", "multi_modal_projector.")):
                if self.vision_tower is None:
                    continue
                n = n.replace("vision_tower.vision_model.embeddings.", "vision_tower.").replace("vision_tower.vision_model.encoder.layers.", "vision_tower.layers.") \
                    .replace("vision_tower.vision_model.post_layernorm", "vision_tower.post_layernorm").replace(".self_attn.", ".").replace(".mlp.fc", ".fc") \
                    .replace("multi_modal_projector.mm_soft_emb_norm.weight", "projector.mm_soft_emb_norm").replace("multi_modal_projector.", "projector.")
                if n in own:
                    own[n].data.copy_(w.reshape(own[n].shape))
                    loaded.append(name)
            else:
                rest.append((name, w))
        return loaded + super().load_weights(rest)
