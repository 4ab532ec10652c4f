"""LLaVA-1.5: CLIP ViT-L/14-336 vision tower + 2-layer MLP projector + Llama decoder.

Parity: reference ``python/sglang/srt/models/llava.py:63-854`` (image features from ``select_layer``, CLS dropped,
projector, features spliced over the image-pad token positions) and ``managers/mm_utils.py:609``
(``general_mm_embed_routine``).  Image pad tokens are replaced with values derived from the image hash before they
reach the scheduler so the radix cache never confuses two different images (reference ``schedule_batch.py:220,392``).
"""
from __future__ import annotations

from typing import Any, Dict, Iterable, List, Optional, Tuple

import torch
import torch.nn.functional as F
from torch import nn

from ..configs.model_config import ModelConfig
from .llama import _DT, LlamaForCausalLM


class CLIPEncoderLayer(nn.Module):
    def __init__(self, h: int, inter: int, heads: int, dtype):
        super().__init__()
        self.heads = heads
        self.layer_norm1 = nn.LayerNorm(h, dtype=dtype)
        self.layer_norm2 = nn.LayerNorm(h, dtype=dtype)
        self.q_proj = nn.Linear(h, h, dtype=dtype)
        self.k_proj = nn.Linear(h, h, dtype=dtype)
        self.v_proj = nn.Linear(h, h, dtype=dtype)
        self.out_proj = nn.Linear(h, h, dtype=dtype)
        self.fc1 = nn.Linear(h, inter, dtype=dtype)
        self.fc2 = nn.Linear(inter, h, dtype=dtype)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        B, N, H = x.shape
        y = self.layer_norm1(x)
        q = self.q_proj(y).view(B, N, self.heads, -1).transpose(1, 2)
        k = self.k_proj(y).view(B, N, self.heads, -1).transpose(1, 2)
        v = self.v_proj(y).view(B, N, self.heads, -1).transpose(1, 2)
        a = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B, N, H)
        x = x + self.out_proj(a)
        y = self.layer_norm2(x)
        y = self.fc1(y)
        y = y * torch.sigmoid(1.702 * y)   # quick_gelu
        return x + self.fc2(y)


class CLIPVisionTower(nn.Module):
    def __init__(self, v: Dict[str, Any], dtype):
        super().__init__()
        h, p, s = v["hidden_size"], v["patch_size"], v["image_size"]
        self.patch_size, self.image_size = p, s
        self.num_patches = (s // p) ** 2
        self.select_layer = v.get("select_layer", -2)
        self.class_embedding = nn.Parameter(torch.zeros(h, dtype=dtype), requires_grad=False)
        self.patch_embedding = nn.Conv2d(3, h, p, p, bias=False, dtype=dtype)
        self.position_embedding = nn.Embedding(self.num_patches + 1, h, dtype=dtype)
        self.pre_layrnorm = nn.LayerNorm(h, dtype=dtype)
        n_layers = v["num_hidden_layers"]
        keep = n_layers + self.select_layer + 1 if self.select_layer < 0 else self.select_layer
        self.layers = nn.ModuleList([CLIPEncoderLayer(h, v["intermediate_size"], v["num_attention_heads"], dtype)
                                     for _ in range(keep)])

    @torch.no_grad()
    def forward(self, pixel_values: torch.Tensor) -> torch.Tensor:
        """pixel_values [B,3,S,S] -> patch features [B, num_patches, h] (CLS dropped)"""
        x = self.patch_embedding(pixel_values.to(self.patch_embedding.weight.dtype)).flatten(2).transpose(1, 2)
        cls = self.class_embedding.expand(x.shape[0], 1, -1)
        x = torch.cat([cls, x], dim=1) + self.position_embedding.weight[None]
        x = self.pre_layrnorm(x)
        for l in self.layers:
            x = l(x)
        return x[:, 1:]


class LlavaForConditionalGeneration(LlamaForCausalLM):
    def __init__(self, cfg: ModelConfig, quant_name: Optional[str] = None, device="cpu"):
        super().__init__(cfg, quant_name, device)
        dtype = _DT[cfg.dtype]
        if cfg.vision.get("skip_tower"):      # EPD --language-only: image rows arrive as embeddings from an encoder server
            self.vision_tower = self.projector = None
            self.num_image_tokens = (cfg.vision["image_size"] // cfg.vision["patch_size"]) ** 2
            return
        self.vision_tower = CLIPVisionTower(cfg.vision, dtype)
        vh = cfg.vision["hidden_size"]
        self.projector = nn.Sequential(nn.Linear(vh, cfg.hidden_size, dtype=dtype), nn.GELU(),
                                       nn.Linear(cfg.hidden_size, cfg.hidden_size, dtype=dtype))
        self.num_image_tokens = self.vision_tower.num_patches

    @torch.no_grad()
    def encode_images(self, pixel_values: torch.Tensor) -> torch.Tensor:
        """[N,3,S,S] -> [N*num_patches, hidden]"""
        f = self.vision_tower(pixel_values)
        return self.projector(f).flatten(0, 1)

    def load_weights(self, weights: Iterable[Tuple[str, torch.Tensor]]) -> List[str]:
        own = dict(self.named_parameters())
        rest, loaded = [], []
        for name, w in weights:
            n = name
            # checkpoint dialects: llava-1.5-hf ("language_model.model.*", "vision_tower.*", "multi_modal_projector.*") and the newer
            # transformers layout that nests all three under "model."
            for pre in ("vision_tower.", "multi_modal_projector.", "language_model."):
                if n.startswith("model." + pre):
                    n = n[len("model."):]
                    break
            if n.startswith("language_model."):
                n = n[len("language_model."):]
                rest.append((n if n.startswith(("model.", "lm_head.")) else "model." + n, w))
                continue
            n = n.replace("vision_tower.vision_model.embeddings.", "vision_tower.").replace(
                "vision_tower.vision_model.encoder.layers.", "vision_tower.layers.").replace(
                "vision_tower.vision_model.pre_layrnorm", "vision_tower.pre_layrnorm").replace(
                ".self_attn.", ".").replace(".mlp.fc", ".fc").replace("multi_modal_projector.linear_1", "projector.0").replace(
                "multi_modal_projector.linear_2", "projector.2")
            if n in own:
                own[n].data.copy_(w.reshape(own[n].shape))
                loaded.append(name)
            elif not n.startswith(("vision_tower", "projector.", "multi_modal_projector")):
                rest.append((name, w))
        return loaded + super().load_weights(rest)
