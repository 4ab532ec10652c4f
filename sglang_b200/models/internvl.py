"""InternVL (2.5 / 3): InternViT tower (CLS token, absolute positions, layer-scale residuals) -> pixel-shuffle 2x2 -> MLP projector -> a
Qwen2 / Llama / InternLM2 decoder.  Images are cut into 448-px tiles (dynamic aspect-ratio grid + thumbnail), every tile gives 256 tokens.

Parity: reference ``python/sglang/srt/models/internvl.py`` + ``multimodal/processors/internvl.py``; checked against HF ``InternVLForConditionalGeneration``.
"""
from __future__ import annotations

from typing import Any, Dict, Iterable, List, Optional, Tuple

import torch
import torch.nn.functional as F
from torch import nn

from ..configs.model_config import ModelConfig
from .llama import _DT, LlamaForCausalLM


class _Norm(nn.Module):
    def __init__(self, n: int, kind: str, eps: float, dtype):
        super().__init__()
        self.rms = kind == "rms_norm"
        self.weight = nn.Parameter(torch.ones(n, dtype=dtype), requires_grad=False)
        self.bias = None if self.rms else nn.Parameter(torch.zeros(n, dtype=dtype), requires_grad=False)
        self.eps = eps

    def forward(self, x):
        if self.rms:
            f = x.float()
            return (f * torch.rsqrt(f.pow(2).mean(-1, keepdim=True) + self.eps)).to(x.dtype) * self.weight
        return F.layer_norm(x, (x.shape[-1],), self.weight, self.bias, self.eps)


class InternVisionLayer(nn.Module):
    def __init__(self, v: Dict[str, Any], dtype):
        super().__init__()
        H, self.heads = v["hidden_size"], v["num_attention_heads"]
        b = v.get("attention_bias", False)
        self.q_proj, self.k_proj, self.v_proj = (nn.Linear(H, H, bias=b, dtype=dtype) for _ in range(3))
        self.projection_layer = nn.Linear(H, H, dtype=dtype)
        self.qk = v.get("use_qk_norm", False)
        if self.qk:
            self.q_norm, self.k_norm = _Norm(H, "rms_norm", 1e-6, dtype), _Norm(H, "rms_norm", 1e-6, dtype)
        kind, eps = v.get("norm_type", "layer_norm"), v.get("layer_norm_eps", 1e-6)
        self.layernorm_before, self.layernorm_after = _Norm(H, kind, eps, dtype), _Norm(H, kind, eps, dtype)
        self.fc1, self.fc2 = nn.Linear(H, v["intermediate_size"], dtype=dtype), nn.Linear(v["intermediate_size"], H, dtype=dtype)
        self.lambda_1 = nn.Parameter(torch.ones(H, dtype=dtype), requires_grad=False)
        self.lambda_2 = nn.Parameter(torch.ones(H, dtype=dtype), requires_grad=False)
        self.act = F.gelu if v.get("hidden_act", "gelu") == "gelu" else (lambda t: t * torch.sigmoid(1.702 * t))

    def forward(self, x):                       # [B, S, H]
        B, S, H = x.shape
        h = self.layernorm_before(x)
        q, k, v = self.q_proj(h), self.k_proj(h), self.v_proj(h)
        if self.qk:
            q, k = self.q_norm(q), self.k_norm(k)
        sp = lambda t: t.view(B, S, self.heads, H // self.heads).transpose(1, 2)
        a = F.scaled_dot_product_attention(sp(q), sp(k), sp(v)).transpose(1, 2).reshape(B, S, H)
        x = x + self.lambda_1 * self.projection_layer(a)
        return x + self.lambda_2 * self.fc2(self.act(self.fc1(self.layernorm_after(x))))


class InternVisionTower(nn.Module):
    def __init__(self, v: Dict[str, Any], dtype):
        super().__init__()
        H, p = v["hidden_size"], v["patch_size"]
        n = (v["image_size"] // p) ** 2
        self.cls_token = nn.Parameter(torch.zeros(1, 1, H, dtype=dtype), requires_grad=False)
        self.patch_embedding = nn.Conv2d(3, H, p, p, dtype=dtype)
        self.position_embeddings = nn.Parameter(torch.zeros(1, n + 1, H, dtype=dtype), requires_grad=False)
        self.layers = nn.ModuleList([InternVisionLayer(v, dtype) for _ in range(v["num_hidden_layers"])])
        self.select_layer = v.get("select_layer", -1)

    def forward(self, px: torch.Tensor) -> torch.Tensor:       # [N, 3, S, S] -> [N, patches, H] (CLS dropped)
        x = self.patch_embedding(px).flatten(2).transpose(1, 2)
        x = torch.cat([self.cls_token.expand(x.shape[0], -1, -1), x], 1) + self.position_embeddings
        n = len(self.layers) if self.select_layer == -1 else (self.select_layer % (len(self.layers) + 1))
        for l in self.layers[:n]:
            x = l(x)
        return x[:, 1:]


def pixel_shuffle(x: torch.Tensor, r: float = 0.5) -> torch.Tensor:
    """[B, W, H, C] -> [B, W*r, H*r, C/r^2] (the exact reshape/permute sequence of the reference implementation)"""
    B, W, Hh, C = x.shape
    x = x.view(B, W, int(Hh * r), int(C / r)).permute(0, 2, 1, 3).contiguous()
    x = x.view(B, int(Hh * r), int(W * r), int(C / (r * r)))
    return x.permute(0, 2, 1, 3).contiguous()


class InternVLForConditionalGeneration(LlamaForCausalLM):
    def __init__(self, cfg: ModelConfig, quant_name: Optional[str] = None, device="cpu"):
        super().__init__(cfg, quant_name, device)
        v, dtype = cfg.vision, _DT[cfg.dtype]
        self.ratio = v.get("downsample_ratio", 0.5)
        m = v["hidden_size"] * int(1 / self.ratio) ** 2
        self.num_image_tokens = int((v["image_size"] // v["patch_size"]) ** 2 * self.ratio ** 2)
        if v.get("skip_tower"):
            self.vision_tower = self.projector = None
            return
        self.vision_tower = InternVisionTower(v, dtype)
        self.projector = nn.ModuleDict({"layer_norm": nn.LayerNorm(m, dtype=dtype), "linear_1": nn.Linear(m, cfg.hidden_size, dtype=dtype),
                                        "linear_2": nn.Linear(cfg.hidden_size, cfg.hidden_size, dtype=dtype)})

    @torch.no_grad()
    def encode_images(self, pixel_values: torch.Tensor) -> torch.Tensor:
        f = self.vision_tower(pixel_values)
        B, n, C = f.shape
        s = int(n ** 0.5)
        f = pixel_shuffle(f.reshape(B, s, s, C), self.ratio).reshape(B, -1, int(C / self.ratio ** 2))
        p = self.projector
        return p["linear_2"](F.gelu(p["linear_1"](p["layer_norm"](f)))).flatten(0, 1)

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
                n = n.replace("vision_tower.embeddings.patch_embeddings.projection.", "vision_tower.patch_embedding.").replace(
                    "vision_tower.embeddings.", "vision_tower.").replace("vision_tower.encoder.layer.", "vision_tower.layers.").replace(
                    ".attention.", ".").replace(".mlp.fc", ".fc").replace("multi_modal_projector.", "projector.")
                if n in own:
                    own[n].data.copy_(w.reshape(own[n].shape))
                    loaded.append(name)
                continue
            rest.append((name, w))
        return loaded + super().load_weights(rest)
