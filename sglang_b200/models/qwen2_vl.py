"""Qwen2-VL: dynamic-resolution vision transformer (2-D rotary, per-image full attention, 2x2 patch merger) feeding a Qwen2 decoder that
uses multimodal rotary positions (M-RoPE: temporal / height / width position per token, one rotary section each).

Parity: reference ``python/sglang/srt/models/qwen2_vl.py`` (``Qwen2VisionTransformer``, ``Qwen2VLForConditionalGeneration``),
``layers/rotary_embedding`` mrope variant, ``multimodal/processors/qwen_vl.py`` (smart resize, patch flattening, rope index).

Engine integration without touching the kernels:

* M-RoPE = ordinary RoPE whose cos/sin row is assembled per token from three positions.  ``prepare_mrope`` builds that [T, rot] table once
  per forward (gathering the three sections from the model's cos/sin cache) and the fused RoPE + KV-write kernel simply indexes it by row
  number (``fb.rope_override``).  Text-only tokens have three equal positions and reduce to plain RoPE;
* decode positions are ``sequence position + delta`` with a per-request delta (images occupy fewer rotary positions than tokens) kept in a
  device vector indexed by request slot — pure device arithmetic, graph-capturable;
* image rows are spliced into the prompt embeddings by the shared ``multimodal/embed.py`` path (variable token count per image).
"""
from __future__ import annotations

from typing import Any, Dict, Iterable, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

from ..configs.model_config import ModelConfig
from ..model_executor.forward_batch import ForwardBatch
from .llama import _DT, LlamaForCausalLM


# ------------------------------------------------------------------------------------------------ vision tower
def _rot_half(x):
    a, b = x[..., : x.shape[-1] // 2], x[..., x.shape[-1] // 2:]
    return torch.cat([-b, a], -1)


class VisionBlock(nn.Module):
    def __init__(self, dim: int, heads: int, mlp_ratio: float, dtype):
        super().__init__()
        self.heads = heads
        self.norm1 = nn.LayerNorm(dim, eps=1e-6, dtype=dtype)
        self.norm2 = nn.LayerNorm(dim, eps=1e-6, dtype=dtype)
        self.qkv = nn.Linear(dim, 3 * dim, dtype=dtype)
        self.proj = nn.Linear(dim, dim, dtype=dtype)
        self.fc1 = nn.Linear(dim, int(dim * mlp_ratio), dtype=dtype)
        self.fc2 = nn.Linear(int(dim * mlp_ratio), dim, dtype=dtype)

    def forward(self, x: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor, seg: List[int]) -> torch.Tensor:
        N, D = x.shape
        q, k, v = self.qkv(self.norm1(x)).view(N, 3, self.heads, D // self.heads).unbind(1)
        q = (q.float() * cos[:, None] + _rot_half(q.float()) * sin[:, None]).to(x.dtype)
        k = (k.float() * cos[:, None] + _rot_half(k.float()) * sin[:, None]).to(x.dtype)
        outs, s = [], 0
        for n in seg:           # full attention inside one image (frame group), nothing across images
            o = F.scaled_dot_product_attention(q[s:s + n].transpose(0, 1), k[s:s + n].transpose(0, 1), v[s:s + n].transpose(0, 1))
            outs.append(o.transpose(0, 1).reshape(n, D))
            s += n
        x = x + self.proj(torch.cat(outs, 0))
        h = self.fc1(self.norm2(x))
        return x + self.fc2(h * torch.sigmoid(1.702 * h))            # quick-GELU


class Qwen2VisionTransformer(nn.Module):
    def __init__(self, v: Dict[str, Any], dtype):
        super().__init__()
        self.embed_dim, self.heads, self.patch, self.merge, self.tps = v["embed_dim"], v["num_heads"], v["patch_size"], v["spatial_merge_size"], \
            v.get("temporal_patch_size", 2)
        self.patch_embed = nn.Linear(3 * self.tps * self.patch * self.patch, self.embed_dim, bias=False, dtype=dtype)      # Conv3d with stride = kernel
        self.blocks = nn.ModuleList([VisionBlock(self.embed_dim, self.heads, v.get("mlp_ratio", 4), dtype) for _ in range(v["depth"])])
        self.ln_q = nn.LayerNorm(self.embed_dim, eps=1e-6, dtype=dtype)
        m = self.embed_dim * self.merge ** 2
        self.merger = nn.Sequential(nn.Linear(m, m, dtype=dtype), nn.GELU(), nn.Linear(m, v["hidden_size"], dtype=dtype))
        hd = self.embed_dim // self.heads
        self.register_buffer("inv_freq", 1.0 / (10000.0 ** (torch.arange(0, hd // 2, 2, dtype=torch.float32) / (hd // 2))), persistent=False)

    def rot_pos_emb(self, grid_thw: np.ndarray) -> Tuple[torch.Tensor, torch.Tensor]:
        ids = []
        m = self.merge
        for t, h, w in grid_thw.tolist():
            hp = np.arange(h)[:, None].repeat(w, 1).reshape(h // m, m, w // m, m).transpose(0, 2, 1, 3).reshape(-1)
            wp = np.arange(w)[None, :].repeat(h, 0).reshape(h // m, m, w // m, m).transpose(0, 2, 1, 3).reshape(-1)
            ids.append(np.tile(np.stack([hp, wp], -1), (t, 1)))
        ids = torch.from_numpy(np.concatenate(ids)).to(self.inv_freq.device)
        freqs = torch.outer(torch.arange(int(grid_thw[:, 1:].max()), dtype=torch.float32, device=self.inv_freq.device), self.inv_freq)
        emb = freqs[ids].flatten(1)                     # [N, head_dim / 2]: h frequencies | w frequencies
        emb = torch.cat([emb, emb], -1)
        return emb.cos(), emb.sin()

    def forward(self, patches: torch.Tensor, grid_thw: np.ndarray) -> torch.Tensor:
        """patches [N, 3*tps*p*p] (merge-block order), grid_thw [n_img, 3] -> [N / merge^2, hidden]"""
        x = self.patch_embed(patches)
        cos, sin = self.rot_pos_emb(grid_thw)
        seg = []
        for t, h, w in grid_thw.tolist():
            seg += [h * w] * t
        for blk in self.blocks:
            x = blk(x, cos, sin, seg)
        return self.merger(self.ln_q(x).view(-1, self.embed_dim * self.merge ** 2))


class _VRMS(nn.Module):
    def __init__(self, n: int, dtype, eps: float = 1e-6):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(n, dtype=dtype), requires_grad=False)
        self.eps = eps

    def forward(self, x):
        f = x.float()
        return (f * torch.rsqrt(f.pow(2).mean(-1, keepdim=True) + self.eps)).to(x.dtype) * self.weight


class VisionBlock25(nn.Module):
    """Qwen2.5-VL block: RMSNorm, biased SwiGLU MLP; attention is full inside the segments it is given (windows or whole frames)"""

    def __init__(self, dim: int, heads: int, inter: int, dtype):
        super().__init__()
        self.heads = heads
        self.norm1, self.norm2 = _VRMS(dim, dtype), _VRMS(dim, dtype)
        self.qkv, self.proj = nn.Linear(dim, 3 * dim, dtype=dtype), nn.Linear(dim, dim, dtype=dtype)
        self.gate_proj, self.up_proj = nn.Linear(dim, inter, dtype=dtype), nn.Linear(dim, inter, dtype=dtype)
        self.down_proj = nn.Linear(inter, dim, dtype=dtype)

    def forward(self, x, cos, sin, seg: List[int]):
        N, D = x.shape
        q, k, v = self.qkv(self.norm1(x)).view(N, 3, self.heads, D // self.heads).unbind(1)
        q = (q.float() * cos[:, None] + _rot_half(q.float()) * sin[:, None]).to(x.dtype)
        k = (k.float() * cos[:, None] + _rot_half(k.float()) * sin[:, None]).to(x.dtype)
        outs, s = [], 0
        for n in seg:
            o = F.scaled_dot_product_attention(q[s:s + n].transpose(0, 1), k[s:s + n].transpose(0, 1), v[s:s + n].transpose(0, 1))
            outs.append(o.transpose(0, 1).reshape(n, D))
            s += n
        x = x + self.proj(torch.cat(outs, 0))
        h = self.norm2(x)
        return x + self.down_proj(F.silu(self.gate_proj(h)) * self.up_proj(h))


class Qwen25VisionTransformer(Qwen2VisionTransformer):
    """Qwen2.5-VL tower: most blocks attend inside ``window_size`` pixel windows (patches are re-ordered window-major for that), the blocks
    listed in ``fullatt_block_indexes`` attend over the whole frame; merger with RMSNorm"""

    def __init__(self, v: Dict[str, Any], dtype):
        nn.Module.__init__(self)
        self.embed_dim, self.heads, self.patch, self.merge, self.tps = v["embed_dim"], v["num_heads"], v["patch_size"], v["spatial_merge_size"], \
            v.get("temporal_patch_size", 2)
        self.window, self.full_blocks = v.get("window_size", 112), set(v.get("fullatt_block_indexes", []))
        self.patch_embed = nn.Linear(3 * self.tps * self.patch * self.patch, self.embed_dim, bias=False, dtype=dtype)
        self.blocks = nn.ModuleList([VisionBlock25(self.embed_dim, self.heads, v["intermediate_size"], dtype) for _ in range(v["depth"])])
        self.ln_q = _VRMS(self.embed_dim, dtype)
        m = self.embed_dim * self.merge ** 2
        self.merger = nn.Sequential(nn.Linear(m, m, dtype=dtype), nn.GELU(), nn.Linear(m, v["hidden_size"], dtype=dtype))
        hd = self.embed_dim // self.heads
        self.register_buffer("inv_freq", 1.0 / (10000.0 ** (torch.arange(0, hd // 2, 2, dtype=torch.float32) / (hd // 2))), persistent=False)

    def window_index(self, grid_thw: np.ndarray):
        """order of the merge-blocks that makes every window contiguous + the window segment lengths in patches"""
        order, seg, base = [], [], 0
        ws = self.window // self.merge // self.patch
        unit = self.merge ** 2
        for t, h, w in grid_thw.tolist():
            lh, lw = h // self.merge, w // self.merge
            idx = np.arange(t * lh * lw).reshape(t, lh, lw)
            ph, pw = ws - lh % ws, ws - lw % ws
            pad = np.pad(idx, ((0, 0), (0, ph), (0, pw)), constant_values=-100)
            nh, nw = (lh + ph) // ws, (lw + pw) // ws
            pad = pad.reshape(t, nh, ws, nw, ws).transpose(0, 1, 3, 2, 4).reshape(t, nh * nw, ws, ws)
            cnt = (pad != -100).sum((2, 3)).reshape(-1)
            flat = pad.reshape(-1)
            order.append(flat[flat != -100] + base)
            seg += [int(c) * unit for c in cnt if c > 0]
            base += t * lh * lw
        return np.concatenate(order), seg

    def forward(self, patches: torch.Tensor, grid_thw: np.ndarray) -> torch.Tensor:
        x = self.patch_embed(patches)
        cos, sin = self.rot_pos_emb(grid_thw)
        order, win_seg = self.window_index(grid_thw)
        unit = self.merge ** 2
        oi = torch.from_numpy(order).to(x.device)
        regroup = lambda t: t.view(-1, unit, t.shape[-1])[oi].reshape(-1, t.shape[-1])
        x, cos, sin = regroup(x), regroup(cos), regroup(sin)
        full_seg = []
        for t, h, w in grid_thw.tolist():
            full_seg += [h * w] * t
        for i, blk in enumerate(self.blocks):
            x = blk(x, cos, sin, full_seg if i in self.full_blocks else win_seg)
        y = self.merger(self.ln_q(x).view(-1, self.embed_dim * unit))
        return y[torch.argsort(oi)]


class VisionBlock3(nn.Module):
    """Qwen3-VL block: LayerNorm, biased qkv / proj, plain (non-gated) tanh-GELU MLP, full attention inside one image"""

    def __init__(self, dim: int, heads: int, inter: int, dtype):
        super().__init__()
        self.heads = heads
        self.norm1, self.norm2 = nn.LayerNorm(dim, eps=1e-6, dtype=dtype), nn.LayerNorm(dim, eps=1e-6, dtype=dtype)
        self.qkv = nn.Linear(dim, 3 * dim, dtype=dtype)
        self.proj = nn.Linear(dim, dim, dtype=dtype)
        self.fc1 = nn.Linear(dim, inter, dtype=dtype)
        self.fc2 = nn.Linear(inter, dim, dtype=dtype)

    def forward(self, x, cos, sin, seg: List[int]):
        N, D = x.shape
        q, k, v = self.qkv(self.norm1(x)).view(N, 3, self.heads, D // self.heads).unbind(1)
        q = (q.float() * cos[:, None] + _rot_half(q.float()) * sin[:, None]).to(x.dtype)
        k = (k.float() * cos[:, None] + _rot_half(k.float()) * sin[:, None]).to(x.dtype)
        outs, s = [], 0
        for n in seg:
            o = F.scaled_dot_product_attention(q[s:s + n].transpose(0, 1), k[s:s + n].transpose(0, 1), v[s:s + n].transpose(0, 1))
            outs.append(o.transpose(0, 1).reshape(n, D))
            s += n
        x = x + self.proj(torch.cat(outs, 0))
        return x + self.fc2(F.gelu(self.fc1(self.norm2(x)), approximate="tanh"))


class _Merger3(nn.Module):
    def __init__(self, dim: int, unit: int, out: int, postshuffle: bool, dtype):
        super().__init__()
        self.m, self.post = dim * unit, postshuffle
        self.norm = nn.LayerNorm(self.m if postshuffle else dim, eps=1e-6, dtype=dtype)
        self.fc1 = nn.Linear(self.m, self.m, dtype=dtype)
        self.fc2 = nn.Linear(self.m, out, dtype=dtype)

    def forward(self, x):
        x = self.norm(x.view(-1, self.m) if self.post else x).view(-1, self.m)
        return self.fc2(F.gelu(self.fc1(x)))


class Qwen3VisionTransformer(Qwen2VisionTransformer):
    """Qwen3-VL tower (reference ``python/sglang/srt/models/qwen3_vl.py``): biased patch embedding, a learned position grid that is bilinearly
    resampled to each image's patch grid, 2-D rotary full attention, and DeepStack — the outputs of ``deepstack_visual_indexes`` blocks go
    through their own mergers and are handed to the language model as extra feature columns (see ``Qwen3VLForConditionalGeneration``)."""

    def __init__(self, v: Dict[str, Any], dtype):
        nn.Module.__init__(self)
        self.embed_dim, self.heads, self.patch, self.merge, self.tps = v["embed_dim"], v["num_heads"], v["patch_size"], v["spatial_merge_size"], \
            v.get("temporal_patch_size", 2)
        self.patch_embed = nn.Linear(3 * self.tps * self.patch * self.patch, self.embed_dim, bias=True, dtype=dtype)
        self.pos_embed = nn.Embedding(v["num_position_embeddings"], self.embed_dim, dtype=dtype)
        self.side = int(v["num_position_embeddings"] ** 0.5)
        self.blocks = nn.ModuleList([VisionBlock3(self.embed_dim, self.heads, v["intermediate_size"], dtype) for _ in range(v["depth"])])
        unit = self.merge ** 2
        self.merger = _Merger3(self.embed_dim, unit, v["hidden_size"], False, dtype)
        self.ds_index = list(v.get("deepstack_visual_indexes") or [])
        self.deepstack_merger_list = nn.ModuleList([_Merger3(self.embed_dim, unit, v["hidden_size"], True, dtype) for _ in self.ds_index])
        hd = self.embed_dim // self.heads
        self.register_buffer("inv_freq", 1.0 / (10000.0 ** (torch.arange(0, hd // 2, 2, dtype=torch.float32) / (hd // 2))), persistent=False)

    def pos_interp(self, grid_thw: np.ndarray) -> torch.Tensor:
        """the learned [side, side] position grid resampled (bilinear, align-corners) to each image's [h, w] patch grid, in merge-block order"""
        W, m, n, dev = self.pos_embed.weight, self.merge, self.side, self.pos_embed.weight.device
        out = []
        for t, h, w in grid_thw.tolist():
            hs, ws = torch.linspace(0, n - 1, h, device=dev), torch.linspace(0, n - 1, w, device=dev)
            h0, w0 = hs.int(), ws.int()
            h1, w1 = (h0 + 1).clamp(max=n - 1), (w0 + 1).clamp(max=n - 1)
            dh, dw = (hs - h0)[:, None], (ws - w0)[None, :]
            g = lambda a, b: W[(a[:, None] * n + b[None, :]).long()].float()               # [h, w, C]
            e = g(h0, w0) * ((1 - dh) * (1 - dw))[..., None] + g(h0, w1) * ((1 - dh) * dw)[..., None] \
                + g(h1, w0) * (dh * (1 - dw))[..., None] + g(h1, w1) * (dh * dw)[..., None]
            e = e.view(h // m, m, w // m, m, -1).permute(0, 2, 1, 3, 4).reshape(h * w, -1)
            out.append(e.repeat(t, 1))
        return torch.cat(out).to(W.dtype)

    def forward(self, patches: torch.Tensor, grid_thw: np.ndarray) -> torch.Tensor:
        """-> [N / merge^2, hidden * (1 + len(deepstack))]: the merged features, then one column block per DeepStack level"""
        x = self.patch_embed(patches) + self.pos_interp(grid_thw)
        cos, sin = self.rot_pos_emb(grid_thw)
        seg = []
        for t, h, w in grid_thw.tolist():
            seg += [h * w] * t
        deep = []
        for i, blk in enumerate(self.blocks):
            x = blk(x, cos, sin, seg)
            if i in self.ds_index:
                deep.append(self.deepstack_merger_list[self.ds_index.index(i)](x))
        return torch.cat([self.merger(x)] + deep, -1)


# ------------------------------------------------------------------------------------------------ M-RoPE helpers
def mrope_index(ids: List[int], grids: List[Tuple[int, int, int]], image_token: int, merge: int) -> Tuple[np.ndarray, int]:
    """[3, len(ids)] rotary positions (temporal, height, width) and the delta ``next position - len(ids)`` for decode (reference
    ``get_rope_index``).  ``grids``: (t, h, w) in PATCH units for every image, in prompt order; its placeholder run has t*h*w/merge^2 tokens."""
    pos = np.zeros((3, len(ids)), dtype=np.int64)
    i, st, g = 0, 0, 0
    arr = np.asarray(ids)
    n = len(ids)
    while i < n:
        if arr[i] == image_token and g < len(grids):
            t, h, w = grids[g]
            h, w = h // merge, w // merge
            cnt = t * h * w
            ti = np.arange(t)[:, None, None].repeat(h, 1).repeat(w, 2).reshape(-1)
            hi = np.arange(h)[None, :, None].repeat(t, 0).repeat(w, 2).reshape(-1)
            wi = np.arange(w)[None, None, :].repeat(t, 0).repeat(h, 1).reshape(-1)
            pos[:, i:i + cnt] = np.stack([ti, hi, wi]) + st
            st += max(t, h, w)
            i += cnt
            g += 1
        else:
            pos[:, i] = st
            st += 1
            i += 1
    return pos, st - n


def prepare_mrope(fb: ForwardBatch, cos_sin: torch.Tensor, pos3: torch.Tensor, sections: List[int], interleaved: bool = False) -> None:
    """assemble the per-token cos/sin rows: rotary dims [0, s0) follow the temporal position, [s0, s0+s1) height, the rest width
    (``interleaved``, Qwen3-VL: pair j follows height when j % 3 == 1 and j < 3 s1, width when j % 3 == 2 and j < 3 s2, else time)"""
    half = cos_sin.shape[1] // 2
    if interleaved:
        j = torch.arange(half, device=cos_sin.device)
        sel = torch.where((j % 3 == 1) & (j < 3 * sections[1]), 1, torch.where((j % 3 == 2) & (j < 3 * sections[2]), 2, 0))
    else:
        sel = torch.repeat_interleave(torch.arange(3, device=cos_sin.device), torch.tensor(sections, device=cos_sin.device))[:half]  # [half] -> axis
    p = pos3.long()[sel, :].t()                                             # [T, half]: which position feeds each rotary pair
    d = torch.arange(half, device=cos_sin.device)[None, :]
    rows = torch.cat([cos_sin[p, d], cos_sin[p, d + half]], -1)             # [T, rot]
    fb.rope_override = (rows.contiguous(), torch.arange(rows.shape[0], dtype=torch.int32, device=rows.device))


class Qwen2VLForConditionalGeneration(LlamaForCausalLM):
    is_mrope = True

    def __init__(self, cfg: ModelConfig, quant_name: Optional[str] = None, device="cpu"):
        super().__init__(cfg, quant_name, device)
        v = cfg.vision
        tower = {"qwen2.5": Qwen25VisionTransformer, "qwen3": Qwen3VisionTransformer}.get(v.get("variant"), Qwen2VisionTransformer)
        if v.get("variant") == "glm4v":
            from .glm4v import Glm4vVisionTransformer as tower
        self.visual = None if v.get("skip_tower") else tower(v, _DT[cfg.dtype])
        self.vision_tower = self.visual            # name the shared multimodal plumbing looks for
        self.merge = v["spatial_merge_size"]
        self.sections = list((cfg.rope_scaling or {}).get("mrope_section", [cfg.head_dim // 2]))
        self.interleaved = bool((cfg.rope_scaling or {}).get("mrope_interleaved", False))
        self.mrope_delta: Optional[torch.Tensor] = None       # [max_slots] i32, set per request at prefill (ModelRunner allocates)

    # ---- vision
    @torch.no_grad()
    def encode_images(self, patches: torch.Tensor, grid_thw: Any = None) -> torch.Tensor:
        return self.visual(patches, np.asarray(grid_thw, dtype=np.int64).reshape(-1, 3))

    # ---- language model with M-RoPE
    def forward(self, fb: ForwardBatch, inputs_embeds=None):
        cs = self.model.layers[0].self_attn.attn.rope.cos_sin if len(self.model.layers) else None
        if cs is not None:
            pos3 = getattr(fb, "mrope_positions", None)
            if pos3 is None and self.mrope_delta is not None:      # decode (or text-only extend): all three axes advance together
                p = fb.positions.long() + self.mrope_delta[fb.req_slots.long()].long() if fb.cu_q is None else fb.positions.long()
                pos3 = p[None, :].expand(3, -1)
            if pos3 is not None:
                prepare_mrope(fb, cs, pos3.clamp(max=cs.shape[0] - 1), self.sections, self.interleaved)
        return self.model(fb, inputs_embeds)

    def load_weights(self, weights: Iterable[Tuple[str, torch.Tensor]]) -> List[str]:
        own = dict(self.named_parameters())
        rest, loaded = [], []
        for name, w in weights:
            if name.startswith("model.visual."):                 # transformers >= 4.52 nests both halves under ``model.``
                name = name[len("model."):]
            elif name.startswith("model.language_model."):
                name = "model." + name[len("model.language_model."):]
            if name.startswith("visual."):
                if self.visual is None:
                    continue
                n = name.replace(".attn.qkv.", ".qkv.").replace(".attn.proj.", ".proj.").replace(".mlp.fc1.", ".fc1.").replace(".mlp.fc2.", ".fc2.") \
                    .replace(".mlp.gate_proj.", ".gate_proj.").replace(".mlp.up_proj.", ".up_proj.").replace(".mlp.down_proj.", ".down_proj.") \
                    .replace("visual.merger.ln_q.", "visual.ln_q.").replace("visual.merger.mlp.", "visual.merger.").replace("visual.patch_embed.proj.", "visual.patch_embed.")
                if n in own:
                    own[n].data.copy_(w.reshape(own[n].shape))
                    loaded.append(name)
                continue
            rest.append((name[len("language_model."):] if name.startswith("language_model.") else name, w))
        return loaded + super().load_weights(rest)


class Qwen2_5_VLForConditionalGeneration(Qwen2VLForConditionalGeneration):
    """same language model and M-RoPE; windowed vision tower (``vision['variant'] == 'qwen2.5'`` is set by the config loader)"""


class Qwen3VLForConditionalGeneration(Qwen2VLForConditionalGeneration):
    """Qwen3-VL (reference ``python/sglang/srt/models/qwen3_vl.py``): Qwen3 decoder (per-head q/k norm) with *interleaved* M-RoPE and DeepStack.
    The tower returns ``hidden * (1 + levels)`` columns per visual token; the first block replaces the placeholder embeddings as usual, level
    ``l`` is added to the residual stream at the same rows after decoder layer ``l`` (``LlamaModel``'s ``post_layer_hook``).  Chunked prefill
    and prefix-cache hits work unchanged because the shared splice path already hands over only the rows of the current chunk."""

    def forward(self, fb: ForwardBatch, inputs_embeds=None):
        H = self.cfg.hidden_size
        if fb.mm_embeds is not None and fb.mm_embeds[1].shape[1] > H:
            rows, emb = fb.mm_embeds
            levels = emb[:, H:].split(H, dim=1)
            fb.mm_embeds = (rows, emb[:, :H].contiguous())

            def hook(layer: int, h):
                if layer >= len(levels):
                    return h
                h = h.resolve() if hasattr(h, "resolve") else h
                h.index_add_(0, rows, levels[layer].to(h.dtype))
                return h

            fb.post_layer_hook = hook
        return super().forward(fb, inputs_embeds)

    def load_weights(self, weights: Iterable[Tuple[str, torch.Tensor]]) -> List[str]:
        def rename():
            for name, w in weights:
                yield name.replace(".mlp.linear_fc1.", ".fc1.").replace(".mlp.linear_fc2.", ".fc2.").replace(".linear_fc1.", ".fc1.") \
                    .replace(".linear_fc2.", ".fc2."), w
        return super().load_weights(rename())


def _make_qwen3_vl_moe():
    from .families import Qwen3MoeForCausalLM

    class Qwen3VLMoeForConditionalGeneration(Qwen3VLForConditionalGeneration, Qwen3MoeForCausalLM):
        """Qwen3-VL with the Qwen3-MoE decoder (reference ``python/sglang/srt/models/qwen3_vl_moe.py``).  Release checkpoints store the experts
        fused and transposed (``experts.gate_up_proj`` [E, H, 2I], ``experts.down_proj`` [E, I, H]); they are unpacked to per-expert tensors
        on the way into the shared MoE loader."""

        def load_weights(self, weights):
            H = self.cfg.hidden_size

            def unfuse():
                for name, w in weights:
                    if name.endswith(".mlp.experts.gate_up_proj") and w.dim() == 3:
                        w = w.transpose(1, 2) if w.shape[1] == H and w.shape[2] != H else w          # -> [E, 2I, H]
                        pre = name[: -len("gate_up_proj")]
                        for e in range(w.shape[0]):
                            g, u = w[e].chunk(2, 0)
                            yield f"{pre}{e}.gate_proj.weight", g
                            yield f"{pre}{e}.up_proj.weight", u
                    elif name.endswith(".mlp.experts.down_proj") and w.dim() == 3:
                        w = w.transpose(1, 2) if w.shape[2] == H and w.shape[1] != H else w          # -> [E, H, I]
                        pre = name[: -len("down_proj")]
                        for e in range(w.shape[0]):
                            yield f"{pre}{e}.down_proj.weight", w[e]
                    else:
                        yield name, w
            return super().load_weights(unfuse())

    return Qwen3VLMoeForConditionalGeneration
