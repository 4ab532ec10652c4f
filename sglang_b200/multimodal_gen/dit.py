"""Denoiser, text encoder and latent decoder of the diffusion runtime.

The denoiser is an MMDiT-style transformer (joint attention over text + latent-patch tokens, adaLN-zero modulation from the
timestep and the pooled text embedding), the family the reference serves in ``multimodal_gen/runtime/models/dits/*``.
Attention goes through :func:`usp_attention`, so the same module runs on one GPU or with its token rows scattered over a
Ulysses x ring grid — linears, norms and modulation are row-local, only attention communicates.
"""
from __future__ import annotations

import math
from dataclasses import asdict, dataclass
from typing import Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from .usp import UspGroup, usp_attention


@dataclass
class DiTConfig:
    hidden_size: int = 64
    num_heads: int = 4
    num_layers: int = 2
    mlp_ratio: float = 4.0
    in_channels: int = 4
    patch_size: int = 2
    text_len: int = 16
    text_vocab: int = 258            # utf-8 bytes + bos + pad
    text_layers: int = 1
    vae_scale: int = 8               # pixel / latent
    vae_channels: int = 32
    scaling_factor: float = 1.0
    prediction: str = "flow"         # "flow" (velocity, rectified flow) or "epsilon" (DDIM)
    qk_norm: bool = True

    def to_dict(self):
        return asdict(self)


def timestep_embedding(t: torch.Tensor, dim: int, max_period: float = 10000.0) -> torch.Tensor:
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(half, dtype=torch.float32, device=t.device) / half)
    args = t.float()[:, None] * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


def sincos_nd(grid: Tuple[int, ...], dim: int, device) -> torch.Tensor:
    """N-d sin/cos position table for a (frames, h, w) or (h, w) patch grid -> [prod(grid), dim]."""
    n = len(grid)
    per = (dim // n) // 2 * 2
    coords = torch.meshgrid(*[torch.arange(g, dtype=torch.float32, device=device) for g in grid], indexing="ij")
    out = []
    for c in coords:
        omega = 1.0 / (10000 ** (torch.arange(per // 2, dtype=torch.float32, device=device) / max(per // 2, 1)))
        a = c.reshape(-1)[:, None] * omega[None]
        out += [torch.sin(a), torch.cos(a)]
    pe = torch.cat(out, dim=-1)
    return F.pad(pe, (0, dim - pe.shape[-1]))


class _Attn(nn.Module):
    def __init__(self, h: int, heads: int, qk_norm: bool):
        super().__init__()
        self.heads, self.hd = heads, h // heads
        self.qkv = nn.Linear(h, 3 * h)
        self.proj = nn.Linear(h, h)
        self.qn = nn.RMSNorm(self.hd, eps=1e-6) if qk_norm else nn.Identity()
        self.kn = nn.RMSNorm(self.hd, eps=1e-6) if qk_norm else nn.Identity()

    def forward(self, x: torch.Tensor, sp: Optional[UspGroup]) -> torch.Tensor:           # x [B, S_local, h]
        B, S, _ = x.shape
        q, k, v = self.qkv(x).view(B, S, 3, self.heads, self.hd).unbind(2)
        q, k = self.qn(q), self.kn(k)
        out = torch.stack([usp_attention(q[b], k[b], v[b], sp) for b in range(B)])
        return self.proj(out.reshape(B, S, -1))


class DiTBlock(nn.Module):
    def __init__(self, c: DiTConfig):
        super().__init__()
        h = c.hidden_size
        self.n1 = nn.LayerNorm(h, elementwise_affine=False, eps=1e-6)
        self.attn = _Attn(h, c.num_heads, c.qk_norm)
        self.n2 = nn.LayerNorm(h, elementwise_affine=False, eps=1e-6)
        self.fc1 = nn.Linear(h, int(h * c.mlp_ratio))
        self.fc2 = nn.Linear(int(h * c.mlp_ratio), h)
        self.mod = nn.Linear(h, 6 * h)

    def forward(self, x, cond, sp):
        sh1, sc1, g1, sh2, sc2, g2 = self.mod(F.silu(cond)).unsqueeze(1).chunk(6, dim=-1)
        x = x + g1 * self.attn(self.n1(x) * (1 + sc1) + sh1, sp)
        return x + g2 * self.fc2(F.gelu(self.fc1(self.n2(x) * (1 + sc2) + sh2), approximate="tanh"))


class DiT(nn.Module):
    def __init__(self, c: DiTConfig):
        super().__init__()
        self.cfg = c
        h, p = c.hidden_size, c.patch_size
        self.x_in = nn.Linear(c.in_channels * p * p, h)
        self.txt_in = nn.Linear(h, h)
        self.t_mlp = nn.Sequential(nn.Linear(256, h), nn.SiLU(), nn.Linear(h, h))
        self.pool_mlp = nn.Sequential(nn.Linear(h, h), nn.SiLU(), nn.Linear(h, h))
        self.blocks = nn.ModuleList(DiTBlock(c) for _ in range(c.num_layers))
        self.n_out = nn.LayerNorm(h, elementwise_affine=False, eps=1e-6)
        self.mod_out = nn.Linear(h, 2 * h)
        self.x_out = nn.Linear(h, c.in_channels * p * p)

    # -- token layout ---------------------------------------------------------------------------------------------------
    def patchify(self, z: torch.Tensor) -> Tuple[torch.Tensor, Tuple[int, int, int]]:
        """z [B, C, F, H, W] -> tokens [B, F*h*w, C*p*p]"""
        B, C, Fr, H, W = z.shape
        p = self.cfg.patch_size
        t = z.reshape(B, C, Fr, H // p, p, W // p, p).permute(0, 2, 3, 5, 1, 4, 6)
        return t.reshape(B, Fr * (H // p) * (W // p), C * p * p), (Fr, H // p, W // p)

    def unpatchify(self, t: torch.Tensor, grid) -> torch.Tensor:
        Fr, gh, gw = grid
        p, C = self.cfg.patch_size, self.cfg.in_channels
        B = t.shape[0]
        return t.reshape(B, Fr, gh, gw, C, p, p).permute(0, 4, 1, 2, 5, 3, 6).reshape(B, C, Fr, gh * p, gw * p)

    def forward(self, z, t, txt, pooled, sp: Optional[UspGroup] = None) -> torch.Tensor:
        """z [B,C,F,H,W] noisy latents (replicated on every SP rank), t [B] in [0,1000], txt [B,L,h], pooled [B,h] -> model output."""
        tok, grid = self.patchify(z)
        x = self.x_in(tok) + sincos_nd(grid if grid[0] > 1 else grid[1:], self.cfg.hidden_size, z.device).to(tok.dtype)
        seq = torch.cat([self.txt_in(txt), x], dim=1)                                     # joint sequence
        S = seq.shape[1]
        if sp is not None and sp.sp_size > 1:
            assert S % sp.sp_size == 0, f"text_len + patches ({S}) must divide the SP size {sp.sp_size}"
            rows = S // sp.sp_size
            seq = seq[:, sp.sp_rank * rows:(sp.sp_rank + 1) * rows]
        cond = self.t_mlp(timestep_embedding(t, 256).to(seq.dtype)) + self.pool_mlp(pooled)
        for blk in self.blocks:
            seq = blk(seq, cond, sp)
        sh, sc = self.mod_out(F.silu(cond)).unsqueeze(1).chunk(2, dim=-1)
        out = self.x_out(self.n_out(seq) * (1 + sc) + sh)
        if sp is not None and sp.sp_size > 1:
            out = _all_gather_rows(out, sp)
        return self.unpatchify(out[:, self.cfg.text_len:], grid)


def _all_gather_rows(x: torch.Tensor, sp: UspGroup) -> torch.Tensor:
    """Gather row blocks in sp_rank order.  Two hops (ulysses then ring) so no extra process group is needed."""
    import torch.distributed as dist

    x = x.contiguous()
    if sp.ulysses_size > 1:
        parts = [torch.empty_like(x) for _ in range(sp.ulysses_size)]
        dist.all_gather(parts, x, group=sp.ulysses)
        x = torch.cat(parts, dim=1)
    if sp.ring_size > 1:
        parts = [torch.empty_like(x) for _ in range(sp.ring_size)]
        dist.all_gather(parts, x, group=sp.ring)
        x = torch.cat(parts, dim=1)
    return x


class TextEncoder(nn.Module):
    """Byte-level bidirectional transformer: prompt -> ([L, h] token states, [h] pooled)."""

    def __init__(self, c: DiTConfig):
        super().__init__()
        self.cfg = c
        h = c.hidden_size
        self.emb = nn.Embedding(c.text_vocab, h)
        self.pos = nn.Embedding(c.text_len, h)
        self.layers = nn.ModuleList(nn.TransformerEncoderLayer(h, c.num_heads, int(h * c.mlp_ratio), dropout=0.0, batch_first=True,
                                                               norm_first=True, activation="gelu") for _ in range(c.text_layers))
        self.norm = nn.LayerNorm(h)

    def tokenize(self, prompts) -> torch.Tensor:
        L, pad = self.cfg.text_len, 257
        rows = [([256] + list(p.encode("utf-8")))[:L] for p in prompts]
        return torch.tensor([r + [pad] * (L - len(r)) for r in rows], dtype=torch.long)

    def forward(self, ids: torch.Tensor):
        x = self.emb(ids) + self.pos(torch.arange(ids.shape[1], device=ids.device))
        for l in self.layers:
            x = l(x)
        x = self.norm(x)
        m = (ids != 257).unsqueeze(-1).to(x.dtype)
        return x, (x * m).sum(1) / m.sum(1).clamp(min=1)


class _Res(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.n1, self.c1 = nn.GroupNorm(min(8, cin), cin), nn.Conv2d(cin, cout, 3, padding=1)
        self.n2, self.c2 = nn.GroupNorm(min(8, cout), cout), nn.Conv2d(cout, cout, 3, padding=1)
        self.skip = nn.Conv2d(cin, cout, 1) if cin != cout else nn.Identity()

    def forward(self, x):
        h = self.c1(F.silu(self.n1(x)))
        return self.skip(x) + self.c2(F.silu(self.n2(h)))


class LatentDecoder(nn.Module):
    """AutoencoderKL-style decoder: [B, C, h, w] latents -> [B, 3, h*s, w*s] in [-1, 1]; frames are decoded as a batch."""

    def __init__(self, c: DiTConfig):
        super().__init__()
        ch, ups = c.vae_channels, int(math.log2(c.vae_scale))
        self.scaling = c.scaling_factor
        self.conv_in = nn.Conv2d(c.in_channels, ch, 3, padding=1)
        self.mid = _Res(ch, ch)
        self.ups = nn.ModuleList()
        for i in range(ups):
            self.ups.append(nn.ModuleList([_Res(ch, max(ch // 2, 8)), nn.Conv2d(max(ch // 2, 8), max(ch // 2, 8), 3, padding=1)]))
            ch = max(ch // 2, 8)
        self.n_out, self.conv_out = nn.GroupNorm(min(8, ch), ch), nn.Conv2d(ch, 3, 3, padding=1)

    def forward(self, z: torch.Tensor, tile: int = 0) -> torch.Tensor:
        if tile and (z.shape[-1] > tile or z.shape[-2] > tile):
            return self._tiled(z, tile)
        x = self.mid(self.conv_in(z / self.scaling))
        for res, conv in self.ups:
            x = conv(F.interpolate(res(x), scale_factor=2.0, mode="nearest"))
        return torch.tanh(self.conv_out(F.silu(self.n_out(x))))

    def _tiled(self, z, tile, overlap: int = 2):
        """Tiled decode with linear blending in the overlap, so peak activation memory is bounded by the tile."""
        s = 2 ** len(self.ups)
        B, _, H, W = z.shape
        out = torch.zeros(B, 3, H * s, W * s, dtype=z.dtype, device=z.device)
        wsum = torch.zeros(1, 1, H * s, W * s, dtype=z.dtype, device=z.device)
        step = tile - overlap
        for y in range(0, H, step):
            for x in range(0, W, step):
                y1, x1 = min(y + tile, H), min(x + tile, W)
                d = self.forward(z[:, :, y:y1, x:x1])
                th, tw = d.shape[-2:]
                wy = torch.minimum(torch.arange(1, th + 1), torch.arange(th, 0, -1)).clamp(max=overlap * s).to(z)
                wx = torch.minimum(torch.arange(1, tw + 1), torch.arange(tw, 0, -1)).clamp(max=overlap * s).to(z)
                w = (wy[:, None] * wx[None])[None, None]
                out[:, :, y * s:y1 * s, x * s:x1 * s] += d * w
                wsum[:, :, y * s:y1 * s, x * s:x1 * s] += w
        return out / wsum
