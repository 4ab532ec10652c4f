"""Plain-PyTorch fp32 reference of every hot op.

Two jobs: (1) the numerics oracle every CUDA kernel test compares against (tests/test_kernels_gpu.py),
(2) the CPU execution path for the no-GPU test tier (reference analogue:
``python/sglang/srt/layers/attention/torch_native_backend.py`` and the ``forward_native`` methods).
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch
import torch.nn.functional as F


def rmsnorm(x: torch.Tensor, weight: torch.Tensor, eps: float) -> torch.Tensor:
    xf = x.float()
    var = xf.pow(2).mean(-1, keepdim=True)
    return (xf * torch.rsqrt(var + eps) * weight.float()).to(x.dtype)


def fused_add_rmsnorm(x: torch.Tensor, residual: torch.Tensor, weight: torch.Tensor, eps: float) -> None:
    """residual <- x + residual ; x <- rmsnorm(residual) * w   (both in place)"""
    s = x.float() + residual.float()
    residual.copy_(s.to(residual.dtype))
    sf = residual.float()
    var = sf.pow(2).mean(-1, keepdim=True)
    x.copy_((sf * torch.rsqrt(var + eps) * weight.float()).to(x.dtype))


def gemma_rmsnorm(x: torch.Tensor, weight: torch.Tensor, eps: float) -> torch.Tensor:
    xf = x.float()
    var = xf.pow(2).mean(-1, keepdim=True)
    return (xf * torch.rsqrt(var + eps) * (1.0 + weight.float())).to(x.dtype)


def layernorm(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], eps: float) -> torch.Tensor:
    return F.layer_norm(x.float(), (x.shape[-1],), weight.float(), None if bias is None else bias.float(), eps).to(x.dtype)


def silu_and_mul(x: torch.Tensor) -> torch.Tensor:
    d = x.shape[-1] // 2
    return (F.silu(x[..., :d].float()) * x[..., d:].float()).to(x.dtype)


def gelu_and_mul(x: torch.Tensor, approximate: str = "none") -> torch.Tensor:
    d = x.shape[-1] // 2
    return (F.gelu(x[..., :d].float(), approximate=approximate) * x[..., d:].float()).to(x.dtype)


def build_cos_sin_cache(head_dim: int, max_pos: int, base: float = 10000.0, rope_scaling: Optional[dict] = None,
                        rotary_dim: Optional[int] = None) -> torch.Tensor:
    """[max_pos, rotary_dim] fp32: cos | sin halves (reference rotary_embedding/base.py; llama3 scaling
    ``rotary_embedding/*`` llama3 variant)"""
    rd = rotary_dim or head_dim
    inv_freq = 1.0 / (base ** (torch.arange(0, rd, 2, dtype=torch.float32) / rd))
    if rope_scaling and rope_scaling.get("rope_type", rope_scaling.get("type")) == "llama3":
        factor = rope_scaling.get("factor", 8.0)
        lo = rope_scaling.get("low_freq_factor", 1.0)
        hi = rope_scaling.get("high_freq_factor", 4.0)
        old = rope_scaling.get("original_max_position_embeddings", 8192)
        wavelen = 2 * math.pi / inv_freq
        smooth = (old / wavelen - lo) / (hi - lo)
        scaled = torch.where(wavelen > old / lo, inv_freq / factor,
                             torch.where(wavelen < old / hi, inv_freq, (1 - smooth) * inv_freq / factor + smooth * inv_freq))
        inv_freq = scaled
    elif rope_scaling and rope_scaling.get("rope_type", rope_scaling.get("type")) == "linear":
        inv_freq = inv_freq / rope_scaling.get("factor", 1.0)
    mscale = 1.0
    kind = (rope_scaling or {}).get("rope_type", (rope_scaling or {}).get("type"))
    if kind in ("yarn", "deepseek_yarn"):
        # YaRN (reference rotary_embedding yarn / deepseek_yarn): interpolate the low frequencies by ``factor``, keep the high ones, blend
        # linearly between the two correction dims; cos / sin carry the attention temperature ("mscale")
        rs = rope_scaling
        factor = float(rs.get("factor", 1.0))
        old = rs.get("original_max_position_embeddings", max_pos)
        fast, slow = rs.get("beta_fast", 32), rs.get("beta_slow", 1)
        corr = lambda rot: (rd * math.log(old / (rot * 2 * math.pi))) / (2 * math.log(base))
        lo, hi = corr(fast), corr(slow)
        if rs.get("truncate", True):
            lo, hi = math.floor(lo), math.ceil(hi)
        lo, hi = max(lo, 0), min(hi, rd - 1)
        if lo == hi:
            hi += 0.001
        ramp = ((torch.arange(rd // 2, dtype=torch.float32) - lo) / (hi - lo)).clamp(0, 1)
        extrap_mask = 1 - ramp
        inv_freq = inv_freq / factor * (1 - extrap_mask) + inv_freq * extrap_mask
        get_ms = lambda sc, m=1.0: 1.0 if sc <= 1 else 0.1 * m * math.log(sc) + 1.0
        if rs.get("attention_factor") is not None:
            mscale = float(rs["attention_factor"])
        elif rs.get("mscale") and rs.get("mscale_all_dim"):
            mscale = get_ms(factor, rs["mscale"]) / get_ms(factor, rs["mscale_all_dim"])
        else:
            mscale = get_ms(factor)
    elif kind in ("longrope", "su"):
        # Phi-3 LongRoPE: per-dimension rescale factors (short / long table chosen by the served context length)
        rs = rope_scaling
        old = rs.get("original_max_position_embeddings", max_pos)
        ext = torch.tensor(rs["long_factor"] if max_pos > old else rs["short_factor"], dtype=torch.float32)
        inv_freq = inv_freq / ext
        sc = max_pos / old
        mscale = 1.0 if sc <= 1.0 else math.sqrt(1 + math.log(sc) / math.log(old))
    elif kind == "dynamic":
        # dynamic NTK evaluated at the served context length (static table: one base for all positions)
        rs = rope_scaling
        factor = float(rs.get("factor", 1.0))
        old = rs.get("original_max_position_embeddings", max_pos)
        if max_pos > old:
            nb = base * ((factor * max_pos / old) - (factor - 1)) ** (rd / (rd - 2))
            inv_freq = 1.0 / (nb ** (torch.arange(0, rd, 2, dtype=torch.float32) / rd))
    t = torch.arange(max_pos, dtype=torch.float32)
    freqs = torch.outer(t, inv_freq)
    return torch.cat([freqs.cos() * mscale, freqs.sin() * mscale], dim=-1)


def apply_rope(x: torch.Tensor, positions: torch.Tensor, cos_sin: torch.Tensor, is_neox: bool = True) -> torch.Tensor:
    """x [T, H, D]; rotary over the first rotary_dim = cos_sin.shape[1] dims"""
    rd = cos_sin.shape[1]
    cs = cos_sin.to(x.device)[positions.long()]
    cos, sin = cs[:, : rd // 2].unsqueeze(1).float(), cs[:, rd // 2:].unsqueeze(1).float()
    xr, xp = x[..., :rd].float(), x[..., rd:]
    if is_neox:
        x1, x2 = xr[..., : rd // 2], xr[..., rd // 2:]
        o = torch.cat([x1 * cos - x2 * sin, x2 * cos + x1 * sin], dim=-1)
    else:
        x1, x2 = xr[..., 0::2], xr[..., 1::2]
        o = torch.stack([x1 * cos - x2 * sin, x2 * cos + x1 * sin], dim=-1).flatten(-2)
    return torch.cat([o.to(x.dtype), xp], dim=-1)


def rope_kv_write(q, k, v, positions, cos_sin, k_cache, v_cache, out_cache_loc, is_neox=True,
                  k_scale: float = 1.0, v_scale: float = 1.0) -> None:
    q.copy_(apply_rope(q, positions, cos_sin, is_neox))
    k.copy_(apply_rope(k, positions, cos_sin, is_neox))
    loc = out_cache_loc.long()
    if k_cache.dtype == torch.float8_e4m3fn:
        k_cache[loc] = (k.float() / k_scale).to(k_cache.dtype)
        v_cache[loc] = (v.float() / v_scale).to(v_cache.dtype)
    else:
        k_cache[loc] = k.to(k_cache.dtype)
        v_cache[loc] = v.to(v_cache.dtype)


def paged_attention(q: torch.Tensor, k_cache: torch.Tensor, v_cache: torch.Tensor, page_table: torch.Tensor,
                    req_slots: torch.Tensor, seq_lens: torch.Tensor, q_lens: Optional[torch.Tensor], page_size: int,
                    scale: float, causal: bool = True, window: int = -1, logit_cap: float = 0.0,
                    k_scale: float = 1.0, v_scale: float = 1.0, custom_mask: Optional[torch.Tensor] = None,
                    return_lse: bool = False):
    """q [T, Hq, D] (ragged: request b owns q_lens[b] rows); caches [slots, Hkv, D].
    The KV of the current tokens is already in the cache (positions seq_len-q_len .. seq_len-1)."""
    T, Hq, D = q.shape
    Hkv = k_cache.shape[1]
    g = Hq // Hkv
    B = req_slots.shape[0]
    out = q.new_empty(T, Hq, v_cache.shape[-1])          # (the value width may differ from the key width: MLA latent rows)
    lse = torch.empty(T, Hq, dtype=torch.float32, device=q.device)
    qs = 0
    mask_off = 0
    for b in range(B):
        L = int(seq_lens[b])
        ql = 1 if q_lens is None else int(q_lens[b])
        npg = (L + page_size - 1) // page_size
        pages = page_table[int(req_slots[b]), :npg].long()
        slots = (pages[:, None] * page_size + torch.arange(page_size, device=q.device)[None, :]).reshape(-1)[:L]
        K = k_cache[slots].float() * k_scale   # [L, Hkv, D]
        V = v_cache[slots].float() * v_scale
        Q = q[qs:qs + ql].float()              # [ql, Hq, D]
        Kx = K.repeat_interleave(g, dim=1)
        Vx = V.repeat_interleave(g, dim=1)
        s = torch.einsum("qhd,khd->hqk", Q, Kx) * scale
        if logit_cap > 0:
            s = logit_cap * torch.tanh(s / logit_cap)
        qpos = torch.arange(L - ql, L, device=q.device)[:, None]
        kpos = torch.arange(L, device=q.device)[None, :]
        if custom_mask is not None:
            m = custom_mask[mask_off:mask_off + ql * L].view(ql, L).bool()
            mask_off += ql * L
            s = s.masked_fill(~m[None], float("-inf"))
        else:
            if causal:
                s = s.masked_fill((kpos > qpos)[None], float("-inf"))
            if window > 0:
                s = s.masked_fill((kpos <= qpos - window)[None], float("-inf"))
        l = torch.logsumexp(s, dim=-1)         # [Hq, ql]
        p = torch.exp(s - l[..., None])
        o = torch.einsum("hqk,khd->qhd", p, Vx)
        out[qs:qs + ql] = o.to(q.dtype)
        lse[qs:qs + ql] = l.transpose(0, 1)
        qs += ql
    return (out, lse) if return_lse else out


def merge_attn_states(o1, lse1, o2, lse2):
    """LSE-weighted merge of two partial attention results (reference merge_attn_states.cu:126)"""
    m = torch.maximum(lse1, lse2)
    w1, w2 = torch.exp(lse1 - m), torch.exp(lse2 - m)
    o = (o1.float() * w1[..., None] + o2.float() * w2[..., None]) / (w1 + w2)[..., None]
    return o.to(o1.dtype), m + torch.log(w1 + w2)


def linear(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    return F.linear(x, w, bias)


# ------------------------------------------------------------------ sampling
def top_k_top_p_min_p_probs(probs: torch.Tensor, top_k: torch.Tensor, top_p: torch.Tensor, min_p: torch.Tensor) -> torch.Tensor:
    """renormalised probabilities after top-k -> top-p -> min-p filtering (fp32)"""
    sp, idx = probs.sort(dim=-1, descending=True)
    csum = sp.cumsum(-1)
    V = probs.shape[-1]
    rank = torch.arange(V, device=probs.device)[None, :]
    keep = rank < top_k.clamp(max=V)[:, None]
    keep &= (csum - sp) < top_p[:, None]
    keep &= sp >= (sp[:, :1] * min_p[:, None])
    sp = torch.where(keep, sp, torch.zeros_like(sp))
    sp = sp / sp.sum(-1, keepdim=True)
    return torch.zeros_like(probs).scatter_(1, idx, sp)


def sample(logits: torch.Tensor, temperature: torch.Tensor, top_k: torch.Tensor, top_p: torch.Tensor,
           min_p: torch.Tensor, uniform: torch.Tensor) -> torch.Tensor:
    """inverse-CDF sampling in sorted-prob order with one uniform per row; top_k==1 rows are argmax"""
    lf = logits.float()
    greedy = top_k == 1
    probs = torch.softmax(lf / temperature.clamp(min=1e-6)[:, None], dim=-1)
    p = top_k_top_p_min_p_probs(probs, top_k, top_p, min_p)
    sp, idx = p.sort(dim=-1, descending=True)
    c = sp.cumsum(-1)
    pick = (c < uniform[:, None] * c[:, -1:]).sum(-1).clamp(max=logits.shape[-1] - 1)
    tok = idx.gather(1, pick[:, None]).squeeze(1)
    return torch.where(greedy, lf.argmax(-1), tok)


def apply_token_bitmask(logits: torch.Tensor, bitmask: torch.Tensor) -> None:
    """bitmask int32 [B, ceil(V/32)], bit=1 allowed (xgrammar convention; reference
    apply_token_bitmask_inplace_cuda.cu:157)"""
    V = logits.shape[-1]
    bits = ((bitmask[:, :, None] >> torch.arange(32, device=bitmask.device, dtype=torch.int32)[None, None, :]) & 1)
    allowed = bits.reshape(bitmask.shape[0], -1)[:, :V].bool()
    logits.masked_fill_(~allowed, float("-inf"))


def apply_penalties(logits, output_counts, prompt_mask, freq_pen, pres_pen, rep_pen) -> None:
    lf = logits.float()
    seen = (output_counts > 0) | prompt_mask
    lf = torch.where(seen & (rep_pen[:, None] != 1.0), torch.where(lf > 0, lf / rep_pen[:, None], lf * rep_pen[:, None]), lf)
    lf = lf - freq_pen[:, None] * output_counts.float() - pres_pen[:, None] * (output_counts > 0).float()
    logits.copy_(lf.to(logits.dtype))


# ------------------------------------------------------------------ MoE
def moe_topk_softmax(router_logits: torch.Tensor, top_k: int, renormalize: bool = True) -> Tuple[torch.Tensor, torch.Tensor]:
    p = torch.softmax(router_logits.float(), dim=-1)
    w, ids = p.topk(top_k, dim=-1)
    if renormalize:
        w = w / w.sum(-1, keepdim=True)
    return w, ids.to(torch.int32)


def fused_moe(x: torch.Tensor, w13: torch.Tensor, w2: torch.Tensor, topk_w: torch.Tensor, topk_ids: torch.Tensor,
              expert_offset: int = 0) -> torch.Tensor:
    """x [T,H]; w13 [E_local, 2I, H]; w2 [E_local, H, I]; ids are global expert ids; experts outside
    [expert_offset, expert_offset+E_local) contribute zero (EP)."""
    T, H = x.shape
    out = torch.zeros(T, H, dtype=torch.float32, device=x.device)
    E = w13.shape[0]
    for e in range(E):
        sel = (topk_ids == e + expert_offset)
        tok, kk = sel.nonzero(as_tuple=True)
        if tok.numel() == 0:
            continue
        h = F.linear(x[tok], w13[e])
        h = silu_and_mul(h)
        y = F.linear(h, w2[e]).float()
        out.index_add_(0, tok, y * topk_w[tok, kk].float()[:, None])
    return out.to(x.dtype)


# ------------------------------------------------------------------ quantisation
FP8_MAX = 448.0


def per_token_group_quant_fp8(x: torch.Tensor, group: int = 128, ue8m0: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
    T, K = x.shape
    xg = x.float().view(T, K // group, group)
    amax = xg.abs().amax(-1).clamp(min=1e-10)
    s = amax / FP8_MAX
    if ue8m0:
        s = torch.exp2(torch.ceil(torch.log2(s)))
    q = (xg / s[..., None]).clamp(-FP8_MAX, FP8_MAX).to(torch.float8_e4m3fn).view(T, K)
    return q, s


def block_quant_weight_fp8(w: torch.Tensor, block: int = 128, ue8m0: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
    N, K = w.shape
    nb, kb = (N + block - 1) // block, (K + block - 1) // block
    wp = torch.zeros(nb * block, kb * block, dtype=torch.float32, device=w.device)
    wp[:N, :K] = w.float()
    wb = wp.view(nb, block, kb, block)
    amax = wb.abs().amax(dim=(1, 3)).clamp(min=1e-10)
    s = amax / FP8_MAX
    if ue8m0:
        s = torch.exp2(torch.ceil(torch.log2(s)))
    q = (wb / s[:, None, :, None]).clamp(-FP8_MAX, FP8_MAX).to(torch.float8_e4m3fn).view(nb * block, kb * block)[:N, :K]
    return q.contiguous(), s


def fp8_block_gemm(xq: torch.Tensor, xs: torch.Tensor, wq: torch.Tensor, ws: torch.Tensor, block: int = 128,
                   out_dtype=torch.bfloat16) -> torch.Tensor:
    T, K = xq.shape
    N = wq.shape[0]
    x = xq.float().view(T, K // block, block) * xs[..., None]
    wsf = ws.repeat_interleave(block, 0)[:N].repeat_interleave(block, 1)[:, :K]
    w = wq.float() * wsf
    return (x.view(T, K) @ w.t()).to(out_dtype)


def tree_anc_to_mask(tree_anc: torch.Tensor, q_lens: torch.Tensor, seq_lens: torch.Tensor) -> torch.Tensor:
    """dense custom mask (flat, per request [q_len, seq_len] row-major) from per-node ancestor bits: keys below the prefix are visible to
    every node, draft key ``prefix + k`` iff bit k of the node's word is set (speculative token-tree verification)"""
    out = []
    anc = tree_anc.to(torch.int64) & 0xFFFFFFFF
    off = 0
    for ql, L in zip(q_lens.tolist(), seq_lens.tolist()):
        p = L - ql
        m = torch.ones(ql, L, dtype=torch.bool, device=tree_anc.device)
        bits = (anc[off:off + ql, None] >> torch.arange(ql, device=tree_anc.device)[None, :]) & 1
        m[:, p:] = bits.bool()
        out.append(m.reshape(-1))
        off += ql
    return torch.cat(out) if out else torch.zeros(0, dtype=torch.bool, device=tree_anc.device)


def causal_conv1d(x: torch.Tensor, w: torch.Tensor, bias, conv_state: torch.Tensor, cu_q, req_slots: torch.Tensor, seq_lens: torch.Tensor,
                  silu: bool = True) -> torch.Tensor:
    """ragged causal depthwise conv with a rolling per-request state (in place): x [T, C], w [C, k], conv_state [slots, C, k-1] -> fp32 [T, C].
    ``cu_q`` None = decode (one token per request)."""
    import torch.nn.functional as F

    T, C = x.shape
    k = w.shape[1]
    out = torch.empty(T, C, dtype=torch.float32, device=x.device)
    B = req_slots.numel()
    cu = cu_q.tolist() if cu_q is not None else list(range(B + 1))
    sl = seq_lens.tolist()
    slots = req_slots.tolist()
    for b in range(B):
        q0, n = cu[b], cu[b + 1] - cu[b]
        if n <= 0:
            continue
        s = slots[b]
        prev = torch.zeros_like(conv_state[s]) if sl[b] - n == 0 else conv_state[s]
        seq = torch.cat([prev.t(), x[q0:q0 + n]], 0)                       # [k-1+n, C]
        conv_state[s] = seq[-(k - 1):].t()
        u = F.conv1d(seq.t().unsqueeze(0).float(), w.float().unsqueeze(1), bias.float() if bias is not None else None, groups=C)[0].t()
        out[q0:q0 + n] = F.silu(u) if silu else u
    return out


def selective_state_update(state: torch.Tensor, req_slots: torch.Tensor, x: torch.Tensor, dt: torch.Tensor, A: torch.Tensor, Bm: torch.Tensor,
                           Cm: torch.Tensor, D) -> torch.Tensor:
    """state fp32 [slots, H, P, N] (in place); x [B, H*P], dt [B, H], A [H], Bm / Cm [B, G*N], D [H] or None -> y fp32 [B, H, P]"""
    Bn = req_slots.numel()
    H, P, N = state.shape[1:]
    G = Bm.shape[1] // N
    rep = H // G
    sl = req_slots.long()
    st = state[sl]
    xh = x.float().view(Bn, H, P)
    dA = torch.exp(dt.float() * A.float()[None, :])
    Bh = Bm.float().reshape(Bn, G, N).repeat_interleave(rep, 1)
    Ch = Cm.float().reshape(Bn, G, N).repeat_interleave(rep, 1)
    st = st * dA[..., None, None] + (dt.float()[..., None] * xh)[..., None] * Bh[:, :, None, :]
    state[sl] = st
    y = torch.einsum("bhpn,bhn->bhp", st, Ch)
    if D is not None:
        y = y + D.float()[None, :, None] * xh
    return y
