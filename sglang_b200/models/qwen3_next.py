"""Qwen3-Next: hybrid of *gated DeltaNet* linear-attention layers and gated softmax-attention layers, zero-centred RMSNorm, MoE with a
sigmoid-gated shared expert.

Parity: reference ``python/sglang/srt/models/qwen3_next.py`` + ``layers/attention/fla/*`` (chunked gated delta rule) + the hybrid-linear attention
backend / ``MambaPool`` state handling; checked against the Hugging Face implementation (``tests/test_hf_parity.py``).

Gated delta rule (per value head; state ``S`` is ``[Dk, Dv]``)::

    S <- exp(g_t) S;   S <- S + k_t (beta_t (v_t - S^T k_t))^T;   o_t = S^T q_t

Prefill uses the chunked WY form: inside a chunk the ``beta``-weighted rank-one corrections couple through a unit lower-triangular system
``(I - A) T = I`` with ``A[t, s] = -beta_t k_t.k_s exp(G_t - G_s)`` (``G`` = running sum of ``g``), solved for all chunks and heads at once with a
batched triangular solve; what is left is GEMM-shaped (``T V``, ``T K``, ``Q K^T``, state updates) plus a short scan over chunks — the shape a
tcgen05 kernel wants.  Decode is the three-line recurrence above.  Conv and recurrent state live in the ``MambaPool`` by request slot.
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Tuple

import torch
import torch.nn.functional as F
from torch import nn

from .. import ops
from ..configs.model_config import ModelConfig
from ..layers.common import RadixAttention, RMSNorm, RotaryEmbedding
from ..layers.linear import ParallelLMHead, QKVParallelLinear, RowParallelLinear, VocabParallelEmbedding
from ..model_executor.forward_batch import ForwardBatch
from ..parallel.parallel_state import get_pp_group, get_tp_group
from .llama import _DT, LlamaMLP
from .mamba import MambaPool, NemotronHForCausalLM


def _l2(x: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    return x * torch.rsqrt((x * x).sum(-1, keepdim=True) + eps)


def gated_delta_step(q, k, v, g, beta, state):
    """one token per row: q, k [R, H, Dk]; v [R, H, Dv]; g, beta [R, H]; state [R, H, Dk, Dv] (updated in place) -> o [R, H, Dv] fp32"""
    q, k = _l2(q.float()) * (q.shape[-1] ** -0.5), _l2(k.float())
    state.mul_(g.float().exp()[..., None, None])
    mem = torch.einsum("rhkv,rhk->rhv", state, k)
    delta = (v.float() - mem) * beta.float()[..., None]
    state.add_(k[..., :, None] * delta[..., None, :])
    return torch.einsum("rhkv,rhk->rhv", state, q)


def gated_delta_chunked(q, k, v, g, beta, state, chunk: int = 64, snap_at: Optional[int] = None):
    """one sequence: q, k [T, H, Dk]; v [T, H, Dv]; g, beta [T, H]; state [H, Dk, Dv] fp32 -> (o [T, H, Dv] fp32, final state)"""
    T, H, Dk = q.shape
    Dv = v.shape[-1]
    pad = (-T) % chunk
    q, k = _l2(q.float()) * (Dk ** -0.5), _l2(k.float())
    v, g, beta = v.float(), g.float(), beta.float()
    if pad:
        z = lambda t: F.pad(t, (0, 0) * (t.dim() - 1) + (0, pad))
        q, k, v, g, beta = z(q), z(k), z(v), z(g), z(beta)             # g = 0, beta = 0 on the padding: state untouched
    c = (T + pad) // chunk
    # [H, c, Q, ...]
    R = lambda t: t.view(c, chunk, H, -1).permute(2, 0, 1, 3)
    qc, kc, vc = R(q), R(k), R(v)
    gc = g.view(c, chunk, H).permute(2, 0, 1)
    bc = beta.view(c, chunk, H).permute(2, 0, 1)
    G = torch.cumsum(gc, -1)                                            # [H, c, Q]
    tri = torch.tril(torch.ones(chunk, chunk, dtype=torch.bool, device=q.device))
    D = torch.exp((G[..., :, None] - G[..., None, :]).masked_fill(~tri, float("-inf")))        # decay t <- s, zero above the diagonal
    kb, vb = kc * bc[..., None], vc * bc[..., None]
    A = -(kb @ kc.transpose(-1, -2)) * D
    A = A.masked_fill(torch.triu(torch.ones(chunk, chunk, dtype=torch.bool, device=q.device)), 0.0)      # strictly lower
    eye = torch.eye(chunk, device=q.device, dtype=torch.float32)
    Tm = torch.linalg.solve_triangular(eye - A, eye.expand_as(A), upper=False)                 # (I - A)^-1
    u = Tm @ vb                                                          # pseudo-values before the carried state is subtracted
    w = Tm @ (kb * G.exp()[..., None])                                   # what the carried state contributes to them
    qk = (qc @ kc.transpose(-1, -2)) * D                                 # causal (diagonal included)
    S = state.float().clone()
    snap = None
    outs = []
    for i in range(c):
        v_new = u[:, i] - w[:, i] @ S                                    # [H, Q, Dv]
        outs.append((qc[:, i] * G[:, i].exp()[..., None]) @ S + qk[:, i] @ v_new)
        last = G[:, i, -1]
        S = S * last.exp()[:, None, None] + (kc[:, i] * (last[:, None] - G[:, i]).exp()[..., None]).transpose(-1, -2) @ v_new
        if snap_at is not None and (i + 1) * chunk == snap_at:
            snap = S.clone()               # prefix-cache snapshot at a chunk boundary (mem_cache/mamba_radix.py)
    o = torch.stack(outs, 1).permute(1, 2, 0, 3).reshape(c * chunk, H, Dv)[:T]
    return (o, S) if snap_at is None else (o, S, snap)


class GatedDeltaNet(nn.Module):
    def __init__(self, cfg: ModelConfig, layer_idx: int, dtype, prefix: str):
        super().__init__()
        tp = get_tp_group()
        self.tp = tp
        W = tp.world_size
        la = cfg.linear_attn
        self.Hk_t, self.Hv_t, self.Dk, self.Dv, self.k = la["num_k_heads"], la["num_v_heads"], la["k_head_dim"], la["v_head_dim"], la["conv_kernel"]
        assert self.Hk_t % W == 0
        self.Hk, self.Hv = self.Hk_t // W, self.Hv_t // W
        self.r = self.Hv_t // self.Hk_t
        self.layer_idx = layer_idx
        self.key_dim, self.value_dim = self.Hk * self.Dk, self.Hv * self.Dv
        self.conv_dim = 2 * self.key_dim + self.value_dim
        H = cfg.hidden_size
        self.in_proj_qkvz = nn.Linear(H, 2 * self.key_dim + 2 * self.value_dim, bias=False, dtype=dtype)       # rows grouped per k-head
        self.in_proj_ba = nn.Linear(H, 2 * self.Hv, bias=False, dtype=dtype)
        self.conv_weight = nn.Parameter(torch.zeros(self.conv_dim, self.k, dtype=dtype), requires_grad=False)
        self.dt_bias = nn.Parameter(torch.ones(self.Hv, dtype=torch.float32), requires_grad=False)
        self.A_log = nn.Parameter(torch.zeros(self.Hv, dtype=torch.float32), requires_grad=False)
        self.norm_weight = nn.Parameter(torch.ones(self.Dv, dtype=dtype), requires_grad=False)
        self.eps = cfg.rms_norm_eps
        self.out_proj = RowParallelLinear(self.Hv_t * self.Dv, H, False, True, dtype, None, prefix + ".out_proj")
        for p in list(self.in_proj_qkvz.parameters()) + list(self.in_proj_ba.parameters()):
            p.requires_grad_(False)
        self.H, self.P, self.N = self.Hv, self.Dk, self.Dv          # MambaPool geometry: [slots, Hv, Dk, Dv]

    # ---- sharded loading
    def load(self, leaf: str, w: torch.Tensor, params, full: str) -> bool:
        r, W = self.tp.rank, self.tp.world_size
        blk = lambda t, n: t.narrow(0, r * n, n)
        if leaf == "in_proj_qkvz.weight":
            per = 2 * self.Dk + 2 * self.Dv * self.r
            params[full].data.copy_(blk(w, self.Hk * per))
        elif leaf == "in_proj_ba.weight":
            params[full].data.copy_(blk(w, self.Hk * 2 * self.r))
        elif leaf == "conv1d.weight":
            w = w.reshape(w.shape[0], -1)
            kd, vd = self.Hk_t * self.Dk, self.Hv_t * self.Dv
            q, k, v = w[:kd], w[kd:2 * kd], w[2 * kd:2 * kd + vd]
            self.conv_weight.data.copy_(torch.cat([blk(q, self.key_dim), blk(k, self.key_dim), blk(v, self.value_dim)], 0))
        elif leaf in ("dt_bias", "A_log"):
            getattr(self, leaf).data.copy_(blk(w, self.Hv).float())
        elif leaf == "norm.weight":
            self.norm_weight.data.copy_(w)
        elif leaf == "out_proj.weight":
            self.out_proj.weight_loader(params[full], w)
        else:
            return False
        return True

    def forward(self, h: torch.Tensor, fb: ForwardBatch) -> torch.Tensor:
        pool: MambaPool = fb.kv_pool.mamba
        conv_s, rec_s = pool.conv[self.layer_idx], pool.ssm[self.layer_idx]
        T = h.shape[0]
        per = 2 * self.Dk + 2 * self.Dv * self.r
        qkvz = self.in_proj_qkvz(h).view(T, self.Hk, per)
        q, k, v, z = qkvz.split([self.Dk, self.Dk, self.r * self.Dv, self.r * self.Dv], dim=-1)
        ba = self.in_proj_ba(h).view(T, self.Hk, 2 * self.r)
        b, a = ba[..., : self.r].reshape(T, self.Hv), ba[..., self.r:].reshape(T, self.Hv)
        mixed = torch.cat([q.reshape(T, -1), k.reshape(T, -1), v.reshape(T, -1)], -1)           # [T, conv_dim]
        z = z.reshape(T, self.Hv, self.Dv)
        beta = torch.sigmoid(b.float())
        g = -self.A_log.exp() * F.softplus(a.float() + self.dt_bias)
        slots = fb.req_slots.long()
        decode = fb.cu_q is None
        # causal conv + SiLU over the ragged batch, rolling state in place (csrc/kernels/mamba.cu)
        u = ops.causal_conv1d(mixed, self.conv_weight, None, conv_s, fb.cu_q, fb.req_slots, fb.seq_lens, fb.max_q_len, silu=True)
        o = ops.gated_delta_update(rec_s, fb.req_slots, u, g, beta, self.Hk) if decode else None      # one pass over the state (native kernel)
        qh, kh, vh = u.split([self.key_dim, self.key_dim, self.value_dim], dim=-1)
        qh = qh.view(T, self.Hk, self.Dk).repeat_interleave(self.r, 1)
        kh = kh.view(T, self.Hk, self.Dk).repeat_interleave(self.r, 1)
        vh = vh.view(T, self.Hv, self.Dv)
        if o is not None:
            pass
        elif decode:
            st = rec_s[slots]
            o = gated_delta_step(qh, kh, vh, g, beta, st)
            rec_s[slots] = st
        else:
            os_, off = [], 0
            for i in range(len(fb.q_lens_cpu)):
                n, s = int(fb.q_lens_cpu[i]), int(slots[i])
                fresh = int(fb.seq_lens_cpu[i]) - n == 0
                st0 = torch.zeros_like(rec_s[s]) if fresh else rec_s[s]
                snap = (fb.mamba_snap or {}).get(i)
                if snap is None:
                    oi, stN = gated_delta_chunked(qh[off:off + n], kh[off:off + n], vh[off:off + n], g[off:off + n], beta[off:off + n], st0)
                else:
                    rel, row = snap
                    oi, stN, sn = gated_delta_chunked(qh[off:off + n], kh[off:off + n], vh[off:off + n], g[off:off + n], beta[off:off + n], st0, snap_at=rel)
                    if sn is not None:
                        rec_s[row] = sn
                        conv_s[row] = mixed[off + rel - (self.k - 1): off + rel].t().to(conv_s.dtype)
                rec_s[s] = stN
                os_.append(oi)
                off += n
            o = torch.cat(os_, 0) if os_ else vh.new_zeros(0, self.Hv, self.Dv).float()
        # gated RMSNorm per value head: norm first, then * silu(z)
        o = o * torch.rsqrt(o.pow(2).mean(-1, keepdim=True) + self.eps)
        o = (o * self.norm_weight.float()).to(h.dtype).float() * F.silu(z.float())
        return self.out_proj(o.reshape(T, self.value_dim).to(h.dtype))


class GatedAttention(nn.Module):
    """softmax attention whose q projection also emits a per-channel output gate; per-head zero-centred q/k norm, partial rotary.
    The checkpoint's q_proj rows are [head][q (hd) | gate (hd)]: to the fused QKV linear that is simply ``2 * heads`` heads of width ``hd``
    (pairs stay together under TP because heads are sharded contiguously)."""

    def __init__(self, cfg: ModelConfig, layer_id: int, dtype, rope, prefix: str):
        super().__init__()
        hd = cfg.head_dim
        self.qkv_proj = QKVParallelLinear(cfg.hidden_size, hd, 2 * cfg.num_attention_heads, cfg.num_key_value_heads, cfg.attention_bias, dtype, None,
                                          prefix + ".qkv_proj")
        self.nh, self.nkv, self.hd = self.qkv_proj.num_heads // 2, self.qkv_proj.num_kv_heads, hd
        self.o_proj = RowParallelLinear(cfg.num_attention_heads * hd, cfg.hidden_size, cfg.attention_bias, True, dtype, None, prefix + ".o_proj")
        self.q_norm, self.k_norm = RMSNorm(hd, cfg.rms_norm_eps, dtype, True), RMSNorm(hd, cfg.rms_norm_eps, dtype, True)
        self.attn = RadixAttention(self.nh, hd, self.nkv, layer_id, rope=rope)

    def load(self, leaf: str, w: torch.Tensor) -> bool:
        proj, kind = leaf.split(".")[0], leaf.split(".")[1]
        if proj not in ("q_proj", "k_proj", "v_proj"):
            return False
        if kind == "weight":
            self.qkv_proj.weight_loader(self.qkv_proj.weight, w, proj[0])
        else:
            from .llama import _load_stacked_bias

            _load_stacked_bias(self.qkv_proj, self.qkv_proj.bias, w, proj[0])
        return True

    def forward(self, x: torch.Tensor, fb: ForwardBatch):
        T = x.shape[0]
        q2, k, v = self.qkv_proj(x).split([self.qkv_proj.q_size, self.qkv_proj.kv_size, self.qkv_proj.kv_size], dim=-1)
        q2 = q2.view(T, self.nh, 2 * self.hd)
        q, gate = q2[..., : self.hd], q2[..., self.hd:].reshape(T, self.nh * self.hd)
        q = self.q_norm(q.reshape(T * self.nh, self.hd)).view(T, self.nh * self.hd)
        k = self.k_norm(k.reshape(T * self.nkv, self.hd)).view(T, self.nkv * self.hd)
        o = self.attn(q.contiguous(), k.contiguous(), v.contiguous(), fb)
        return self.o_proj(o * torch.sigmoid(gate.float()).to(o.dtype))


class _Block(nn.Module):
    def __init__(self, cfg: ModelConfig, kind: str, idx: int, gid: int, dtype, rope):
        super().__init__()
        from .families import QwenSparseMoeBlock

        self.kind = "mamba" if kind == "linear_attention" else "attention"
        pre = f"model.layers.{gid}"
        self.input_layernorm = RMSNorm(cfg.hidden_size, cfg.rms_norm_eps, dtype, True)
        self.post_attention_layernorm = RMSNorm(cfg.hidden_size, cfg.rms_norm_eps, dtype, True)
        if self.kind == "mamba":
            self.linear_attn = GatedDeltaNet(cfg, idx, dtype, pre + ".linear_attn")
        else:
            self.self_attn = GatedAttention(cfg, idx, dtype, rope, pre + ".self_attn")
        sparse = cfg.num_local_experts > 0 and gid not in (cfg.mlp_only_layers or []) and (gid + 1) % max(1, cfg.decoder_sparse_step) == 0
        self.mlp = QwenSparseMoeBlock(cfg, gid, dtype, None, pre + ".mlp") if sparse else LlamaMLP(cfg, dtype, None, pre + ".mlp")

    def forward(self, x, fb):
        h = self.input_layernorm(x)
        y = self.linear_attn(h, fb) if self.kind == "mamba" else self.self_attn(h, fb)
        x = x + (y.resolve() if hasattr(y, "resolve") else y)
        m = self.mlp(self.post_attention_layernorm(x))
        return x + (m.resolve() if hasattr(m, "resolve") else m)


class Qwen3NextForCausalLM(NemotronHForCausalLM):
    def __init__(self, cfg: ModelConfig, quant_name: Optional[str] = None, device="cpu"):
        nn.Module.__init__(self)
        assert get_pp_group().world_size == 1
        self.cfg = cfg
        dtype = _DT[cfg.dtype]
        bb = nn.Module()
        bb.embed_tokens = VocabParallelEmbedding(cfg.vocab_size, cfg.hidden_size, dtype)
        rope = RotaryEmbedding(cfg.head_dim, cfg.max_position_embeddings, cfg.rope_theta, cfg.rope_scaling, cfg.rope_is_neox, cfg.partial_rotary_factor,
                               device)
        blocks, n_l, n_a = [], 0, 0
        for i, kind in enumerate(cfg.layer_types):
            if kind == "linear_attention":
                blocks.append(_Block(cfg, kind, n_l, i, dtype, rope))
                n_l += 1
            else:
                blocks.append(_Block(cfg, kind, n_a, i, dtype, rope))
                n_a += 1
        bb.all_layers = nn.ModuleList(blocks)
        bb.num_mamba = n_l
        bb.layers = [b for b in blocks if b.kind == "attention"]
        bb.norm = RMSNorm(cfg.hidden_size, cfg.rms_norm_eps, dtype, True)
        self.model = bb
        self.lm_head = None if cfg.tie_word_embeddings else ParallelLMHead(cfg.vocab_size, cfg.hidden_size, dtype)

    def forward(self, fb: ForwardBatch, inputs_embeds=None):
        m = self.model
        h = inputs_embeds if inputs_embeds is not None else (m.embed_tokens(fb.next_token_buf, fb.req_slots) if fb.input_ids is None
                                                             else m.embed_tokens(fb.input_ids))
        for blk in m.all_layers:
            h = blk(h, fb)
        return m.norm(h)

    def make_mamba_pool(self, slots: int, device) -> Optional[MambaPool]:
        la = next((b.linear_attn for b in self.model.all_layers if b.kind == "mamba"), None)
        return None if la is None else MambaPool(self.model.num_mamba, slots, la.conv_dim, la.k, la.Hv, la.Dk, la.Dv, _DT[self.cfg.dtype], device)

    def load_weights(self, weights: Iterable[Tuple[str, torch.Tensor]]) -> List[str]:
        params = dict(self.named_parameters())
        mods = dict(self.named_modules())
        loaded = []
        for name, w in weights:
            n = name.replace("model.layers.", "model.all_layers.")
            if n == "lm_head.weight" and self.lm_head is None:
                continue
            if ".linear_attn." in n:
                mn = n[: n.index(".linear_attn") + 12]
                if mods[mn].load(n[len(mn) + 1:], w, params, n):
                    loaded.append(name)
                continue
            if ".self_attn." in n and n.split(".self_attn.")[1].split(".")[0] in ("q_proj", "k_proj", "v_proj"):
                mn = n[: n.index(".self_attn") + 10]
                if mods[mn].load(n[len(mn) + 1:], w):
                    loaded.append(name)
                continue
            if ".mlp.experts." in n:
                ex = mods.get(n[: n.index(".experts") + 8])
                tail = n.split(".experts.")[1]
                if ex is None:
                    continue
                if tail.split(".")[0].isdigit():
                    eid, proj = int(tail.split(".")[0]), tail.split(".")[1]
                    ex.weight_loader(ex.w2 if proj == "down_proj" else ex.w13, w, eid, {"gate_proj": "w1", "up_proj": "w3", "down_proj": "w2"}[proj])
                elif tail.startswith("gate_up_proj"):
                    for e in range(w.shape[0]):
                        g_, u_ = w[e].chunk(2, 0)
                        ex.weight_loader(ex.w13, g_, e, "w1")
                        ex.weight_loader(ex.w13, u_, e, "w3")
                elif tail.startswith("down_proj"):
                    for e in range(w.shape[0]):
                        ex.weight_loader(ex.w2, w[e], e, "w2")
                loaded.append(name)
                continue
            target, shard = n, None
            for merged, orig, sid in ((".gate_up_proj", ".gate_proj", 0), (".gate_up_proj", ".up_proj", 1)):
                if orig + "." in n:
                    target, shard = n.replace(orig, merged), sid
                    break
            if target not in params:
                continue
            mod = mods[target.rsplit(".", 1)[0]]
            if hasattr(mod, "weight_loader"):
                mod.weight_loader(params[target], w, shard) if shard is not None else mod.weight_loader(params[target], w)
            else:
                params[target].data.copy_(w)
            loaded.append(name)
        return loaded


class Qwen3_5ForCausalLM(Qwen3NextForCausalLM):
    """Qwen3.5 text decoder (reference ``python/sglang/srt/models/qwen3_5_text.py``): the Qwen3-Next hybrid with a dense (or MoE) MLP.  Its
    checkpoints store the DeltaNet input projections separately and flat (``in_proj_qkv`` = [q | k | v], ``in_proj_z``, ``in_proj_b``,
    ``in_proj_a``); they are regrouped per key head into this module's ``in_proj_qkvz`` / ``in_proj_ba`` layout while loading, so the forward
    path (and its TP sharding by key head) is shared.  Text positions make the interleaved M-RoPE identical to plain rotary embedding; the
    multimodal checkpoints' ``model.language_model.*`` names load as the text tower."""

    def load_weights(self, weights: Iterable[Tuple[str, torch.Tensor]]) -> List[str]:
        la = self.cfg.linear_attn
        Hk, Hv, Dk, Dv = la["num_k_heads"], la["num_v_heads"], la["k_head_dim"], la["v_head_dim"]
        r = Hv // Hk
        pend: dict = {}

        def regroup():
            for name, w in weights:
                if name.startswith("model.language_model."):
                    name = "model." + name[len("model.language_model."):]
                elif name.startswith(("model.visual.", "visual.", "mtp.")):
                    continue
                leaf = name.rsplit(".linear_attn.", 1)[-1] if ".linear_attn." in name else ""
                if leaf not in ("in_proj_qkv.weight", "in_proj_z.weight", "in_proj_b.weight", "in_proj_a.weight"):
                    yield name, w
                    continue
                pre = name[: name.index(".linear_attn.") + len(".linear_attn.")]
                d = pend.setdefault(pre, {})
                d[leaf.split(".")[0]] = w
                if "in_proj_qkv" in d and "in_proj_z" in d:
                    qkv, z = d.pop("in_proj_qkv"), d.pop("in_proj_z")
                    Hin = qkv.shape[1]
                    q, k, v = qkv[: Hk * Dk].view(Hk, Dk, Hin), qkv[Hk * Dk: 2 * Hk * Dk].view(Hk, Dk, Hin), qkv[2 * Hk * Dk:].view(Hk, r * Dv, Hin)
                    yield pre + "in_proj_qkvz.weight", torch.cat([q, k, v, z.view(Hk, r * Dv, Hin)], 1).reshape(-1, Hin)
                if "in_proj_b" in d and "in_proj_a" in d:
                    b, a = d.pop("in_proj_b"), d.pop("in_proj_a")
                    Hin = b.shape[1]
                    yield pre + "in_proj_ba.weight", torch.cat([b.view(Hk, r, Hin), a.view(Hk, r, Hin)], 1).reshape(-1, Hin)

        return super().load_weights(regroup())


class Qwen3_5MoeForCausalLM(Qwen3_5ForCausalLM):
    pass
