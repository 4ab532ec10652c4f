"""State-space and hybrid decoders: Mamba-2 (``Mamba2ForCausalLM``) and Nemotron-H style hybrids (``NemotronHForCausalLM``: every layer is
``x + mixer(norm(x))`` with the mixer chosen per layer by a pattern string — ``M`` Mamba-2, ``*`` attention, ``-`` MLP).

Parity: reference ``python/sglang/srt/models/{mamba2,nemotron_h}.py``, ``layers/attention/mamba/*`` (mixer, causal-conv1d and selective-scan
ops), ``mem_cache/memory_pool.py`` ``MambaPool`` (per-request conv + SSM state), hybrid-linear attention backend wrapper.

How it fits this engine:

* recurrent state lives in a ``MambaPool`` indexed by the *request slot* (the same index the page table uses), so continuous batching,
  chunked prefill (the state simply carries over between chunks), retraction (a re-prefill starts from position 0 and re-initialises the
  state) and the overlap scheduler need no special cases; attention layers of a hybrid use the ordinary paged KV pool;
* prefix caching is switched off for these families (a radix hit would skip tokens whose state was never computed — the reference's
  ``MambaRadixCache`` keeps state snapshots; not carried over yet);
* prefill uses the chunked *SSD* formulation of Mamba-2 (intra-chunk work is three batched GEMM-shaped contractions, inter-chunk a short
  scan over chunk states) — the shape a tcgen05 kernel wants; decode is the one-step recurrence.  Both are plain PyTorch here: exactness
  first (tests check the SSD path against the step recurrence), a fused sm_100a selective-scan/conv kernel is the follow-up;
* tensor parallelism shards heads (and B/C groups) — z | x | B | C | dt blocks of ``in_proj`` and the conv channels are sliced per block.
"""
from __future__ import annotations

import zlib
from typing import Dict, Iterable, List, Optional, Tuple

import torch
import torch.nn.functional as F
from torch import nn

from .. import ops
from ..configs.model_config import ModelConfig
from ..layers.common import RadixAttention, RMSNorm
from ..layers.linear import ColumnParallelLinear, ParallelLMHead, QKVParallelLinear, RowParallelLinear, VocabParallelEmbedding
from ..model_executor.forward_batch import ForwardBatch
from ..parallel.parallel_state import get_pp_group, get_tp_group
from .llama import _DT


class MambaPool:
    """conv_state [L, slots, conv_dim, k-1], ssm_state [L, slots, H, P, N] (fp32), slot = request slot"""

    def __init__(self, num_layers: int, slots: int, conv_dim: int, d_conv: int, heads: int, head_dim: int, d_state: int, dtype, device):
        self.conv = torch.zeros(num_layers, slots, conv_dim, d_conv - 1, dtype=dtype, device=device)
        self.ssm = torch.zeros(num_layers, slots, heads, head_dim, d_state, dtype=torch.float32, device=device)

    def bytes(self) -> int:
        return self.conv.numel() * self.conv.element_size() + self.ssm.numel() * 4


def _segsum(x: torch.Tensor) -> torch.Tensor:
    """stable segment sum: out[..., i, j] = sum_{j < k <= i} x[..., k] (lower triangle), -inf above the diagonal"""
    T = x.size(-1)
    x = x[..., None].expand(*x.shape, T)
    lower = torch.tril(torch.ones(T, T, dtype=torch.bool, device=x.device), -1)
    x = x.masked_fill(~lower, 0)
    s = torch.cumsum(x, dim=-2)
    keep = torch.tril(torch.ones(T, T, dtype=torch.bool, device=x.device), 0)
    return s.masked_fill(~keep, float("-inf"))


def ssd_chunk_scan(x: torch.Tensor, dt: torch.Tensor, A: torch.Tensor, B: torch.Tensor, C: torch.Tensor, state: torch.Tensor,
                   chunk: int = 64, return_states: bool = False):
    """Mamba-2 SSD over one sequence.  x [T, H, P], dt [T, H] (after softplus), A [H] (negative), B / C [T, H, N] (already expanded from
    groups), state [H, P, N] fp32 -> (y [T, H, P] fp32 WITHOUT the D skip term, final state)."""
    T, H, P = x.shape
    pad = (-T) % chunk
    if pad:
        z = lambda t: F.pad(t, (0, 0) * (t.dim() - 1) + (0, pad))
        x, dt, B, C = z(x), z(dt), z(B), z(C)          # dt = 0 on the padding: decay 1, no input -> the state is untouched
    c = (T + pad) // chunk
    X = (x.float() * dt.float()[..., None]).view(c, chunk, H, P)
    Ad = (dt.float() * A.float()[None, :]).view(c, chunk, H).permute(2, 0, 1)           # [H, c, q]
    Bc, Cc = B.float().view(c, chunk, H, -1), C.float().view(c, chunk, H, -1)
    Acs = torch.cumsum(Ad, dim=-1)
    L = torch.exp(_segsum(Ad))                                                          # [H, c, q, q]
    Yd = torch.einsum("clhn,cshn,hcls,cshp->clhp", Cc, Bc, L, X)
    decay = torch.exp(Acs[..., -1:] - Acs)                                              # [H, c, q]
    st = torch.einsum("clhn,hcl,clhp->chpn", Bc, decay, X)                              # chunk-local end states
    st = torch.cat([state.float()[None], st], 0)                                        # [c + 1, H, P, N]
    dc = torch.exp(_segsum(F.pad(Acs[..., -1], (1, 0))))                                # [H, c + 1, c + 1]
    ns = torch.einsum("hzc,chpn->zhpn", dc, st)
    prev, final = ns[:-1], ns[-1]
    Yo = torch.einsum("clhn,chpn,hcl->clhp", Cc, prev, torch.exp(Acs))
    y = (Yd + Yo).reshape(c * chunk, H, P)[:T]
    return (y, final, ns) if return_states else (y, final)      # ns[j] = state after j * chunk tokens of this call


def ssm_step(x, dt, A, B, C, state):
    """one token per row: x [R, H, P], dt [R, H], B / C [R, H, N], state [R, H, P, N] (updated in place) -> y [R, H, P] fp32"""
    dA = torch.exp(dt.float() * A.float()[None, :])
    state.mul_(dA[..., None, None]).add_((dt.float()[..., None] * x.float())[..., None] * B.float()[:, :, None, :])
    return torch.einsum("rhpn,rhn->rhp", state, C.float())


class Mamba2Mixer(nn.Module):
    def __init__(self, cfg: ModelConfig, layer_idx: int, dtype, prefix: str):
        super().__init__()
        tp = get_tp_group()
        self.tp = tp
        W = tp.world_size
        m = cfg.mamba
        self.H_total, self.P, self.N, self.G_total, self.k = m["num_heads"], m["head_dim"], m["d_state"], m["n_groups"], m["d_conv"]
        assert self.H_total % W == 0 and (self.G_total % W == 0 or self.G_total == 1), \
            "Mamba-2 TP: heads must be divisible by tp_size, B/C groups divisible or a single (replicated) group"
        self.group_replicated = self.G_total % W != 0
        self.H, self.G = self.H_total // W, (self.G_total if self.group_replicated else self.G_total // W)
        self.d_inner = self.H * self.P
        self.conv_dim = self.d_inner + 2 * self.G * self.N
        self.layer_idx = layer_idx                      # index into the MambaPool
        self.chunk = m.get("chunk_size", 64)
        proj_out = 2 * self.d_inner + 2 * self.G * self.N + self.H
        self.in_proj = nn.Linear(cfg.hidden_size, proj_out, bias=m.get("use_bias", False), dtype=dtype)
        self.conv_weight = nn.Parameter(torch.zeros(self.conv_dim, self.k, dtype=dtype), requires_grad=False)
        self.conv_bias = nn.Parameter(torch.zeros(self.conv_dim, dtype=dtype), requires_grad=False) if m.get("use_conv_bias", True) else None
        self.A_log = nn.Parameter(torch.zeros(self.H, dtype=torch.float32), requires_grad=False)
        self.D = nn.Parameter(torch.ones(self.H, dtype=torch.float32), requires_grad=False)
        self.dt_bias = nn.Parameter(torch.zeros(self.H, dtype=torch.float32), requires_grad=False)
        self.norm_weight = nn.Parameter(torch.ones(self.d_inner, dtype=dtype), requires_grad=False)
        ng = int(m.get("norm_groups", self.G_total))          # Nemotron-H: one RMS group per B/C group; HF Mamba-2: a single full-width group
        assert ng == 1 or ng % W == 0
        self.norm_groups = max(1, ng // W)
        self.norm_full_width = ng == 1 and W > 1                # the variance then spans all ranks' channels
        self.use_norm, self.norm_before_gate = bool(m.get("rms_norm", True)), bool(m.get("norm_before_gate", False))
        self.eps = cfg.rms_norm_eps
        num = lambda v, d: float(v) if isinstance(v, (int, float)) else d       # JSON has no literal for infinity: anything else means "unbounded"
        lim = m.get("dt_limit") or (0.0, float("inf"))
        lim = (num(lim[0], 0.0), num(lim[1], float("inf")))
        self.dt_limit = lim if (lim[0] > 0 or lim[1] != float("inf")) else None
        self.out_proj = RowParallelLinear(self.H_total * self.P, cfg.hidden_size, m.get("use_bias", False), True, dtype, None, prefix + ".out_proj")
        for p in self.in_proj.parameters():
            p.requires_grad_(False)

    # ---- sharded loading: the fused tensors are [z | x | B | C | dt] (in_proj rows) and [x | B | C] (conv channels)
    def _blocks(self, with_z_dt: bool) -> List[Tuple[int, int]]:
        W, r = self.tp.world_size, self.tp.rank
        di, gn = self.H_total * self.P, self.G_total * self.N
        segs = ([di] if with_z_dt else []) + [di, gn, gn] + ([self.H_total] if with_z_dt else [])
        out, off = [], 0
        for s in segs:
            if s == gn and self.group_replicated:
                out.append((off, s))
            else:
                out.append((off + r * (s // W), s // W))
            off += s
        return out

    def load_fused(self, param: torch.Tensor, w: torch.Tensor, with_z_dt: bool) -> None:
        parts = [w.narrow(0, o, n) for o, n in self._blocks(with_z_dt)]
        param.data.copy_(torch.cat(parts, 0).reshape(param.shape))

    def load_heads(self, param: torch.Tensor, w: torch.Tensor, per_head: int = 1) -> None:
        n = param.shape[0]
        param.data.copy_(w.narrow(0, self.tp.rank * n, n).to(param.dtype))

    # ---- compute
    def _gated_norm(self, y: torch.Tensor, z: torch.Tensor) -> torch.Tensor:
        if not self.use_norm:                                   # Falcon-H1 without mamba_rms_norm: gate only
            return (y.float() * F.silu(z.float())).to(z.dtype)
        if self.norm_before_gate:                               # normalise first, gate afterwards
            g = self.norm_groups
            yg = y.float().view(y.shape[0], g, -1)
            yg = yg * torch.rsqrt(yg.pow(2).mean(-1, keepdim=True) + self.eps)
            return (yg.view(y.shape[0], -1) * self.norm_weight.float() * F.silu(z.float())).to(z.dtype)
        y = y.float() * F.silu(z.float())
        if self.norm_full_width:
            ss = y.pow(2).sum(-1, keepdim=True)
            ss = self.tp.all_reduce(ss.contiguous())
            return (y * torch.rsqrt(ss / (self.H_total * self.P) + self.eps) * self.norm_weight.float()).to(z.dtype)
        g = self.norm_groups
        yg = y.view(y.shape[0], g, -1)
        yg = yg * torch.rsqrt(yg.pow(2).mean(-1, keepdim=True) + self.eps)
        return (yg.view(y.shape[0], -1) * self.norm_weight.float()).to(z.dtype)

    def forward(self, h: torch.Tensor, fb: ForwardBatch) -> torch.Tensor:
        pool: MambaPool = fb.kv_pool.mamba
        conv_s, ssm_s = pool.conv[self.layer_idx], pool.ssm[self.layer_idx]
        T = h.shape[0]
        zxbcdt = self.in_proj(h)
        di, gn = self.d_inner, self.G * self.N
        z, xBC, dt = zxbcdt.split([di, di + 2 * gn, self.H], dim=-1)
        dt = F.softplus(dt.float() + self.dt_bias)
        if self.dt_limit is not None:
            dt = dt.clamp(self.dt_limit[0], self.dt_limit[1])
        A = -torch.exp(self.A_log)
        rep = self.H // self.G
        slots = fb.req_slots.long()
        # causal conv + SiLU over the whole ragged batch, rolling per-request state updated in place (csrc/kernels/mamba.cu; fp32 out)
        u_all = ops.causal_conv1d(xBC, self.conv_weight, self.conv_bias, conv_s, fb.cu_q, fb.req_slots, fb.seq_lens, fb.max_q_len, silu=True)
        if fb.cu_q is None:      # decode: one token per request — ONE pass over the fp32 state (update + read-out + D skip)
            x, Bm, Cm = u_all.split([di, gn, gn], dim=-1)
            y = ops.selective_state_update(ssm_s, fb.req_slots, x, dt, A.float(), Bm, Cm, self.D.float())
        else:
            ys = []
            ql, sl = fb.q_lens_cpu, fb.seq_lens_cpu
            off = 0
            for b in range(len(ql)):
                n, s = int(ql[b]), int(slots[b])
                fresh = int(sl[b]) - n == 0
                x, Bm, Cm = u_all[off:off + n].split([di, gn, gn], dim=-1)
                st0 = torch.zeros_like(ssm_s[s]) if fresh else ssm_s[s]
                xh = x.view(n, self.H, self.P)
                snap = (fb.mamba_snap or {}).get(b)
                yb, stN, ns = ssd_chunk_scan(xh, dt[off:off + n], A, Bm.view(n, self.G, self.N).repeat_interleave(rep, 1),
                                             Cm.view(n, self.G, self.N).repeat_interleave(rep, 1), st0, self.chunk, return_states=True)
                ssm_s[s] = stN
                if snap is not None:       # prefix-cache snapshot at a scan boundary inside this chunk (mem_cache/mamba_radix.py)
                    rel, row = snap
                    ssm_s[row] = ns[rel // self.chunk]
                    conv_s[row] = xBC[off + rel - (self.k - 1): off + rel].t().to(conv_s.dtype)
                ys.append(yb + self.D[None, :, None] * xh.float())
                off += n
            y = torch.cat(ys, 0) if ys else torch.zeros(0, self.H, self.P, device=h.device)
        y = self._gated_norm(y.reshape(T, di), z)
        return self.out_proj(y.to(h.dtype))


class _ReluSqMLP(nn.Module):
    def __init__(self, cfg: ModelConfig, dtype, prefix: str):
        super().__init__()
        self.up_proj = ColumnParallelLinear(cfg.hidden_size, cfg.intermediate_size, cfg.mlp_bias, False, dtype, None, prefix + ".up_proj")
        self.down_proj = RowParallelLinear(cfg.intermediate_size, cfg.hidden_size, cfg.mlp_bias, True, dtype, None, prefix + ".down_proj")

    def forward(self, x, fb=None):
        return self.down_proj(torch.relu(self.up_proj(x)).pow(2))


class _Attn(nn.Module):
    """Nemotron-H attention block: no positional encoding (the SSM layers carry order), GQA"""

    def __init__(self, cfg: ModelConfig, attn_idx: int, dtype, prefix: str):
        super().__init__()
        self.qkv_proj = QKVParallelLinear(cfg.hidden_size, cfg.head_dim, cfg.num_attention_heads, cfg.num_key_value_heads, cfg.attention_bias,
                                          dtype, None, prefix + ".qkv_proj")
        self.o_proj = RowParallelLinear(cfg.num_attention_heads * cfg.head_dim, cfg.hidden_size, cfg.attention_bias, True, dtype, None,
                                        prefix + ".o_proj")
        self.q_size, self.kv_size = self.qkv_proj.q_size, self.qkv_proj.kv_size
        self.attn = RadixAttention(self.qkv_proj.num_heads, cfg.head_dim, self.qkv_proj.num_kv_heads, attn_idx, rope=None)

    def forward(self, x, fb: ForwardBatch):
        q, k, v = self.qkv_proj(x).split([self.q_size, self.kv_size, self.kv_size], dim=-1)
        return self.o_proj(self.attn(q.contiguous(), k.contiguous(), v.contiguous(), fb))


class _Block(nn.Module):
    def __init__(self, kind: str, mixer: nn.Module, cfg: ModelConfig, dtype):
        super().__init__()
        self.kind = kind
        self.norm = RMSNorm(cfg.hidden_size, cfg.rms_norm_eps, dtype)
        self.mixer = mixer
        if kind == "attention":
            self.self_attn = mixer          # the runner sizes the KV pool from ``layers[i].self_attn.attn``

    def forward(self, x, fb):
        y = self.mixer(self.norm(x), fb)
        y = y.resolve() if hasattr(y, "resolve") else y
        return x + y


class HybridBackbone(nn.Module):
    def __init__(self, cfg: ModelConfig, device):
        super().__init__()
        self.cfg = cfg
        dtype = _DT[cfg.dtype]
        assert get_pp_group().world_size == 1, "hybrid / SSM decoders: pipeline parallelism is not wired"
        self.embed_tokens = VocabParallelEmbedding(cfg.vocab_size, cfg.hidden_size, dtype)
        kinds = {"M": "mamba", "*": "attention", "-": "mlp"}
        pat = cfg.mamba.get("pattern") or "M" * cfg.num_hidden_layers
        assert len(pat) == cfg.num_hidden_layers, "hybrid pattern length must equal num_hidden_layers"
        blocks, n_m, n_a = [], 0, 0
        for i, ch in enumerate(pat):
            kind = kinds[ch]
            pre = f"model.layers.{i}.mixer"
            if kind == "mamba":
                mixer, n_m = Mamba2Mixer(cfg, n_m, dtype, pre), n_m + 1
            elif kind == "attention":
                mixer, n_a = _Attn(cfg, n_a, dtype, pre), n_a + 1
            else:
                mixer = _ReluSqMLP(cfg, dtype, pre)
            blocks.append(_Block(kind, mixer, cfg, dtype))
        self.all_layers = nn.ModuleList(blocks)
        self.num_mamba, self.num_attn = n_m, n_a
        self.norm = RMSNorm(cfg.hidden_size, cfg.rms_norm_eps, dtype)

    @property
    def layers(self):
        """what the runner iterates to size the paged KV pool: the attention blocks only"""
        return [b for b in self.all_layers if b.kind == "attention"]

    def forward(self, fb: ForwardBatch, inputs_embeds=None):
        if inputs_embeds is not None:
            h = inputs_embeds
        elif fb.input_ids is None:
            h = self.embed_tokens(fb.next_token_buf, fb.req_slots)
        else:
            h = self.embed_tokens(fb.input_ids)
        for blk in self.all_layers:
            h = blk(h, fb)
        return self.norm(h)


class NemotronHForCausalLM(nn.Module):
    """canonical names: model.embed_tokens, model.layers.N.norm, model.layers.N.mixer.{in_proj, conv1d, A_log, D, dt_bias, norm, out_proj |
    q_proj, k_proj, v_proj, o_proj | up_proj, down_proj}, model.norm, lm_head.  HF checkpoints say ``backbone.`` / ``embeddings`` / ``norm_f``."""
    supports_cuda_graph = False
    supports_mamba_snapshots = True         # Mamba-2 mixers export the chunk-boundary state of the SSD scan (Mamba2Mixer.forward)
    snapshot_granularity = 64
    disable_radix_cache = True              # unless --mamba-cache-slots > 0 (recurrent-state snapshots, mem_cache/mamba_radix.py)
    stacked_params = [(".qkv_proj", ".q_proj", "q"), (".qkv_proj", ".k_proj", "k"), (".qkv_proj", ".v_proj", "v")]

    def __init__(self, cfg: ModelConfig, quant_name: Optional[str] = None, device="cpu"):
        super().__init__()
        self.cfg = cfg
        self.model = HybridBackbone(cfg, device)
        self.lm_head = None if cfg.tie_word_embeddings else ParallelLMHead(cfg.vocab_size, cfg.hidden_size, _DT[cfg.dtype])

    def forward(self, fb: ForwardBatch, inputs_embeds=None):
        return self.model(fb, inputs_embeds)

    def compute_logits(self, hidden: torch.Tensor, key: str = "logits") -> torch.Tensor:
        head = self.lm_head if self.lm_head is not None else self.model.embed_tokens
        y = ops.linear(hidden, head.weight)
        tp = get_tp_group()
        if tp.world_size > 1:
            y = tp.all_gather(y, dim=-1)
        return y[:, : self.cfg.vocab_size]

    # ---- recurrent state
    def make_mamba_pool(self, slots: int, device) -> Optional[MambaPool]:
        m = next((b.mixer for b in self.model.all_layers if b.kind == "mamba"), None)
        if m is None:
            return None
        return MambaPool(self.model.num_mamba, slots, m.conv_dim, m.k, m.H, m.P, m.N, _DT[self.cfg.dtype], device)

    # ---- weights
    @staticmethod
    def remap(name: str) -> str:
        name = name.replace("backbone.", "model.", 1) if name.startswith("backbone.") else name
        return name.replace("model.embeddings.", "model.embed_tokens.").replace("model.embedding.", "model.embed_tokens.").replace(
            "model.norm_f.", "model.norm.")

    def load_weights(self, weights: Iterable[Tuple[str, torch.Tensor]]) -> List[str]:
        params = dict(self.named_parameters())
        mods = dict(self.named_modules())
        loaded = []
        for name, w in weights:
            name = self.remap(name).replace("model.layers.", "model.all_layers.")
            if name == "lm_head.weight" and self.lm_head is None:
                continue
            mixer_name = name[: name.index(".mixer") + 6] if ".mixer." in name else None
            mixer = mods.get(mixer_name) if mixer_name else None
            if isinstance(mixer, Mamba2Mixer):
                leaf = name[len(mixer_name) + 1:]
                if leaf in ("in_proj.weight", "in_proj.bias"):
                    mixer.load_fused(params[name], w, True)
                elif leaf == "conv1d.weight":
                    mixer.load_fused(mixer.conv_weight, w.reshape(w.shape[0], -1), False)
                elif leaf == "conv1d.bias":
                    mixer.load_fused(mixer.conv_bias, w, False)
                elif leaf in ("A_log", "D", "dt_bias"):
                    mixer.load_heads(getattr(mixer, leaf), w)
                elif leaf == "norm.weight":
                    mixer.load_heads(mixer.norm_weight, w)
                elif leaf.startswith("out_proj."):
                    mixer.out_proj.weight_loader(params[name], w)
                else:
                    continue
                loaded.append(name)
                continue
            target, shard = name, None
            for merged, orig, sid in self.stacked_params:
                if orig + "." in name:
                    target, shard = name.replace(orig, merged), sid
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

    def dummy_checkpoint(self, cfg: ModelConfig, device, dtype, seed: int = 0, scale: float = 0.02):
        """full-size (unsharded) synthetic checkpoint in HF naming, independent of the TP layout"""
        def rnd(name, *shape, s=scale):
            g = torch.Generator(device="cpu").manual_seed((seed * 1000003 + zlib.crc32(name.encode())) & 0x7FFFFFFF)
            return (torch.randn(*shape, generator=g) * s).to(dtype)

        m = cfg.mamba
        H, P, N, G, k = m["num_heads"], m["head_dim"], m["d_state"], m["n_groups"], m["d_conv"]
        di, hid = H * P, cfg.hidden_size
        yield "backbone.embeddings.weight", rnd("emb", cfg.vocab_size, hid)
        pat = m.get("pattern") or "M" * cfg.num_hidden_layers
        for i, ch in enumerate(pat):
            pre = f"backbone.layers.{i}."
            yield pre + "norm.weight", torch.ones(hid, dtype=dtype)
            if ch == "M":
                yield pre + "mixer.in_proj.weight", rnd(pre + "in", 2 * di + 2 * G * N + H, hid)
                yield pre + "mixer.conv1d.weight", rnd(pre + "cw", di + 2 * G * N, 1, k, s=0.3)
                yield pre + "mixer.conv1d.bias", rnd(pre + "cb", di + 2 * G * N, s=0.1)
                g = torch.Generator(device="cpu").manual_seed(zlib.crc32((pre + "A").encode()) & 0x7FFFFFFF)
                yield pre + "mixer.A_log", torch.log(torch.rand(H, generator=g) * 3 + 0.5)
                yield pre + "mixer.D", torch.ones(H) + rnd(pre + "D", H, s=0.1).float()
                yield pre + "mixer.dt_bias", rnd(pre + "dtb", H, s=0.5).float() - 1.0
                yield pre + "mixer.norm.weight", torch.ones(di, dtype=dtype) + rnd(pre + "nw", di, s=0.05)
                yield pre + "mixer.out_proj.weight", rnd(pre + "out", hid, di)
            elif ch == "*":
                hd, nh, nkv = cfg.head_dim, cfg.num_attention_heads, cfg.num_key_value_heads
                yield pre + "mixer.q_proj.weight", rnd(pre + "q", nh * hd, hid)
                yield pre + "mixer.k_proj.weight", rnd(pre + "k", nkv * hd, hid)
                yield pre + "mixer.v_proj.weight", rnd(pre + "v", nkv * hd, hid)
                yield pre + "mixer.o_proj.weight", rnd(pre + "o", hid, nh * hd)
            else:
                yield pre + "mixer.up_proj.weight", rnd(pre + "up", cfg.intermediate_size, hid)
                yield pre + "mixer.down_proj.weight", rnd(pre + "down", hid, cfg.intermediate_size)
        yield "backbone.norm_f.weight", torch.ones(hid, dtype=dtype)
        if not cfg.tie_word_embeddings:
            yield "lm_head.weight", rnd("lm_head", cfg.vocab_size, hid)


class Mamba2ForCausalLM(NemotronHForCausalLM):
    """pure Mamba-2 stack (every layer is an SSM mixer); same checkpoint naming (``backbone.layers.N.mixer.*``)"""


# ================================================================================================ Bamba (Mamba-2 / attention mixers + a gated MLP in EVERY layer)
class _BambaBlock(nn.Module):
    def __init__(self, cfg: ModelConfig, kind: str, idx: int, gid: int, dtype, rope):
        super().__init__()
        from .llama import LlamaAttention, LlamaMLP

        self.kind = kind
        pre = f"model.layers.{gid}"
        self.input_layernorm = RMSNorm(cfg.hidden_size, cfg.rms_norm_eps, dtype)
        self.pre_ff_layernorm = RMSNorm(cfg.hidden_size, cfg.rms_norm_eps, dtype)
        if kind == "mamba":
            self.mamba = Mamba2Mixer(cfg, idx, dtype, pre + ".mamba")
        else:
            self.self_attn = LlamaAttention(cfg, idx, dtype, None, rope, pre + ".self_attn")
        self.feed_forward = LlamaMLP(cfg, dtype, None, pre + ".feed_forward")

    def forward(self, x, fb):
        h = self.input_layernorm(x)
        y = self.mamba(h, fb) if self.kind == "mamba" else self.self_attn(h, fb)
        x = x + (y.resolve() if hasattr(y, "resolve") else y)
        m = self.feed_forward(self.pre_ff_layernorm(x))
        return x + (m.resolve() if hasattr(m, "resolve") else m)


class BambaBackbone(nn.Module):
    def __init__(self, cfg: ModelConfig, device):
        super().__init__()
        from ..layers.common import RotaryEmbedding

        self.cfg = cfg
        dtype = _DT[cfg.dtype]
        assert get_pp_group().world_size == 1
        self.embed_tokens = VocabParallelEmbedding(cfg.vocab_size, cfg.hidden_size, dtype)
        rope = RotaryEmbedding(cfg.head_dim, cfg.max_position_embeddings, cfg.rope_theta, cfg.rope_scaling, cfg.rope_is_neox,
                               cfg.partial_rotary_factor, device)
        pat = cfg.mamba["pattern"]
        blocks, n_m, n_a = [], 0, 0
        for i, ch in enumerate(pat):
            if ch == "M":
                blocks.append(_BambaBlock(cfg, "mamba", n_m, i, dtype, rope))
                n_m += 1
            else:
                blocks.append(_BambaBlock(cfg, "attention", n_a, i, dtype, rope))
                n_a += 1
        self.all_layers = nn.ModuleList(blocks)
        self.num_mamba, self.num_attn = n_m, n_a
        self.norm = RMSNorm(cfg.hidden_size, cfg.rms_norm_eps, dtype)

    @property
    def layers(self):
        return [b for b in self.all_layers if b.kind == "attention"]

    def forward(self, fb: ForwardBatch, inputs_embeds=None):
        if inputs_embeds is not None:
            h = inputs_embeds
        elif fb.input_ids is None:
            h = self.embed_tokens(fb.next_token_buf, fb.req_slots)
        else:
            h = self.embed_tokens(fb.input_ids)
        for blk in self.all_layers:
            h = blk(h, fb)
        return self.norm(h)


class BambaForCausalLM(NemotronHForCausalLM):
    stacked_params = [(".qkv_proj", ".q_proj", "q"), (".qkv_proj", ".k_proj", "k"), (".qkv_proj", ".v_proj", "v"),
                      (".gate_up_proj", ".gate_proj", 0), (".gate_up_proj", ".up_proj", 1)]

    def __init__(self, cfg: ModelConfig, quant_name: Optional[str] = None, device="cpu"):
        nn.Module.__init__(self)
        self.cfg = cfg
        self.model = BambaBackbone(cfg, device)
        self.lm_head = None if cfg.tie_word_embeddings else ParallelLMHead(cfg.vocab_size, cfg.hidden_size, _DT[cfg.dtype])

    def make_mamba_pool(self, slots: int, device) -> Optional[MambaPool]:
        m = next((b.mamba for b in self.model.all_layers if b.kind == "mamba"), None)
        if m is None:
            return None
        return MambaPool(self.model.num_mamba, slots, m.conv_dim, m.k, m.H, m.P, m.N, _DT[self.cfg.dtype], device)

    def load_weights(self, weights: Iterable[Tuple[str, torch.Tensor]]) -> List[str]:
        from .llama import _load_stacked_bias

        params = dict(self.named_parameters())
        mods = dict(self.named_modules())
        loaded = []
        for name, w in weights:
            name = name.replace("model.final_layernorm.", "model.norm.").replace("model.layers.", "model.all_layers.")
            if name == "lm_head.weight" and self.lm_head is None:
                continue
            if ".mamba." in name:
                mixer_name = name[: name.index(".mamba") + 6]
                mixer, leaf = mods[mixer_name], name[name.index(".mamba") + 7:]
                if leaf in ("in_proj.weight", "in_proj.bias"):
                    mixer.load_fused(params[name], w, True)
                elif leaf == "conv1d.weight":
                    mixer.load_fused(mixer.conv_weight, w.reshape(w.shape[0], -1), False)
                elif leaf == "conv1d.bias":
                    mixer.load_fused(mixer.conv_bias, w, False)
                elif leaf in ("A_log", "D", "dt_bias"):
                    mixer.load_heads(getattr(mixer, leaf), w)
                elif leaf == "norm.weight":
                    mixer.load_heads(mixer.norm_weight, w)
                elif leaf.startswith("out_proj."):
                    mixer.out_proj.weight_loader(params[name], w)
                else:
                    continue
                loaded.append(name)
                continue
            target, shard = name, None
            for merged, orig, sid in self.stacked_params:
                if orig + "." in name:
                    target, shard = name.replace(orig, merged), sid
                    break
            if target not in params:
                continue
            mod = mods[target.rsplit(".", 1)[0]]
            if hasattr(mod, "weight_loader"):
                if target.endswith(".bias") and shard is not None:
                    _load_stacked_bias(mod, params[target], w, shard)
                else:
                    mod.weight_loader(params[target], w, shard) if shard is not None else mod.weight_loader(params[target], w)
            else:
                params[target].data.copy_(w)
            loaded.append(name)
        return loaded


# ================================================================================================ Mamba-1 (selective scan per channel): Mamba, FalconMamba, Jamba
def selective_scan_chunked(u, dt, A, B, C, state, chunk: int = 16, snap_at: Optional[int] = None):
    """u, dt [T, Dn]; A [Dn, N] (negative); B, C [T, N]; state [Dn, N] fp32 -> (y [T, Dn] without the D skip, final state).
    h_t = exp(dt_t A) h_{t-1} + dt_t B_t u_t, y_t = h_t . C_t.  Inside a chunk the recurrence is solved in closed form with decay factors
    exp(sum_{s<k<=t} dt_k A) (all <= 1, so nothing overflows); chunks are chained through the carried state."""
    T, Dn = u.shape
    u, dt, B, C, A = u.float(), dt.float(), B.float(), C.float(), A.float()
    ys = []
    snap = None
    for s in range(0, T, chunk):
        e = min(T, s + chunk)
        q = e - s
        la = dt[s:e, :, None] * A[None]                                   # [q, Dn, N] log decay per step
        cs = torch.cumsum(la, 0)
        inp = (dt[s:e] * u[s:e])[:, :, None] * B[s:e, None, :]            # [q, Dn, N]
        # L[t, k] = exp(cs_t - cs_k) for k <= t
        diff = cs[:, None] - cs[None, :]                                  # [q, q, Dn, N]
        mask = torch.tril(torch.ones(q, q, dtype=torch.bool, device=u.device))[:, :, None, None]
        L = torch.exp(diff.masked_fill(~mask, float("-inf")))
        h = torch.exp(cs) * state[None] + torch.einsum("tkdn,kdn->tdn", L, inp)
        ys.append(torch.einsum("tdn,tn->td", h, C[s:e]))
        state = h[-1]
        if snap_at is not None and e == snap_at:
            snap = state.clone()          # prefix-cache snapshot at a chunk boundary (mem_cache/mamba_radix.py)
    y = torch.cat(ys, 0) if ys else u.new_zeros(0, Dn)
    return (y, state) if snap_at is None else (y, state, snap)


class Mamba1Mixer(nn.Module):
    def __init__(self, cfg: ModelConfig, layer_idx: int, dtype, prefix: str):
        super().__init__()
        tp = get_tp_group()
        self.tp = tp
        W = tp.world_size
        m = cfg.mamba
        self.Dn_total, self.N, self.k, self.R = m["d_inner"], m["d_state"], m["d_conv"], m["dt_rank"]
        assert self.Dn_total % W == 0
        self.Dn = self.Dn_total // W
        self.layer_idx = layer_idx
        self.in_proj = nn.Linear(cfg.hidden_size, 2 * self.Dn, bias=m.get("use_bias", False), dtype=dtype)          # local [x | z] channels
        self.conv_weight = nn.Parameter(torch.zeros(self.Dn, self.k, dtype=dtype), requires_grad=False)
        self.conv_bias = nn.Parameter(torch.zeros(self.Dn, dtype=dtype), requires_grad=False) if m.get("use_conv_bias", True) else None
        self.x_proj = RowParallelLinear(self.Dn_total, self.R + 2 * self.N, False, True, dtype, None, prefix + ".x_proj")
        self.dt_proj = nn.Linear(self.R, self.Dn, bias=True, dtype=dtype)
        self.A_log = nn.Parameter(torch.zeros(self.Dn, self.N, dtype=torch.float32), requires_grad=False)
        self.D = nn.Parameter(torch.ones(self.Dn, dtype=torch.float32), requires_grad=False)
        self.out_proj = RowParallelLinear(self.Dn_total, cfg.hidden_size, m.get("use_bias", False), True, dtype, None, prefix + ".out_proj")
        kind = m.get("bcdt_norm")            # None | "plain" (FalconMamba: weightless RMS) | "weight" (Jamba)
        self.bcdt = kind
        self.eps = m.get("mixer_rms_eps", cfg.rms_norm_eps)
        if kind == "weight":
            self.dt_layernorm, self.b_layernorm, self.c_layernorm = RMSNorm(self.R, cfg.rms_norm_eps, dtype), RMSNorm(self.N, cfg.rms_norm_eps, dtype), \
                RMSNorm(self.N, cfg.rms_norm_eps, dtype)
        for p in list(self.in_proj.parameters()) + list(self.dt_proj.parameters()):
            p.requires_grad_(False)
        self.chunk = m.get("scan_chunk", 16)
        # MambaPool geometry: heads = local channels, head_dim = 1
        self.conv_dim, self.H, self.P = self.Dn, self.Dn, 1

    def shard(self, w: torch.Tensor, dim: int = 0) -> torch.Tensor:
        return w.narrow(dim, self.tp.rank * self.Dn, self.Dn)

    def _n(self, t, mod):
        if self.bcdt == "weight":
            return mod(t.contiguous())
        if self.bcdt == "plain":
            f = t.float()
            return (f * torch.rsqrt(f.pow(2).mean(-1, keepdim=True) + self.eps)).to(t.dtype)
        return t

    def forward(self, h: torch.Tensor, fb: ForwardBatch) -> torch.Tensor:
        pool: MambaPool = fb.kv_pool.mamba
        conv_s, ssm_s = pool.conv[self.layer_idx], pool.ssm[self.layer_idx]          # [slots, Dn, k-1], [slots, Dn, 1, N]
        T = h.shape[0]
        x, z = self.in_proj(h).split([self.Dn, self.Dn], dim=-1)
        A = -torch.exp(self.A_log)
        slots = fb.req_slots.long()
        decode = fb.cu_q is None
        # causal conv (+ bias) + SiLU over the ragged batch, rolling state in place (csrc/kernels/mamba.cu)
        u = ops.causal_conv1d(x, self.conv_weight, self.conv_bias, conv_s, fb.cu_q, fb.req_slots, fb.seq_lens, fb.max_q_len, silu=True).to(h.dtype)
        dbc = self.x_proj(u)
        dbc = dbc.resolve() if hasattr(dbc, "resolve") else dbc
        dtr, Bm, Cm = dbc.split([self.R, self.N, self.N], dim=-1)
        dtr, Bm, Cm = self._n(dtr, getattr(self, "dt_layernorm", None)), self._n(Bm, getattr(self, "b_layernorm", None)), \
            self._n(Cm, getattr(self, "c_layernorm", None))
        dt = F.softplus(self.dt_proj(dtr).float())
        if decode:
            st = ssm_s[slots][:, :, 0]                                                 # [B, Dn, N]
            st = torch.exp(dt[:, :, None] * A[None]) * st + (dt * u.float())[:, :, None] * Bm.float()[:, None, :]
            ssm_s[slots] = st.unsqueeze(2)
            y = torch.einsum("bdn,bn->bd", st, Cm.float())
        else:
            ys, off = [], 0
            for b in range(len(fb.q_lens_cpu)):
                n, s = int(fb.q_lens_cpu[b]), int(slots[b])
                fresh = int(fb.seq_lens_cpu[b]) - n == 0
                st0 = torch.zeros_like(ssm_s[s][:, 0]) if fresh else ssm_s[s][:, 0]
                snap = (fb.mamba_snap or {}).get(b)
                if snap is None:
                    yb, stN = selective_scan_chunked(u[off:off + n], dt[off:off + n], A, Bm[off:off + n], Cm[off:off + n], st0, self.chunk)
                else:       # also export the state at a chunk boundary for the prefix cache
                    rel, row = snap
                    yb, stN, sn = selective_scan_chunked(u[off:off + n], dt[off:off + n], A, Bm[off:off + n], Cm[off:off + n], st0, self.chunk, snap_at=rel)
                    if sn is not None:
                        ssm_s[row] = sn.unsqueeze(1)
                        conv_s[row] = x[off + rel - (self.k - 1): off + rel].t().to(conv_s.dtype)
                ssm_s[s] = stN.unsqueeze(1)
                ys.append(yb)
                off += n
            y = torch.cat(ys, 0) if ys else u.new_zeros(0, self.Dn).float()
        y = (y + self.D[None] * u.float()) * F.silu(z.float())
        return self.out_proj(y.to(h.dtype))


def _load_mamba1(mixer: Mamba1Mixer, leaf: str, w: torch.Tensor, params: Dict[str, torch.Tensor], full: str) -> bool:
    if leaf in ("in_proj.weight", "in_proj.bias"):
        x, z = w.chunk(2, 0)
        params[full].data.copy_(torch.cat([mixer.shard(x), mixer.shard(z)], 0))
    elif leaf == "conv1d.weight":
        mixer.conv_weight.data.copy_(mixer.shard(w.reshape(w.shape[0], -1)))
    elif leaf == "conv1d.bias":
        mixer.conv_bias.data.copy_(mixer.shard(w))
    elif leaf == "x_proj.weight":
        mixer.x_proj.weight_loader(params[full], w)
    elif leaf in ("dt_proj.weight", "dt_proj.bias"):
        params[full].data.copy_(mixer.shard(w))
    elif leaf in ("A_log", "D"):
        getattr(mixer, leaf).data.copy_(mixer.shard(w).float())
    elif leaf.startswith("out_proj."):
        mixer.out_proj.weight_loader(params[full], w) if leaf.endswith("weight") else params[full].data.copy_(w)
    elif leaf.split(".")[0] in ("dt_layernorm", "b_layernorm", "c_layernorm"):
        params[full].data.copy_(w)
    else:
        return False
    return True


class _M1Block(nn.Module):
    def __init__(self, cfg, idx, dtype, prefix):
        super().__init__()
        self.kind = "mamba"
        self.norm = RMSNorm(cfg.hidden_size, cfg.rms_norm_eps, dtype)
        self.mixer = Mamba1Mixer(cfg, idx, dtype, prefix + ".mixer")

    def forward(self, x, fb):
        y = self.mixer(self.norm(x), fb)
        return x + (y.resolve() if hasattr(y, "resolve") else y)


class MambaForCausalLM(NemotronHForCausalLM):
    """Mamba-1 (state-spaces/mamba-*-hf) and FalconMamba (``cfg.mamba['bcdt_norm'] == 'plain'``)"""

    def __init__(self, cfg: ModelConfig, quant_name: Optional[str] = None, device="cpu"):
        nn.Module.__init__(self)
        self.cfg = cfg
        dtype = _DT[cfg.dtype]
        bb = nn.Module()
        bb.embed_tokens = VocabParallelEmbedding(cfg.vocab_size, cfg.hidden_size, dtype)
        bb.all_layers = nn.ModuleList([_M1Block(cfg, i, dtype, f"model.layers.{i}") for i in range(cfg.num_hidden_layers)])
        bb.norm = RMSNorm(cfg.hidden_size, cfg.rms_norm_eps, dtype)
        bb.layers, bb.num_mamba = [], cfg.num_hidden_layers
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
        mx = self.model.all_layers[0].mixer
        return MambaPool(self.model.num_mamba, slots, mx.Dn, mx.k, mx.Dn, 1, mx.N, _DT[self.cfg.dtype], device)

    def load_weights(self, weights):
        params = dict(self.named_parameters())
        mods = dict(self.named_modules())
        loaded = []
        for name, w in weights:
            n = self.remap(name).replace("model.layers.", "model.all_layers.")
            if n == "lm_head.weight" and self.lm_head is None:
                continue
            if ".mixer." in n:
                mn = n[: n.index(".mixer") + 6]
                if _load_mamba1(mods[mn], n[len(mn) + 1:], w, params, n):
                    loaded.append(name)
                continue
            if n in params:
                mod = mods[n.rsplit(".", 1)[0]]
                mod.weight_loader(params[n], w) if hasattr(mod, "weight_loader") else params[n].data.copy_(w)
                loaded.append(name)
        return loaded

    def dummy_checkpoint(self, cfg, device, dtype, seed: int = 0, scale: float = 0.02):
        for n, p in self.named_parameters():
            g = torch.Generator(device="cpu").manual_seed((seed * 1000003 + zlib.crc32(n.encode())) & 0x7FFFFFFF)
            if n.endswith("A_log"):
                p.data.copy_(torch.log(torch.rand(p.shape, generator=g) * 3 + 0.5))
            elif p.dim() > 1:
                p.data.copy_((torch.randn(p.shape, generator=g) * 0.05).to(p.dtype))
            elif "norm" in n or n.endswith(".D"):
                p.data.fill_(1.0)
            else:
                p.data.copy_((torch.randn(p.shape, generator=g) * 0.02).to(p.dtype))
        return iter(())


class FalconMambaForCausalLM(MambaForCausalLM):
    pass


# ---- Jamba: Mamba-1 mixers with normalised dt / B / C, attention without positional encoding, MoE every other layer
class _JambaMoE(nn.Module):
    def __init__(self, cfg: ModelConfig, layer_id: int, dtype, prefix: str):
        super().__init__()
        from ..layers.linear import ReplicatedLinear
        from ..layers.moe import FusedMoE
        from .mixtral import _EP_SIZE

        self.router = ReplicatedLinear(cfg.hidden_size, cfg.num_local_experts, False, dtype, None, prefix + ".router")
        self.experts = FusedMoE(cfg.num_local_experts, cfg.num_experts_per_tok, cfg.hidden_size, cfg.intermediate_size, dtype, False, _EP_SIZE["v"],
                                None, layer_id)

    def forward(self, x):
        return self.experts(x, self.router(x))


class _JambaBlock(nn.Module):
    def __init__(self, cfg: ModelConfig, kind: str, idx: int, gid: int, moe: bool, dtype):
        super().__init__()
        from .llama import LlamaAttention, LlamaMLP

        self.kind = kind
        pre = f"model.layers.{gid}"
        self.input_layernorm = RMSNorm(cfg.hidden_size, cfg.rms_norm_eps, dtype)
        self.pre_ff_layernorm = RMSNorm(cfg.hidden_size, cfg.rms_norm_eps, dtype)
        if kind == "mamba":
            self.mamba = Mamba1Mixer(cfg, idx, dtype, pre + ".mamba")
        else:
            self.self_attn = LlamaAttention(cfg, idx, dtype, None, None, pre + ".self_attn")          # Jamba attention has no rotary embedding
        self.feed_forward = _JambaMoE(cfg, gid, dtype, pre + ".feed_forward") if moe else LlamaMLP(cfg, dtype, None, pre + ".feed_forward")

    def forward(self, x, fb):
        h = self.input_layernorm(x)
        y = self.mamba(h, fb) if self.kind == "mamba" else self.self_attn(h, fb)
        x = x + (y.resolve() if hasattr(y, "resolve") else y)
        m = self.feed_forward(self.pre_ff_layernorm(x))
        return x + (m.resolve() if hasattr(m, "resolve") else m)


class JambaForCausalLM(BambaForCausalLM):
    def __init__(self, cfg: ModelConfig, quant_name: Optional[str] = None, device="cpu"):
        nn.Module.__init__(self)
        self.cfg = cfg
        dtype = _DT[cfg.dtype]
        bb = nn.Module()
        bb.embed_tokens = VocabParallelEmbedding(cfg.vocab_size, cfg.hidden_size, dtype)
        blocks, n_m, n_a = [], 0, 0
        for i, ch in enumerate(cfg.mamba["pattern"]):
            moe = bool(cfg.mamba["moe_layers"][i])
            if ch == "M":
                blocks.append(_JambaBlock(cfg, "mamba", n_m, i, moe, dtype))
                n_m += 1
            else:
                blocks.append(_JambaBlock(cfg, "attention", n_a, i, moe, dtype))
                n_a += 1
        bb.all_layers = nn.ModuleList(blocks)
        bb.num_mamba, bb.num_attn = n_m, n_a
        bb.norm = RMSNorm(cfg.hidden_size, cfg.rms_norm_eps, dtype)
        bb.layers = [b for b in blocks if b.kind == "attention"]
        self.model = bb
        self.lm_head = None if cfg.tie_word_embeddings else ParallelLMHead(cfg.vocab_size, cfg.hidden_size, dtype)

    def forward(self, fb: ForwardBatch, inputs_embeds=None):
        return MambaForCausalLM.forward(self, fb, inputs_embeds)

    def make_mamba_pool(self, slots: int, device) -> Optional[MambaPool]:
        mx = next((b.mamba for b in self.model.all_layers if b.kind == "mamba"), None)
        return None if mx is None else MambaPool(self.model.num_mamba, slots, mx.Dn, mx.k, mx.Dn, 1, mx.N, _DT[self.cfg.dtype], device)

    def load_weights(self, weights):
        from .llama import _load_stacked_bias

        params = dict(self.named_parameters())
        mods = dict(self.named_modules())
        loaded = []
        for name, w in weights:
            n = name.replace("model.final_layernorm.", "model.norm.").replace("model.layers.", "model.all_layers.")
            if n == "lm_head.weight" and self.lm_head is None:
                continue
            if ".mamba." in n:
                mn = n[: n.index(".mamba") + 6]
                if _load_mamba1(mods[mn], n[len(mn) + 1:], w, params, n):
                    loaded.append(name)
                continue
            if ".feed_forward.experts." in n:
                ex = mods[n[: n.index(".experts") + 8]]
                tail = n.split(".experts.")[1]
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
            for merged, orig, sid in self.stacked_params:
                if orig + "." in n:
                    target, shard = n.replace(orig, merged), sid
                    break
            if target not in params:
                continue
            mod = mods[target.rsplit(".", 1)[0]]
            if hasattr(mod, "weight_loader"):
                if target.endswith(".bias") and shard is not None:
                    _load_stacked_bias(mod, params[target], w, shard)
                else:
                    mod.weight_loader(params[target], w, shard) if shard is not None else mod.weight_loader(params[target], w)
            else:
                params[target].data.copy_(w)
            loaded.append(name)
        return loaded


# ================================================================================================ Granite 4 (GraniteMoeHybrid): Mamba-2 / attention mixers, MoE + shared MLP, Granite multipliers
class _GraniteMoE(nn.Module):
    def __init__(self, cfg: ModelConfig, gid: int, dtype, prefix: str):
        super().__init__()
        from ..layers.linear import ReplicatedLinear
        from ..layers.moe import FusedMoE
        from .mixtral import _EP_SIZE

        self.router = ReplicatedLinear(cfg.hidden_size, cfg.num_local_experts, False, dtype, None, prefix + ".router")
        self.experts = FusedMoE(cfg.num_local_experts, cfg.num_experts_per_tok, cfg.hidden_size, cfg.intermediate_size, dtype, True, _EP_SIZE["v"], None, gid)

    def forward(self, x):
        return self.experts(x, self.router(x))


class _SharedMLP(nn.Module):
    def __init__(self, cfg: ModelConfig, dtype, prefix: str):
        super().__init__()
        from ..layers.linear import MergedColumnParallelLinear

        S = cfg.shared_expert_intermediate_size
        self.gate_up_proj = MergedColumnParallelLinear(cfg.hidden_size, [S, S], False, dtype, None, prefix + ".gate_up_proj")
        self.down_proj = RowParallelLinear(S, cfg.hidden_size, False, True, dtype, None, prefix + ".down_proj")

    def forward(self, x):
        return self.down_proj(ops.silu_and_mul(self.gate_up_proj(x)))


class _GraniteHybridBlock(nn.Module):
    def __init__(self, cfg: ModelConfig, kind: str, idx: int, gid: int, dtype, rope):
        super().__init__()
        from .llama import LlamaAttention

        self.kind = kind
        pre = f"model.layers.{gid}"
        self.rm = float(cfg.residual_multiplier or 1.0)
        self.input_layernorm = RMSNorm(cfg.hidden_size, cfg.rms_norm_eps, dtype)
        self.post_attention_layernorm = RMSNorm(cfg.hidden_size, cfg.rms_norm_eps, dtype)
        if kind == "mamba":
            self.mamba = Mamba2Mixer(cfg, idx, dtype, pre + ".mamba")
        else:
            self.self_attn = LlamaAttention(cfg, idx, dtype, None, rope, pre + ".self_attn")
            if cfg.attention_multiplier:
                self.self_attn.attn.scale = float(cfg.attention_multiplier)
        self.block_sparse_moe = _GraniteMoE(cfg, gid, dtype, pre + ".block_sparse_moe") if cfg.num_local_experts > 0 else None
        self.shared_mlp = _SharedMLP(cfg, dtype, pre + ".shared_mlp") if cfg.shared_expert_intermediate_size else None

    def forward(self, x, fb):
        r = lambda y: y.resolve() if hasattr(y, "resolve") else y
        h = self.input_layernorm(x)
        x = x + r(self.mamba(h, fb) if self.kind == "mamba" else self.self_attn(h, fb)) * self.rm
        h = self.post_attention_layernorm(x)
        y = r(self.block_sparse_moe(h)) if self.block_sparse_moe is not None else 0
        if self.shared_mlp is not None:
            y = y + r(self.shared_mlp(h))
        return x + y * self.rm


class GraniteMoeHybridForCausalLM(BambaForCausalLM):
    def __init__(self, cfg: ModelConfig, quant_name: Optional[str] = None, device="cpu"):
        nn.Module.__init__(self)
        from ..layers.common import RotaryEmbedding

        self.cfg = cfg
        dtype = _DT[cfg.dtype]
        bb = nn.Module()
        bb.embed_tokens = VocabParallelEmbedding(cfg.vocab_size, cfg.hidden_size, dtype)
        rope = None
        if cfg.pos_emb == "rope":
            rope = RotaryEmbedding(cfg.head_dim, cfg.max_position_embeddings, cfg.rope_theta, cfg.rope_scaling, cfg.rope_is_neox, cfg.partial_rotary_factor,
                                   device)
        blocks, n_m, n_a = [], 0, 0
        for i, ch in enumerate(cfg.mamba["pattern"]):
            if ch == "M":
                blocks.append(_GraniteHybridBlock(cfg, "mamba", n_m, i, dtype, rope))
                n_m += 1
            else:
                blocks.append(_GraniteHybridBlock(cfg, "attention", n_a, i, dtype, rope))
                n_a += 1
        bb.all_layers = nn.ModuleList(blocks)
        bb.num_mamba, bb.num_attn = n_m, n_a
        bb.layers = [b for b in blocks if b.kind == "attention"]
        bb.norm = RMSNorm(cfg.hidden_size, cfg.rms_norm_eps, dtype)
        self.model = bb
        self.lm_head = None if cfg.tie_word_embeddings else ParallelLMHead(cfg.vocab_size, cfg.hidden_size, dtype)

    def forward(self, fb: ForwardBatch, inputs_embeds=None):
        m = self.model
        h = inputs_embeds if inputs_embeds is not None else (m.embed_tokens(fb.next_token_buf, fb.req_slots) if fb.input_ids is None
                                                             else m.embed_tokens(fb.input_ids))
        if self.cfg.embedding_multiplier:
            h = h * float(self.cfg.embedding_multiplier)
        for blk in m.all_layers:
            h = blk(h, fb)
        return m.norm(h)

    def compute_logits(self, hidden, key: str = "logits"):
        y = super().compute_logits(hidden, key)
        return y / float(self.cfg.logits_scaling) if (self.cfg.logits_scaling or 1.0) != 1.0 else y

    def load_weights(self, weights):
        mods = dict(self.named_modules())

        def rename():
            for name, w in weights:
                n = name.replace("model.layers.", "model.all_layers.")
                if ".block_sparse_moe.router.layer." in n:
                    yield name.replace(".router.layer.", ".router."), w
                elif ".block_sparse_moe.input_linear.weight" in n or ".block_sparse_moe.output_linear.weight" in n:
                    ex = mods[n[: n.index(".block_sparse_moe") + 17] + ".experts"]
                    for e in range(w.shape[0]):
                        if "input_linear" in n:
                            g_, u_ = w[e].chunk(2, 0)
                            ex.weight_loader(ex.w13, g_, e, "w1")
                            ex.weight_loader(ex.w13, u_, e, "w3")
                        else:
                            ex.weight_loader(ex.w2, w[e], e, "w2")
                elif ".shared_mlp.input_linear.weight" in n:
                    g_, u_ = w.chunk(2, 0)
                    yield name.replace("input_linear", "gate_proj"), g_
                    yield name.replace("input_linear", "up_proj"), u_
                elif ".shared_mlp.output_linear.weight" in n:
                    yield name.replace("output_linear", "down_proj"), w
                else:
                    yield name, w
        return super().load_weights(rename())


# ================================================================================================ Falcon-H1: Mamba-2 and attention IN PARALLEL in every layer + muP multipliers
class _FalconH1Block(nn.Module):
    def __init__(self, cfg: ModelConfig, idx: int, dtype, rope):
        super().__init__()
        from .llama import LlamaAttention, LlamaMLP

        self.kind = "attention"            # every layer owns a KV-cache layer (and a MambaPool layer)
        pre = f"model.layers.{idx}"
        self.input_layernorm = RMSNorm(cfg.hidden_size, cfg.rms_norm_eps, dtype)
        self.pre_ff_layernorm = RMSNorm(cfg.hidden_size, cfg.rms_norm_eps, dtype)
        self.mamba = Mamba2Mixer(cfg, idx, dtype, pre + ".mamba")
        self.self_attn = LlamaAttention(cfg, idx, dtype, None, rope, pre + ".self_attn")
        self.feed_forward = LlamaMLP(cfg, dtype, None, pre + ".feed_forward")

    def forward(self, x, fb):
        r = lambda y: y.resolve() if hasattr(y, "resolve") else y
        h = self.input_layernorm(x)
        x = x + r(self.mamba(h, fb)) + r(self.self_attn(h, fb))         # the muP multipliers are folded into the weights at load time
        return x + r(self.feed_forward(self.pre_ff_layernorm(x)))


class FalconH1ForCausalLM(BambaForCausalLM):
    """all scalar muP multipliers (ssm in / out, per-section zxbcdt, attention in / out, key, MLP gate / down) are folded into the projection weights
    when the checkpoint is loaded; only the embedding and LM-head multipliers remain at run time"""

    def __init__(self, cfg: ModelConfig, quant_name: Optional[str] = None, device="cpu"):
        nn.Module.__init__(self)
        from ..layers.common import RotaryEmbedding

        self.cfg = cfg
        dtype = _DT[cfg.dtype]
        bb = nn.Module()
        bb.embed_tokens = VocabParallelEmbedding(cfg.vocab_size, cfg.hidden_size, dtype)
        rope = RotaryEmbedding(cfg.head_dim, cfg.max_position_embeddings, cfg.rope_theta, cfg.rope_scaling, cfg.rope_is_neox, cfg.partial_rotary_factor, device)
        blocks = [_FalconH1Block(cfg, i, dtype, rope) for i in range(cfg.num_hidden_layers)]
        bb.all_layers = nn.ModuleList(blocks)
        bb.num_mamba = bb.num_attn = cfg.num_hidden_layers
        bb.layers = blocks
        bb.norm = RMSNorm(cfg.hidden_size, cfg.rms_norm_eps, dtype)
        self.model = bb
        self.lm_head = None if cfg.tie_word_embeddings else ParallelLMHead(cfg.vocab_size, cfg.hidden_size, dtype)

    def forward(self, fb: ForwardBatch, inputs_embeds=None):
        m = self.model
        h = inputs_embeds if inputs_embeds is not None else (m.embed_tokens(fb.next_token_buf, fb.req_slots) if fb.input_ids is None
                                                             else m.embed_tokens(fb.input_ids))
        h = h * float(self.cfg.embedding_multiplier or 1.0)
        for blk in m.all_layers:
            h = blk(h, fb)
        return m.norm(h)

    def compute_logits(self, hidden, key: str = "logits"):
        return NemotronHForCausalLM.compute_logits(self, hidden, key) * float(self.cfg.falcon_h1.get("lm_head_multiplier", 1.0))

    def make_mamba_pool(self, slots: int, device) -> Optional[MambaPool]:
        mx = self.model.all_layers[0].mamba
        return MambaPool(self.model.num_mamba, slots, mx.conv_dim, mx.k, mx.H, mx.P, mx.N, _DT[self.cfg.dtype], device)

    def load_weights(self, weights):
        f = self.cfg.falcon_h1
        m = self.cfg.mamba
        di, gn, H = m["num_heads"] * m["head_dim"], m["n_groups"] * m["d_state"], m["num_heads"]
        zx = f.get("ssm_multipliers") or [1.0] * 5
        mup = torch.cat([torch.full((n,), float(v)) for n, v in zip((di, di, gn, gn, H), zx)]) * float(f.get("ssm_in_multiplier", 1.0))
        gate_m, down_m = f.get("mlp_multipliers") or (1.0, 1.0)
        scale = {".self_attn.q_proj.": float(f.get("attention_in_multiplier", 1.0)), ".self_attn.v_proj.": float(f.get("attention_in_multiplier", 1.0)),
                 ".self_attn.k_proj.": float(f.get("attention_in_multiplier", 1.0)) * float(f.get("key_multiplier", 1.0)),
                 ".self_attn.o_proj.": float(f.get("attention_out_multiplier", 1.0)), ".mamba.out_proj.": float(f.get("ssm_out_multiplier", 1.0)),
                 ".feed_forward.gate_proj.": float(gate_m), ".feed_forward.down_proj.": float(down_m)}

        def fold():
            for name, w in weights:
                if name.endswith(".mamba.in_proj.weight"):
                    w = w * mup[:, None].to(w.dtype)
                else:
                    for k, v in scale.items():
                        if k in name and v != 1.0:
                            w = w * v
                            break
                yield name, w
        return super().load_weights(fold())
