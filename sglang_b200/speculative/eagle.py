"""EAGLE-style feature-level speculative decoding (chain drafts, top-1).

Parity: reference ``python/sglang/srt/speculative/eagle_worker_v2.py:1050`` (draft extend after prefill / after verify, multi-step
draft), ``models/llama_eagle.py`` (draft = ``fc(2H -> H)`` over ``[embed(next token) ; target feature]`` + ONE decoder layer, sharing
the target's embedding and LM head), ``eagle_info.py`` (features captured from the verify pass).

How it runs here:
* the target's final hidden states are tapped (``ModelRunner.hidden_tap`` for prefill chunks, ``last_verify_hidden`` for verify
  passes) and the newest ``window`` (token position, feature) pairs are kept per request;
* ``propose`` re-runs the one-layer draft over that window into a scratch KV pool (no second page table: the draft is one layer, the
  window is short), then rolls forward ``num_draft`` steps feeding its OWN output feature back in (EAGLE's autoregression in feature
  space) and taking the arg-max token at each step;
* verification is the shared ragged-extend path of ``BaseSpecWorker`` (chain drafts => plain causal attention on the tcgen05 kernel).
Tree drafts (topk > 1) need custom-mask attention; ``speculative_eagle_topk`` > 1 is accepted and clamped to the chain.
"""
from __future__ import annotations

import logging
from typing import Any, List, Optional, Tuple

import numpy as np
import torch
from torch import nn

from ..configs.model_config import ModelConfig, load_model_config
from ..managers.schedule_batch import ForwardMode, ScheduleBatch
from ..mem_cache.memory_pool import KVPool
from ..model_executor.forward_batch import ForwardBatch
from ..models.llama import _DT, LlamaDecoderLayer
from ..layers.common import RMSNorm, RotaryEmbedding
from .spec_worker import BaseSpecWorker

logger = logging.getLogger(__name__)


class EagleDraftModel(nn.Module):
    def __init__(self, cfg: ModelConfig, device: Any):
        super().__init__()
        dtype = _DT[cfg.dtype]
        H = cfg.hidden_size
        self.fc = nn.Linear(2 * H, H, bias=False, dtype=dtype)
        rope = RotaryEmbedding(cfg.head_dim, cfg.max_position_embeddings, cfg.rope_theta, cfg.rope_scaling, cfg.rope_is_neox,
                               cfg.partial_rotary_factor, device)
        self.layer = LlamaDecoderLayer(cfg, 0, dtype, None, rope, "model.layers.0")
        self.layer.input_layernorm = _Identity()        # EAGLE's first (only) layer has no input norm: fc output feeds attention

    def forward(self, fb: ForwardBatch, tok_embeds: torch.Tensor, feats: torch.Tensor) -> torch.Tensor:
        x = self.fc(torch.cat([tok_embeds, feats.to(tok_embeds.dtype)], dim=-1))
        h, res = self.layer(x, None, fb)
        h = h.resolve() if hasattr(h, "resolve") else h
        return h + res                                  # predicted next-position feature

    def load_weights(self, weights) -> List[str]:
        params = dict(self.named_parameters())
        loaded = []
        remap = [("model.fc.", "fc."), ("fc.", "fc."), ("model.layers.0.", "layer."), ("layers.0.", "layer.")]
        stacked = [(".qkv_proj", ".q_proj", "q"), (".qkv_proj", ".k_proj", "k"), (".qkv_proj", ".v_proj", "v"),
                   (".gate_up_proj", ".gate_proj", 0), (".gate_up_proj", ".up_proj", 1)]
        mods = dict(self.named_modules())
        for name, w in weights:
            for a, b in remap:
                if name.startswith(a):
                    name = b + name[len(a):]
                    break
            target, shard = name, None
            for merged, orig, sid in stacked:
                if orig in name:
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


class _Identity(nn.Module):
    def forward(self, x, residual=None):
        return x if residual is None else (x, residual)


class EagleSpecWorker(BaseSpecWorker):
    def __init__(self, args: Any, runner: Any, scheduler: Any, window: int = 48):
        super().__init__(args, runner, scheduler)
        self.topk = max(1, int(args.speculative_eagle_topk or 1))
        self.window = window
        tcfg = runner.cfg
        path = args.speculative_draft_model_path
        cfg = load_model_config(path) if path and not path.startswith("dummy:") else tcfg
        self.cfg = cfg
        self.draft = self.make_draft(cfg, runner).to(runner.device)
        if path and not path.startswith("dummy:"):
            from ..model_loader.loader import iter_safetensors

            self.draft.load_weights(iter_safetensors(path))
        else:
            g = torch.Generator(device="cpu").manual_seed(args.random_seed + 29)
            for p in self.draft.parameters():
                if p.dim() > 1:
                    p.data.copy_((torch.randn(p.shape, generator=g) * 0.02).to(p.dtype))
        a0 = self.draft.layer.self_attn.attn
        self.ps = 16
        self.pool = KVPool(2 + (window + 2 * self.num_draft + self.topk * max(1, args.speculative_num_steps) + 32) // self.ps, self.ps, 1, a0.num_kv_heads,
                           getattr(a0, "kv_head_dim", None) or cfg.head_dim, runner.dtype, str(runner.device))
        # the feature window lives ON the request object (``req._eagle``: (position of the first kept row, features [n, H])) so it
        # is dropped with the request — no registry to clean up on finish / abort / retraction
        runner.hidden_tap = self._tap_prefill

    def make_draft(self, cfg: ModelConfig, runner: Any) -> nn.Module:
        return EagleDraftModel(cfg, runner.device)

    # ---- feature bookkeeping
    def _append(self, req: Any, start_pos: int, rows: torch.Tensor) -> None:
        cur = getattr(req, "_eagle", None)
        if cur is not None and cur[0] + cur[1].shape[0] == start_pos:
            first, buf = cur
            buf = torch.cat([buf, rows], 0)
        else:
            first, buf = start_pos, rows          # gap (cache hit / retraction): restart the window
        if buf.shape[0] > self.window:
            first += buf.shape[0] - self.window
            buf = buf[-self.window:]
        req._eagle = (first, buf.detach())

    def _tap_prefill(self, sb: ScheduleBatch, hidden: torch.Tensor) -> None:
        cu = np.concatenate([[0], np.cumsum(sb.q_lens)])
        for i, r in enumerate(sb.reqs):
            lo, hi = int(cu[i]), int(cu[i + 1])
            keep = min(hi - lo, self.window)
            self._append(r, int(sb.prefix_lens[i]) + (hi - lo) - keep, hidden[hi - keep:hi])

    def on_verified_hidden(self, req: Any, start_pos: int, hidden: torch.Tensor) -> None:
        self._append(req, start_pos, hidden)

    # ---- drafting
    @torch.no_grad()
    def propose(self, req: Any, max_draft: int) -> List[int]:
        st = getattr(req, "_eagle", None)
        if st is None or max_draft <= 0:
            return []
        first, feats = st
        toks = req.origin_input_ids + req.output_ids
        L = len(toks) - 1                      # position of the newest token (no target feature yet)
        n = feats.shape[0]
        if first + n != L or n == 0:           # features must end right before the newest token
            return []
        dev = self.runner.device
        target = self.runner.model
        embed = target.model.embed_tokens
        out: List[int] = []
        npages = self.pool.num_pages - 1
        pt = torch.arange(1, npages + 1, dtype=torch.int32, device=dev).view(1, -1)
        pt = torch.cat([torch.zeros_like(pt), pt], 0)
        # window pass: inputs (token p+1, feature p) for p = first .. L-1, draft KV rows 0 .. n-1
        ids = torch.tensor(toks[first + 1: L + 1], dtype=torch.int32, device=dev)
        f_in = feats
        base = 0
        for step in range(max_draft):
            m = ids.shape[0]
            rel = torch.arange(base, base + m, dtype=torch.int32, device=dev)
            fb = ForwardBatch(mode=ForwardMode.EXTEND, batch_size=1, num_tokens=m, input_ids=ids, positions=rel + first,
                              out_cache_loc=(rel + self.ps).to(torch.int32), req_slots=torch.ones(1, dtype=torch.int32, device=dev),
                              seq_lens=torch.tensor([base + m], dtype=torch.int32, device=dev),
                              cu_q=torch.tensor([0, m], dtype=torch.int32, device=dev), q_lens=torch.tensor([m], dtype=torch.int32, device=dev),
                              last_idx=None, sample_slots=None, max_q_len=m, max_kv_len=base + m, page_table=pt, page_size=self.ps,
                              kv_pool=self.pool)
            h = self.draft(fb, embed(ids), f_in)
            last = h[-1:]
            logits = target.compute_logits(last)       # features are post-final-norm hidden states (what the LM head consumes)
            t = int(logits.argmax(-1))
            out.append(t)
            base += m
            ids = torch.tensor([t], dtype=torch.int32, device=dev)
            f_in = last                         # autoregression in feature space
            if base + 1 >= self.window + 2 * self.num_draft + 16:
                break
        return out

    @torch.no_grad()
    def propose_tree(self, req: Any, max_nodes: int):
        """EAGLE-2 style token tree (``--speculative-eagle-topk k > 1``): the window pass scores the children of the last token; at every
        further depth the ``k`` best frontier nodes (by cumulative log-probability) are expanded in ONE draft forward whose tree mask lets
        a node see the window and its own ancestors; the ``max_nodes`` best candidates overall form the tree.  Cumulative scores never
        increase along a path, so the best-N set is closed under "parent of" (reference eagle_utils.py build_tree / organize_draft_results)."""
        if self.topk <= 1:
            return super().propose_tree(req, max_nodes)
        st = getattr(req, "_eagle", None)
        if st is None or max_nodes <= 0:
            return [], []
        first, feats = st
        toks = req.origin_input_ids + req.output_ids
        L = len(toks) - 1
        n = feats.shape[0]
        if first + n != L or n == 0:
            return [], []
        dev, target, k = self.runner.device, self.runner.model, self.topk
        embed = target.model.embed_tokens
        npages = self.pool.num_pages - 1
        pt = torch.arange(1, npages + 1, dtype=torch.int32, device=dev).view(1, -1)
        pt = torch.cat([torch.zeros_like(pt), pt], 0)
        i32 = lambda v: torch.tensor(v, dtype=torch.int32, device=dev)

        def run(ids, f_in, positions, base, mask):
            m = ids.shape[0]
            rel = torch.arange(base, base + m, dtype=torch.int32, device=dev)
            fb = ForwardBatch(mode=ForwardMode.EXTEND, batch_size=1, num_tokens=m, input_ids=ids, positions=positions,
                              out_cache_loc=(rel + self.ps).to(torch.int32), req_slots=torch.ones(1, dtype=torch.int32, device=dev),
                              seq_lens=i32([base + m]), cu_q=i32([0, m]), q_lens=i32([m]), last_idx=None, sample_slots=None, max_q_len=m,
                              max_kv_len=base + m, page_table=pt, page_size=self.ps, kv_pool=self.pool, custom_mask=mask)
            return self.draft(fb, embed(ids), f_in)

        ids = i32(toks[first + 1: L + 1])
        h = run(ids, feats, torch.arange(n, dtype=torch.int32, device=dev) + first, 0, None)
        # candidate pool: (score, token, parent candidate index or -1, depth, feature row to feed when expanded)
        cand_tok: List[int] = []
        cand_par: List[int] = []
        cand_score: List[float] = []
        cand_depth: List[int] = []
        row_of: dict = {}                 # candidate index -> draft KV row (only expanded candidates have one)
        feat_of: dict = {}                # candidate index -> the parent's output feature (input of this candidate's expansion)

        def add_children(parent: int, feat: torch.Tensor, score: float, depth: int) -> List[int]:
            lp = torch.log_softmax(target.compute_logits(feat).float(), -1)[0]
            v, t = torch.topk(lp, k)
            out = []
            for s_, t_ in zip(v.tolist(), t.tolist()):
                cand_tok.append(int(t_))
                cand_par.append(parent)
                cand_score.append(score + float(s_))
                cand_depth.append(depth)
                feat_of[len(cand_tok) - 1] = feat
                out.append(len(cand_tok) - 1)
            return out

        frontier = add_children(-1, h[-1:], 0.0, 1)
        base = n
        steps = max(1, int(self.args.speculative_num_steps))
        cap = self.pool.num_slots - self.ps - 1
        for _ in range(steps - 1):
            if len(cand_tok) >= 4 * max_nodes or base + len(frontier) >= cap:
                break
            m = len(frontier)
            mask = torch.zeros(m, base + m, dtype=torch.bool, device=dev)
            mask[:, :n] = True
            for j, c in enumerate(frontier):
                a = cand_par[c]
                while a >= 0:
                    mask[j, row_of[a]] = True
                    a = cand_par[a]
                mask[j, base + j] = True
                row_of[c] = base + j
            hx = run(i32([cand_tok[c] for c in frontier]), torch.cat([feat_of[c] for c in frontier], 0),
                     i32([first + n - 1 + cand_depth[c] for c in frontier]), base, mask.reshape(-1))
            base += m
            new = []
            for j, c in enumerate(frontier):
                new += add_children(c, hx[j: j + 1], cand_score[c], cand_depth[c] + 1)
            frontier = sorted(new, key=lambda c: (-cand_score[c], c))[:k]
        order = sorted(range(len(cand_tok)), key=lambda c: (-cand_score[c], c))[:max_nodes]
        keep = sorted(order)                                  # creation order is topological
        idx = {c: j for j, c in enumerate(keep)}
        assert all(cand_par[c] < 0 or cand_par[c] in idx for c in keep)
        return [cand_tok[c] for c in keep], [idx[cand_par[c]] if cand_par[c] >= 0 else -1 for c in keep]


# ================================================================================================ EAGLE-3
class Eagle3DraftModel(nn.Module):
    """EAGLE-3 draft (reference ``models/llama_eagle3.py``): ``fc`` fuses the hidden states of several target layers (low / mid / high) into one
    feature; ONE decoder layer whose attention reads ``[input_layernorm(embed(token)) ; hidden_norm(feature)]`` (2H wide); own final norm and
    LM head over a reduced draft vocabulary with a ``d2t`` offset table back to target ids.  Later draft steps feed the layer's own
    pre-norm output back in as the feature."""

    def __init__(self, cfg: ModelConfig, device: Any, n_aux: int = 3, target_hidden: Optional[int] = None, draft_vocab: Optional[int] = None):
        super().__init__()
        from ..layers.linear import QKVParallelLinear

        dtype = _DT[cfg.dtype]
        H = cfg.hidden_size
        self.n_aux, self.Ht = n_aux, target_hidden or H
        self.fc = nn.Linear(self.Ht * n_aux, H, bias=False, dtype=dtype)
        rope = RotaryEmbedding(cfg.head_dim, cfg.max_position_embeddings, cfg.rope_theta, cfg.rope_scaling, cfg.rope_is_neox,
                               cfg.partial_rotary_factor, device)
        self.layer = LlamaDecoderLayer(cfg, 0, dtype, None, rope, "model.layers.0")
        sa = self.layer.self_attn
        sa.qkv_proj = QKVParallelLinear(2 * H, cfg.head_dim, cfg.num_attention_heads, cfg.num_key_value_heads, cfg.attention_bias, dtype, None,
                                        "model.layers.0.self_attn.qkv_proj", tp_group=sa.qkv_proj.tp)
        self.layer.hidden_norm = RMSNorm(H, cfg.rms_norm_eps, dtype)
        self.norm = RMSNorm(H, cfg.rms_norm_eps, dtype)
        self.draft_vocab = draft_vocab or cfg.vocab_size
        self.lm_head = nn.Linear(H, self.draft_vocab, bias=False, dtype=dtype)
        self.register_buffer("d2t", torch.zeros(self.draft_vocab, dtype=torch.long), persistent=False)

    def fuse(self, aux: torch.Tensor) -> torch.Tensor:
        return self.fc(aux.to(self.fc.weight.dtype))

    def forward(self, fb: ForwardBatch, tok_embeds: torch.Tensor, hidden: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        L = self.layer
        residual = hidden
        x = torch.cat([L.input_layernorm(tok_embeds), L.hidden_norm(hidden)], dim=-1)
        a = L.self_attn(x, fb)
        a = a.resolve() if hasattr(a, "resolve") else a
        h, residual = L.post_attention_layernorm(a, residual.clone())
        m = L.mlp(h)
        m = m.resolve() if hasattr(m, "resolve") else m
        normed, pre = self.norm(m, residual)
        return normed, pre                  # (what the draft LM head reads, the feature the next draft step consumes)

    def to_target_ids(self, draft_ids: torch.Tensor) -> torch.Tensor:
        return draft_ids + self.d2t[draft_ids]

    def load_weights(self, weights) -> List[str]:
        params = dict(self.named_parameters())
        loaded = []
        stacked = [(".qkv_proj", ".q_proj", "q"), (".qkv_proj", ".k_proj", "k"), (".qkv_proj", ".v_proj", "v"),
                   (".gate_up_proj", ".gate_proj", 0), (".gate_up_proj", ".up_proj", 1)]
        mods = dict(self.named_modules())
        for name, w in weights:
            if name == "d2t":
                self.d2t.copy_(w.to(torch.long)[: self.d2t.shape[0]])
                loaded.append(name)
                continue
            n = name
            for a, b in (("model.midlayer.", "layer."), ("midlayer.", "layer."), ("model.layers.0.", "layer."), ("layers.0.", "layer."),
                         ("model.fc.", "fc."), ("model.norm.", "norm.")):
                if n.startswith(a):
                    n = b + n[len(a):]
                    break
            target, shard = n, None
            for merged, orig, sid in stacked:
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


def default_aux_layers(num_layers: int) -> List[int]:
    """the three taps EAGLE-3 trains with: low, middle, high (reference ``set_eagle3_layers_to_capture``: [2, L // 2, L - 3])"""
    return sorted({min(num_layers - 1, max(0, x)) for x in (2, num_layers // 2, num_layers - 3)})


class Eagle3SpecWorker(EagleSpecWorker):
    """features = concatenated target hidden states of ``aux_layers`` (captured inside the target's forward); chain drafts"""

    def __init__(self, args: Any, runner: Any, scheduler: Any, window: int = 48):
        BaseSpecWorker.__init__(self, args, runner, scheduler)
        self.topk, self.window = 1, window
        tcfg = runner.cfg
        path = args.speculative_draft_model_path
        real = bool(path) and not path.startswith("dummy:")
        cfg = load_model_config(path) if real else tcfg
        extra = {}
        if real:
            import json
            import os

            with open(os.path.join(path, "config.json")) as f:
                extra = json.load(f)
        ec = extra.get("eagle_config") or {}
        self.aux_layers = list(ec.get("eagle_aux_hidden_state_layer_ids") or default_aux_layers(tcfg.num_hidden_layers))
        self.cfg = cfg
        self.draft = Eagle3DraftModel(cfg, runner.device, len(self.aux_layers), extra.get("target_hidden_size") or tcfg.hidden_size,
                                      extra.get("draft_vocab_size")).to(runner.device)
        if real:
            from ..model_loader.loader import iter_safetensors

            got = self.draft.load_weights(iter_safetensors(path))
            if not any(n.startswith("lm_head") for n in got):      # checkpoints without a head share the target's
                head = runner.model.lm_head if runner.model.lm_head is not None else runner.model.model.embed_tokens
                self.draft.lm_head.weight.data.copy_(head.weight[: self.draft.draft_vocab])
        else:
            g = torch.Generator(device="cpu").manual_seed(args.random_seed + 31)
            for p in self.draft.parameters():
                if p.dim() > 1:
                    p.data.copy_((torch.randn(p.shape, generator=g) * 0.02).to(p.dtype))
        a0 = self.draft.layer.self_attn.attn
        self.ps = 16
        self.pool = KVPool(2 + (window + 2 * self.num_draft + 32) // self.ps, self.ps, 1, a0.num_kv_heads, cfg.head_dim, runner.dtype, str(runner.device))
        runner.aux_layer_ids = set(self.aux_layers)
        runner.hidden_tap = self._tap_prefill

    @torch.no_grad()
    def propose(self, req: Any, max_draft: int) -> List[int]:
        st = getattr(req, "_eagle", None)
        if st is None or max_draft <= 0:
            return []
        first, feats = st
        toks = req.origin_input_ids + req.output_ids
        L = len(toks) - 1
        n = feats.shape[0]
        if first + n != L or n == 0:
            return []
        dev = self.runner.device
        embed = self.runner.model.model.embed_tokens
        npages = self.pool.num_pages - 1
        pt = torch.arange(1, npages + 1, dtype=torch.int32, device=dev).view(1, -1)
        pt = torch.cat([torch.zeros_like(pt), pt], 0)
        i32 = lambda v: torch.tensor(v, dtype=torch.int32, device=dev)
        ids = i32(toks[first + 1: L + 1])
        hidden = self.draft.fuse(feats)                      # [n, H]: one fused feature per window position
        out: List[int] = []
        base = 0
        for _ in range(max_draft):
            m = ids.shape[0]
            rel = torch.arange(base, base + m, dtype=torch.int32, device=dev)
            fb = ForwardBatch(mode=ForwardMode.EXTEND, batch_size=1, num_tokens=m, input_ids=ids, positions=rel + first,
                              out_cache_loc=(rel + self.ps).to(torch.int32), req_slots=torch.ones(1, dtype=torch.int32, device=dev),
                              seq_lens=i32([base + m]), cu_q=i32([0, m]), q_lens=i32([m]), last_idx=None, sample_slots=None, max_q_len=m,
                              max_kv_len=base + m, page_table=pt, page_size=self.ps, kv_pool=self.pool)
            normed, pre = self.draft(fb, embed(ids), hidden)
            t = int(self.draft.to_target_ids(self.draft.lm_head(normed[-1:]).argmax(-1))[0])
            out.append(t)
            base += m
            ids, hidden = i32([t]), pre[-1:]                 # autoregression on the draft's own (pre-norm) feature
            if base + 1 >= self.window + 2 * self.num_draft + 16:
                break
        return out

    def propose_tree(self, req: Any, max_nodes: int):
        return BaseSpecWorker.propose_tree(self, req, max_nodes)


# ================================================================================================ MTP / NextN (DeepSeek-V3 style)
class MtpDraftModel(nn.Module):
    """multi-token-prediction module shipped inside the target checkpoint as layer ``num_hidden_layers`` (reference ``models/deepseek_nextn.py``):
    ``eh_proj([enorm(embed(token)) ; hnorm(previous hidden)])`` -> one full decoder layer of the target's own kind (MLA + MoE for DeepSeek) ->
    ``shared_head.norm``; the LM head is the target's.  Same driver as EAGLE: autoregression on its own output hidden state."""

    def __init__(self, cfg: ModelConfig, device: Any, layer_cls: Any):
        super().__init__()
        dtype = _DT[cfg.dtype]
        H = cfg.hidden_size
        self.enorm, self.hnorm = RMSNorm(H, cfg.rms_norm_eps, dtype), RMSNorm(H, cfg.rms_norm_eps, dtype)
        self.eh_proj = nn.Linear(2 * H, H, bias=False, dtype=dtype)
        rope = RotaryEmbedding(cfg.head_dim, cfg.max_position_embeddings, cfg.rope_theta, cfg.rope_scaling, cfg.rope_is_neox,
                               cfg.partial_rotary_factor, device)
        # the MTP layer sits AFTER the dense prefix: build it with the index of the last layer so MoE families make it sparse
        self.layer = layer_cls(cfg, 0, dtype, None, rope, f"model.layers.{cfg.num_hidden_layers - 1}")
        self.norm = RMSNorm(H, cfg.rms_norm_eps, dtype)
        self.src_layer = cfg.num_hidden_layers

    def forward(self, fb: ForwardBatch, tok_embeds: torch.Tensor, feats: torch.Tensor) -> torch.Tensor:
        x = self.eh_proj(torch.cat([self.enorm(tok_embeds), self.hnorm(feats.to(tok_embeds.dtype))], dim=-1))
        h, res = self.layer(x, None, fb)
        h = h.resolve() if hasattr(h, "resolve") else h
        return self.norm(h) if res is None else self.norm(h, res)[0]

    def load_weights(self, weights) -> List[str]:
        """picks ``model.layers.{L}.*`` out of the TARGET checkpoint"""
        pre = f"model.layers.{self.src_layer}."
        params = dict(self.named_parameters())
        mods = dict(self.named_modules())
        stacked = [(".qkv_proj", ".q_proj", "q"), (".qkv_proj", ".k_proj", "k"), (".qkv_proj", ".v_proj", "v"),
                   (".gate_up_proj", ".gate_proj", 0), (".gate_up_proj", ".up_proj", 1)]
        shard_of = {"gate_proj": "w1", "up_proj": "w3", "down_proj": "w2"}
        loaded = []
        for name, w in weights:
            if not name.startswith(pre):
                continue
            n = name[len(pre):]
            if n.startswith(("embed_tokens.", "shared_head.head.")):
                continue                        # shared with the target
            n = n.replace("shared_head.norm.", "norm.") if n.startswith("shared_head.norm.") else (n if n.startswith(("enorm.", "hnorm.", "eh_proj.")) else "layer." + n)
            if ".mlp.experts." in n and n.split(".experts.")[1].split(".")[0].isdigit():
                ex = mods.get("layer.mlp.experts")
                if ex is not None:
                    eid, proj = int(n.split(".experts.")[1].split(".")[0]), n.split(".")[-2]
                    ex.weight_loader(ex.w2 if proj == "down_proj" else ex.w13, w, eid, shard_of[proj])
                    loaded.append(name)
                continue
            if n.endswith("mlp.gate.e_score_correction_bias"):
                g = mods.get("layer.mlp.gate")
                if g is not None and getattr(g, "e_score_correction_bias", None) is not None:
                    g.e_score_correction_bias.data.copy_(w.float())
                    loaded.append(name)
                continue
            target, shard = n, None
            for merged, orig, sid in stacked:
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


class MtpSpecWorker(EagleSpecWorker):
    """``--speculative-algorithm NEXTN`` (alias MTP): the draft is the target checkpoint's own MTP layer (``--speculative-draft-model-path``
    defaults to the target path)"""

    def __init__(self, args: Any, runner: Any, scheduler: Any, window: int = 48):
        if not args.speculative_draft_model_path and not str(args.model_path).startswith("dummy:"):
            args.speculative_draft_model_path = args.model_path
        super().__init__(args, runner, scheduler, window)

    def make_draft(self, cfg: ModelConfig, runner: Any) -> nn.Module:
        layer_cls = type(runner.model.model.layers[-1])
        return MtpDraftModel(runner.cfg, runner.device, layer_cls)
