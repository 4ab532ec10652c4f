"""Speculative decoding workers.

Parity: reference ``python/sglang/srt/speculative/`` (algorithm registry ``spec_info.py:31-118``; ``NGRAMWorker``
``ngram_worker.py:71`` backed by the C++ corpus ``kernels/jit/csrc/ngram_corpus``; ``StandaloneWorkerV2``; greedy /
rejection verification ``kernels/aot/csrc/speculative/eagle_utils.cu``, ``speculative_sampling.cu``; accept-length
statistics in ``meta_info``).

Design: a decode step of B requests becomes ONE ragged *extend* of sum(k_i + 1) tokens — request i feeds
``[last_token, d_1 .. d_k]`` at positions ``p .. p+k`` over its cached prefix — which is exactly the shape the tcgen05
prefill kernel already handles (causal over prefix + new tokens), so chain drafts need no custom-mask attention.
Verification samples every position with the request's own sampling parameters and accepts draft ``d_j`` iff the target
sample at position j-1 equals it; because the proposal is deterministic this emits tokens distributed exactly as the
target model would (greedy => identical output to non-speculative decoding).
"""
from __future__ import annotations

import logging
from typing import Any, List, Optional

import numpy as np
import torch

from .. import ops
from ..managers.schedule_batch import ForwardMode, ScheduleBatch

logger = logging.getLogger(__name__)


class SpecResult:
    def __init__(self, B: int):
        self.accept_lens = np.ones(B, dtype=np.int64)           # tokens emitted per request (accepted drafts + 1)
        self.accepted_tokens: List[List[int]] = [[] for _ in range(B)]
        self.next_token_ids = np.full(B, -1, dtype=np.int64)
        self.next_token_logprobs = None
        self.top_logprobs = None
        self.hidden_states = None

    def wait(self) -> None:
        pass


class BaseSpecWorker:
    def __init__(self, args: Any, runner: Any, scheduler: Any):
        self.args, self.runner, self.scheduler = args, runner, scheduler
        self.num_draft = max(1, args.speculative_num_draft_tokens)
        self.total_drafted = 0
        self.total_accepted = 0
        self.total_verify = 0

    def propose(self, req: Any, max_draft: int) -> List[int]:
        raise NotImplementedError

    def on_tokens_accepted(self, req: Any, new_tokens: List[int]) -> None:
        pass

    def on_verified_hidden(self, req: Any, start_pos: int, hidden: torch.Tensor) -> None:
        pass

    def forward_batch_speculative(self, batch: ScheduleBatch) -> SpecResult:
        """``batch``: decode snapshot (one slot already allocated at position p = seq_len-1 for every request)."""
        sch, kv, runner = self.scheduler, self.scheduler.kv, self.runner
        B = len(batch.reqs)
        ids, pos, loc, qlens, plens = [], [], [], [], []
        drafts: List[List[int]] = []
        for i, r in enumerate(batch.reqs):
            p = int(batch.positions[i])
            room = min(self.num_draft, r.sampling_params.max_new_tokens - len(r.output_ids) - 1,
                       sch.max_context_len - p - 2)
            d = self.propose(r, room) if room > 0 else []
            if r.grammar is not None:
                d = []      # grammar-constrained requests verify one token at a time (masks depend on accepted text)
            if d and not kv.alloc_for_extend(r, p + len(d) + 1):
                d = []      # no KV room for the draft: fall back to plain decode for this request
            drafts.append(d)
            last = r.output_ids[-1] if r.output_ids else r.origin_input_ids[-1]
            ids.append(np.asarray([last] + d, dtype=np.int32))
            pos.append(np.arange(p, p + len(d) + 1, dtype=np.int32))
            loc.append(kv.slots_for(r, p, p + len(d) + 1))
            qlens.append(len(d) + 1)
            plens.append(p)
        vb = ScheduleBatch(reqs=batch.reqs, forward_mode=ForwardMode.TARGET_VERIFY)
        vb.input_ids = np.concatenate(ids)
        vb.positions = np.concatenate(pos)
        vb.out_cache_loc = np.concatenate(loc)
        vb.q_lens = np.asarray(qlens, dtype=np.int32)
        vb.prefix_lens = np.asarray(plens, dtype=np.int32)
        vb.req_slots = batch.req_slots
        vb.seq_lens = (vb.prefix_lens + vb.q_lens).astype(np.int32)
        vb.sample_mask = np.ones(B, dtype=bool)
        d0 = (batch.pt_rows, batch.pt_cols, batch.pt_vals)
        d1 = kv.take_deltas()
        vb.pt_rows, vb.pt_cols, vb.pt_vals = (np.concatenate([a, b]) for a, b in zip(d0, d1))
        vb.spec_info = {"verify_all_positions": True}
        vb.has_grammar = batch.has_grammar
        toks_all = runner.forward_verify(vb)          # np.int64 [T]: target sample at every fed position
        res = SpecResult(B)
        cu = np.concatenate([[0], np.cumsum(vb.q_lens)])
        vh = runner.last_verify_hidden
        for i, r in enumerate(batch.reqs):
            t = toks_all[cu[i]:cu[i + 1]]
            d = drafts[i]
            a = 0
            while a < len(d) and int(t[a]) == d[a]:
                a += 1
            out = [int(x) for x in t[: a + 1]]    # a accepted drafts (== target samples) + the bonus/correction token
            res.accept_lens[i] = len(out)
            res.accepted_tokens[i] = out
            res.next_token_ids[i] = out[-1]
            r.kv_len += a                          # KV of last_token and the a accepted drafts is now valid
            batch.seq_lens[i] += a
            kv.trim(r, r.kv_len)                   # rejected drafts' pages go back (next decode re-allocates on demand)
            self.total_drafted += len(d)
            self.total_accepted += a
            self.total_verify += 1
            if vh is not None:                     # target features of the positions that stay (last token + accepted drafts)
                self.on_verified_hidden(r, plens[i], vh[cu[i]:cu[i] + a + 1])
            self.on_tokens_accepted(r, out)
        # keep the device-side "future map" coherent for anyone who reads it
        slots = torch.from_numpy(np.asarray(batch.req_slots, dtype=np.int64)).to(runner.device)
        runner.next_token_buf[slots] = torch.from_numpy(res.next_token_ids.astype(np.int32)).to(runner.device)
        # running-batch vectors advance by the accepted drafts as well
        rb = sch.running_batch
        idx = {id(r): j for j, r in enumerate(rb.reqs)}
        for i, r in enumerate(batch.reqs):
            j = idx.get(id(r))
            if j is not None:
                rb.seq_lens[j] = r.kv_len
        return res

    def accept_rate(self) -> float:
        return self.total_accepted / max(1, self.total_drafted)

    def accept_length(self) -> float:
        return 1.0 + self.total_accepted / max(1, self.total_verify)


class NgramSpecWorker(BaseSpecWorker):
    """drafts from an n-gram trie over everything this server has seen (prompts + generations)"""

    def __init__(self, args: Any, runner: Any, scheduler: Any, max_depth: int = 12):
        super().__init__(args, runner, scheduler)
        from .. import sglb_rt

        self.corpus = sglb_rt.NgramCorpus(max_depth, 1 << 22)
        self.max_depth = max_depth
        self._seen_prompt = set()

    def propose(self, req: Any, max_draft: int) -> List[int]:
        if id(req) not in self._seen_prompt:
            self._seen_prompt.add(id(req))
            self.corpus.insert(np.asarray(req.origin_input_ids, dtype=np.int32))
        ctx = (req.origin_input_ids + req.output_ids)[-self.max_depth:]
        return self.corpus.match_chain(np.asarray(ctx, dtype=np.int32), max_draft, 2).tolist()

    def on_tokens_accepted(self, req: Any, new_tokens: List[int]) -> None:
        tail = (req.origin_input_ids + req.output_ids)[-(self.max_depth - 1):] + list(new_tokens)
        self.corpus.insert(np.asarray(tail, dtype=np.int32))
        if req.finished():
            self._seen_prompt.discard(id(req))


class StandaloneSpecWorker(BaseSpecWorker):
    """a small draft LM proposes the chain autoregressively (reference ``StandaloneWorkerV2`` / EAGLE chain mode with
    topk=1).  The draft model keeps NO KV cache of its own: it re-reads a bounded window of recent tokens, which is cheap
    for draft-sized models and keeps the worker free of a second page table."""

    def __init__(self, args: Any, runner: Any, scheduler: Any, window: int = 64):
        super().__init__(args, runner, scheduler)
        from ..configs.model_config import load_model_config
        from ..model_loader.loader import load_model

        path = args.speculative_draft_model_path or args.model_path
        self.cfg = load_model_config(path)
        if runner.dtype == torch.float32:
            self.cfg.dtype = "float32"
        self.window = window
        self.draft = load_model(self.cfg, path, "dummy" if path.startswith("dummy:") else "safetensors", runner.device, None,
                                args.random_seed + 17)
        from ..mem_cache.memory_pool import KVPool

        a0 = self.draft.model.layers[0].self_attn.attn
        self.ps = 16
        self.pool = KVPool(2 + (window + 32) // self.ps, self.ps, len(self.draft.model.layers), a0.num_kv_heads, self.cfg.head_dim,
                           runner.dtype, str(runner.device))

    @torch.no_grad()
    def propose(self, req: Any, max_draft: int) -> List[int]:
        from ..model_executor.forward_batch import ForwardBatch, SamplingTensors

        dev = self.runner.device
        ctx = (req.origin_input_ids + req.output_ids)[-self.window:]
        out: List[int] = []
        npages = self.pool.num_pages - 1
        pt = torch.arange(1, npages + 1, dtype=torch.int32, device=dev).view(1, -1)
        pt = torch.cat([torch.zeros_like(pt), pt], 0)
        # one prefill over the window, then single-token steps reusing the scratch KV
        toks = list(ctx)
        start = 0
        for step in range(max_draft):
            n = len(toks) - start
            ids = torch.tensor(toks[start:], dtype=torch.int32, device=dev)
            pos = torch.arange(start, len(toks), dtype=torch.int32, device=dev)
            fb = ForwardBatch(mode=ForwardMode.EXTEND, batch_size=1, num_tokens=n, input_ids=ids, positions=pos,
                              out_cache_loc=(pos + self.ps).to(torch.int32), req_slots=torch.ones(1, dtype=torch.int32, device=dev),
                              seq_lens=torch.tensor([len(toks)], dtype=torch.int32, device=dev),
                              cu_q=torch.tensor([0, n], dtype=torch.int32, device=dev), q_lens=torch.tensor([n], dtype=torch.int32, device=dev),
                              last_idx=None, sample_slots=None, max_q_len=n, max_kv_len=len(toks), page_table=pt, page_size=self.ps,
                              kv_pool=self.pool)
            h = self.draft(fb)
            logits = self.draft.compute_logits(h[-1:])
            t = int(logits.argmax(-1))
            out.append(t)
            start = len(toks)
            toks.append(t)
            if len(toks) >= self.window + 32:
                break
        return out


def create_spec_worker(args: Any, runner: Any, scheduler: Any) -> BaseSpecWorker:
    algo = (args.speculative_algorithm or "").upper()
    if algo == "NGRAM":
        return NgramSpecWorker(args, runner, scheduler)
    if algo in ("EAGLE", "EAGLE3"):
        from .eagle import EagleSpecWorker

        return EagleSpecWorker(args, runner, scheduler)
    if algo == "STANDALONE":
        return StandaloneSpecWorker(args, runner, scheduler)
    raise ValueError(f"unknown speculative algorithm {args.speculative_algorithm}")
