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
from typing import Any, Dict, List

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
        self.total_tree_verifies = 0       # verifications that needed a tree mask
        self.total_kv_moves = 0            # KV rows compacted after accepting a non-leading path

    def propose(self, req: Any, max_draft: int) -> List[int]:
        raise NotImplementedError

    def propose_tree(self, req: Any, max_nodes: int):
        """(tokens, parents): a draft token TREE in topological order; ``parents[i]`` is the index of node i's parent among the drafts, -1 for
        children of the last committed token.  Default: the chain from ``propose`` (reference eagle_utils build_tree / ngram tree mask)."""
        d = self.propose(req, max_nodes)
        return d, list(range(-1, len(d) - 1))

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
        trees: List[List[int]] = []
        any_tree = False
        for i, r in enumerate(batch.reqs):
            p = int(batch.positions[i])
            room = min(self.num_draft, r.sampling_params.max_new_tokens - len(r.output_ids) - 1,
                       sch.max_context_len - p - 2)
            d, par = self.propose_tree(r, room) if room > 0 else ([], [])
            if r.grammar is not None:
                d, par = [], []      # grammar-constrained requests verify one token at a time (masks depend on accepted text)
            if d and not kv.alloc_for_extend(r, p + len(d) + 1):
                d, par = [], []      # no KV room for the draft: fall back to plain decode for this request
            d, par = list(d), [int(x) for x in par]
            drafts.append(d)
            trees.append(par)
            depth = [0] * (len(d) + 1)           # node 0 = the last committed token; draft j is node j + 1
            for j, pj in enumerate(par):
                assert -1 <= pj < j, "draft tree must be in topological order"
                depth[j + 1] = depth[pj + 1] + 1
                any_tree |= pj != j - 1
            last = r.output_ids[-1] if r.output_ids else r.origin_input_ids[-1]
            ids.append(np.asarray([last] + d, dtype=np.int32))
            pos.append(p + np.asarray(depth, dtype=np.int32))
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
        tree_anc = None
        if any_tree and max(qlens) <= 32:
            # compact form for the tcgen05 kernel: bit k of anc[node] = draft-region key k is an ancestor of the node (or the node itself)
            anc_all = []
            for i in range(B):
                anc = [1] * qlens[i]
                for j, pj in enumerate(trees[i]):
                    anc[j + 1] = anc[pj + 1] | (1 << (j + 1))
                anc_all.extend(anc)
            tree_anc = np.asarray(anc_all, dtype=np.uint32).view(np.int32)
            vb.spec_info["tree_anc"] = tree_anc
        if any_tree and tree_anc is not None and getattr(runner, "supports_tree_anc", lambda: False)():
            self.total_tree_verifies += 1        # masked tcgen05 verify: no O(q x context) host mask
        elif any_tree:
            self.total_tree_verifies += 1
            masks = []
            for i in range(B):
                ql, p = qlens[i], plens[i]
                m = np.zeros((ql, p + ql), dtype=bool)
                m[:, :p] = True
                m[:, p] = True                        # every node descends from the last committed token
                for j, pj in enumerate(trees[i]):
                    m[j + 1, p: p + ql] = m[pj + 1, p: p + ql]
                    m[j + 1, p + j + 1] = True
                masks.append(m.reshape(-1))
            vb.spec_info["custom_mask"] = np.concatenate(masks)
        vb.has_grammar = batch.has_grammar
        toks_all = runner.forward_verify(vb)          # np.int64 [T]: target sample at every fed node
        res = SpecResult(B)
        cu = np.concatenate([[0], np.cumsum(vb.q_lens)])
        vh = runner.last_verify_hidden
        mv_src, mv_dst = [], []
        for i, r in enumerate(batch.reqs):
            t = toks_all[cu[i]:cu[i + 1]]
            d, par = drafts[i], trees[i]
            kids: Dict[int, List[int]] = {}
            for j, pj in enumerate(par):
                kids.setdefault(pj + 1, []).append(j + 1)
            node, path = 0, []                     # walk: follow the child that equals the target's sample at the current node
            while True:
                nxt = next((c for c in kids.get(node, ()) if d[c - 1] == int(t[node])), None)
                if nxt is None:
                    break
                path.append(nxt)
                node = nxt
            a = len(path)
            out = [d[c - 1] for c in path] + [int(t[node])]     # accepted drafts (== target samples) + the bonus / correction token
            if path != list(range(1, a + 1)):      # accepted nodes are not the first a slots: compact their KV rows
                sl = loc[i]
                mv_src.append(sl[np.asarray(path, dtype=np.int64)])
                mv_dst.append(sl[1: a + 1])
            res.accept_lens[i] = len(out)
            res.accepted_tokens[i] = out
            res.next_token_ids[i] = out[-1]
            r.kv_len += a                          # KV of last_token and the a accepted drafts is now valid
            batch.seq_lens[i] += a
            self.total_drafted += len(d)
            self.total_accepted += a
            self.total_verify += 1
            if vh is not None:                     # target features of the nodes that stay (last token + accepted path)
                rows = torch.as_tensor([0] + path, dtype=torch.long, device=vh.device) + int(cu[i])
                self.on_verified_hidden(r, plens[i], vh[rows])
            self.on_tokens_accepted(r, out)
        if mv_src:
            self.total_kv_moves += int(sum(x.size for x in mv_src))
            dev = runner.device
            runner.kv_pool.move_slots(torch.from_numpy(np.concatenate(mv_src).astype(np.int64)).to(dev),
                                      torch.from_numpy(np.concatenate(mv_dst).astype(np.int64)).to(dev))
        for r in batch.reqs:
            kv.trim(r, r.kv_len)                   # rejected drafts' pages go back (next decode re-allocates on demand)
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
    """drafts from an n-gram index over everything this server has seen (prompts + generations).  ``--speculative-ngram-index trie``: the
    bounded-depth counting trie (chain and tree drafts); ``suffix_automaton``: longest-suffix matches of unbounded length over the whole
    token stream (chain drafts; reference ``ngram_corpus/suffix_automaton.cpp``)."""

    def __init__(self, args: Any, runner: Any, scheduler: Any, max_depth: int = 12):
        super().__init__(args, runner, scheduler)
        from .. import sglb_rt

        self.index = getattr(args, "speculative_ngram_index", "trie") or "trie"
        self.max_depth = max_depth if self.index == "trie" else 64
        self.corpus = sglb_rt.NgramCorpus(max_depth, 1 << 22) if self.index == "trie" else sglb_rt.SuffixAutomaton(1 << 23)
        self._seen_prompt = set()
        self._pending: dict = {}          # suffix automaton: generated tokens not yet in the index, per request

    def _ctx(self, req: Any) -> np.ndarray:
        if self._pending:
            self._flush_finished()
        if id(req) not in self._seen_prompt:
            self._seen_prompt.add(id(req))
            self.corpus.insert(np.asarray(req.origin_input_ids, dtype=np.int32))
        return np.asarray((req.origin_input_ids + req.output_ids)[-self.max_depth:], dtype=np.int32)

    def propose(self, req: Any, max_draft: int) -> List[int]:
        if self.index != "trie":
            return self.corpus.match_chain(self._ctx(req), max_draft, 2)[0].tolist()
        return self.corpus.match_chain(self._ctx(req), max_draft, 2).tolist()

    def propose_tree(self, req: Any, max_nodes: int):
        branch = int(self.args.speculative_eagle_topk or 1)
        if branch <= 1 or self.index != "trie":
            return super().propose_tree(req, max_nodes)
        toks, parents = self.corpus.match_tree(self._ctx(req), max_nodes, branch, 2)      # breadth-first over the trie's continuations
        return np.asarray(toks).tolist(), np.asarray(parents).tolist()

    def on_tokens_accepted(self, req: Any, new_tokens: List[int]) -> None:
        if self.index == "trie":
            tail = (req.origin_input_ids + req.output_ids)[-(self.max_depth - 1):] + list(new_tokens)
            self.corpus.insert(np.asarray(tail, dtype=np.int32))
        else:
            # the automaton indexes ONE token stream, so a request's generation enters it in closed pieces (with a little left context so
            # that matches can run across the piece boundary), not token by token interleaved with other requests
            n = self._pending.get(id(req), (req, 0))[1] + len(new_tokens)
            if n >= 32:
                allt = req.origin_input_ids + req.output_ids + list(new_tokens)          # (called before the tokens are appended to the request)
                self.corpus.insert(np.asarray(allt[-(n + 16):], dtype=np.int32))
                n = 0
            self._pending[id(req)] = (req, n)
            self._flush_finished()
        if req.finished():
            self._seen_prompt.discard(id(req))

    def _flush_finished(self) -> None:
        """index the unindexed tail of requests that have finished since the last call (the finish test runs after ``on_tokens_accepted``)"""
        for k, (r, n) in list(self._pending.items()):
            if r.finished():
                if n > 0:
                    allt = r.origin_input_ids + r.output_ids
                    self.corpus.insert(np.asarray(allt[-(n + 16):], dtype=np.int32))
                self._pending.pop(k, None)
                self._seen_prompt.discard(k)


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
    if getattr(args, "dllm_algorithm", None):
        from ..dllm import DllmWorker

        return DllmWorker(args, runner, scheduler)
    if algo == "NGRAM":
        return NgramSpecWorker(args, runner, scheduler)
    if algo in ("NEXTN", "MTP"):
        from .eagle import MtpSpecWorker

        return MtpSpecWorker(args, runner, scheduler)
    if algo == "EAGLE3":
        from .eagle import Eagle3SpecWorker

        return Eagle3SpecWorker(args, runner, scheduler)
    if algo == "EAGLE":
        from .eagle import EagleSpecWorker

        return EagleSpecWorker(args, runner, scheduler)
    if algo == "STANDALONE":
        return StandaloneSpecWorker(args, runner, scheduler)
    raise ValueError(f"unknown speculative algorithm {args.speculative_algorithm}")
