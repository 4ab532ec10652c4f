"""The block-diffusion decode step (reference ``srt/dllm/mixin/scheduler.py``), expressed on the verify path."""
from __future__ import annotations

import json
from dataclasses import dataclass, field
from typing import Any, Dict, List

import numpy as np
import torch

from ..managers.schedule_batch import ForwardMode, ScheduleBatch
from ..speculative.spec_worker import BaseSpecWorker, SpecResult
from .algorithms import ALGORITHMS

# architectures with a published block size / mask id (reference srt/dllm/config.py DLLM_PARAMS)
KNOWN = {"LLaDA2MoeModelLM": (32, 156895), "SDARForCausalLM": (4, 151669), "SDARMoeForCausalLM": (4, 151669)}


@dataclass
class DllmConfig:
    algorithm: str = "LowConfidence"
    block_size: int = 4
    mask_id: int = 0
    algorithm_config: Dict[str, Any] = field(default_factory=dict)

    @staticmethod
    def from_args(args: Any, architecture: str = "") -> "DllmConfig":
        bs, mid = KNOWN.get(architecture, (4, None))
        cfg: Dict[str, Any] = {}
        src = getattr(args, "dllm_algorithm_config", None)
        if isinstance(src, dict):
            cfg = dict(src)
        elif src:
            if str(src).lstrip().startswith("{"):
                cfg = json.loads(src)
            else:
                import yaml

                cfg = yaml.safe_load(open(src)) or {}
        bs = int(cfg.pop("block_size", getattr(args, "dllm_block_size", None) or bs))
        mid = cfg.pop("mask_id", getattr(args, "dllm_mask_id", None) if getattr(args, "dllm_mask_id", None) is not None else mid)
        if mid is None:
            raise ValueError(f"diffusion-LLM decoding needs --dllm-mask-id for architecture {architecture!r}")
        return DllmConfig(args.dllm_algorithm, bs, int(mid), cfg)


class DllmWorker(BaseSpecWorker):
    def __init__(self, args: Any, runner: Any, scheduler: Any):
        super().__init__(args, runner, scheduler)
        arch = getattr(getattr(runner, "cfg", None), "architecture", "") or ""
        self.cfg = DllmConfig.from_args(args, arch)
        if self.cfg.algorithm not in ALGORITHMS:
            raise ValueError(f"unknown dLLM algorithm {self.cfg.algorithm!r}; registered: {sorted(ALGORITHMS)}")
        self.algo = ALGORITHMS[self.cfg.algorithm](self.cfg.mask_id, self.cfg.algorithm_config)
        self.total_blocks = 0
        self.total_forwards = 0

    def forward_batch_speculative(self, batch: ScheduleBatch) -> SpecResult:
        sch, kv, runner = self.scheduler, self.scheduler.kv, self.runner
        B = len(batch.reqs)
        sizes, plens, locs = [], [], []
        for i, r in enumerate(batch.reqs):
            p = int(batch.positions[i])
            bs = max(0, min(self.cfg.block_size, r.sampling_params.max_new_tokens - len(r.output_ids), sch.max_context_len - p - 2))
            if bs and not kv.alloc_for_extend(r, p + bs + 1):
                bs = 0                                  # no KV room: emit nothing this step; retraction will make room
            sizes.append(bs)
            plens.append(p)
            locs.append(kv.slots_for(r, p, p + bs + 1))
        maxb = max(sizes + [1])
        dev = runner.device
        ids = torch.full((B, maxb), self.cfg.mask_id, dtype=torch.long, device=dev)
        valid = torch.zeros((B, maxb), dtype=torch.bool, device=dev)
        for i, bs in enumerate(sizes):
            valid[i, :bs] = True
        ids = torch.where(valid, ids, torch.zeros_like(ids))              # padding columns are never masks
        state = self.algo.init_state(ids)
        masks = []
        for i in range(B):                                 # row 0 (the last committed token) stays causal; block rows see everything
            ql, p = sizes[i] + 1, plens[i]
            m = np.ones((ql, p + ql), dtype=bool)
            m[0, p + 1:] = False
            masks.append(m.reshape(-1))
        vb = ScheduleBatch(reqs=batch.reqs, forward_mode=ForwardMode.TARGET_VERIFY)
        vb.positions = np.concatenate([p + np.arange(bs + 1, dtype=np.int32) for p, bs in zip(plens, sizes)])
        vb.out_cache_loc = np.concatenate(locs)
        vb.q_lens = np.asarray([bs + 1 for bs in sizes], dtype=np.int32)
        vb.prefix_lens = np.asarray(plens, dtype=np.int32)
        vb.req_slots = batch.req_slots
        vb.seq_lens = (vb.prefix_lens + vb.q_lens).astype(np.int32)
        vb.sample_mask = np.ones(B, dtype=bool)
        vb.has_grammar = False
        d0, d1 = (batch.pt_rows, batch.pt_cols, batch.pt_vals), kv.take_deltas()
        deltas = tuple(np.concatenate([a, b]) for a, b in zip(d0, d1))
        empty = tuple(np.zeros(0, dtype=a.dtype) for a in deltas)
        cu = np.concatenate([[0], np.cumsum(vb.q_lens)])
        rows = torch.as_tensor(np.concatenate([np.arange(cu[i] + 1, cu[i + 1]) for i in range(B)]) if sum(sizes) else np.zeros(0), dtype=torch.long, device=dev)
        sel = valid.reshape(-1).nonzero().squeeze(-1)
        lasts = [r.output_ids[-1] if r.output_ids else r.origin_input_ids[-1] for r in batch.reqs]
        done = torch.zeros(B, dtype=torch.bool, device=dev)
        for it in range(self.algo.max_steps(maxb) + 1):
            host = ids.cpu().numpy()
            vb.input_ids = np.concatenate([np.asarray([lasts[i]] + host[i, :sizes[i]].tolist(), dtype=np.int32) for i in range(B)])
            vb.pt_rows, vb.pt_cols, vb.pt_vals = deltas if it == 0 else empty
            vb.spec_info = {"verify_all_positions": True, "custom_mask": np.concatenate(masks), "return_logits": True}
            runner.forward_verify(vb)
            self.total_forwards += 1
            lg = runner.last_verify_logits                                   # [T, V]
            full = torch.zeros((B * maxb, lg.shape[-1]), dtype=lg.dtype, device=dev)
            full[sel] = lg[rows]
            new, fin = self.algo.step(ids, full.view(B, maxb, -1), state)
            new = torch.where(valid, new, ids)
            done |= fin | ~(valid & (ids == self.cfg.mask_id)).any(dim=1) & (new == ids).all(dim=1)
            ids = torch.where(done.unsqueeze(1), ids, new)
            if bool(done.all()):
                break
        res = SpecResult(B)
        host = ids.cpu().numpy()
        for i, r in enumerate(batch.reqs):
            out = [int(t) for t in host[i, :sizes[i]]]
            res.accept_lens[i] = len(out)
            res.accepted_tokens[i] = out
            res.next_token_ids[i] = out[-1] if out else lasts[i]
            if not out:                                # no room for a block: one ordinary next-token step from the causal row
                out = [int(torch.argmax(lg[int(cu[i])]))]
                res.accept_lens[i], res.accepted_tokens[i], res.next_token_ids[i] = 1, out, out[0]
            adv = len(out) - 1                         # KV of the last token and of the block's first bs-1 tokens is final; the block's
            r.kv_len += adv                            # last token is re-fed as next step's causal row 0 (identical for a last position)
            batch.seq_lens[i] += adv
            self.total_blocks += 1
            self.total_verify += 1
            self.total_accepted += max(len(out) - 1, 0)
            self.total_drafted += max(len(out) - 1, 0)
            kv.trim(r, r.kv_len)
        slots = torch.from_numpy(np.asarray(batch.req_slots, dtype=np.int64)).to(dev)
        runner.next_token_buf[slots] = torch.from_numpy(res.next_token_ids.astype(np.int32)).to(dev)
        rb = sch.running_batch
        idx = {id(r): j for j, r in enumerate(rb.reqs)}
        for i, r in enumerate(batch.reqs):
            j = idx.get(id(r))
            if j is not None:
                rb.seq_lens[j] = r.kv_len
        return res
