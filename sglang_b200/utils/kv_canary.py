"""KV canary + token oracle: deterministic plumbing tests for the scheduler / cache / TP stack without a real model.

Parity: reference ``python/sglang/srt/kv_canary/*`` (shadow buffer per KV slot holding (token id, position, request) that
is written next to the real KV and verified before attention; ``token_oracle`` sampling backend that dictates the
"sampled" tokens; perturbation injectors that prove the detector fires).

The canary hooks ``ModelRunner._forward``: for every token written this step it records ``(token, position)`` at its
slot; for every request it then re-derives the slot of every position < seq_len from the page table and checks the
recorded (token, position) against the request's actual token sequence.  Any aliasing, stale page-table entry, premature
free or wrong prefix-cache hit shows up as a mismatch.
"""
from __future__ import annotations

import logging
from typing import Any, Callable, List, Optional

import numpy as np

logger = logging.getLogger(__name__)


class KVCanaryError(AssertionError):
    pass


class KVCanary:
    def __init__(self, num_slots: int, page_size: int, mode: str = "raise"):
        self.tok = np.full(num_slots, -1, dtype=np.int64)
        self.pos = np.full(num_slots, -1, dtype=np.int64)
        self.ps, self.mode = page_size, mode
        self.violations: List[str] = []
        self.checked = 0

    def on_batch(self, sb: Any, kv: Any) -> None:
        """call BEFORE the forward of ``sb`` (its out_cache_loc / positions are final, input ids known on the host)"""
        decode = sb.forward_mode.is_decode()
        B = len(sb.reqs)
        if decode:
            for i, r in enumerate(sb.reqs):
                ids = r.all_ids()
                p = int(sb.positions[i])
                if p < len(ids):
                    self.tok[sb.out_cache_loc[i]] = ids[p]
                    self.pos[sb.out_cache_loc[i]] = p
        else:
            self.tok[sb.out_cache_loc] = sb.input_ids
            self.pos[sb.out_cache_loc] = sb.positions
        # verify every slot that attention will read
        for i, r in enumerate(sb.reqs):
            if r.req_slot is None:
                continue
            L = int(sb.seq_lens[i])
            ids = r.all_ids()
            L = min(L, len(ids))
            ppos = np.arange(L)
            slots = kv.page_table[r.req_slot, ppos // self.ps].astype(np.int64) * self.ps + ppos % self.ps
            bad = np.flatnonzero((self.tok[slots] != np.asarray(ids[:L])) | (self.pos[slots] != ppos))
            self.checked += L
            if bad.size:
                j = int(bad[0])
                msg = (f"KV canary: rid={r.rid} position {j} reads slot {int(slots[j])} holding (tok={int(self.tok[slots[j]])}, "
                       f"pos={int(self.pos[slots[j]])}) but expects (tok={ids[j]}, pos={j}); {bad.size} bad of {L}")
                self.violations.append(msg)
                if self.mode == "raise":
                    raise KVCanaryError(msg)
                logger.error(msg)


def install_canary(engine: Any, mode: str = "raise") -> KVCanary:
    runner, sch = engine.runner, engine.scheduler
    canary = KVCanary(runner.kv_pool.num_slots, runner.page_size, mode)
    inner = runner.forward_batch_generation

    def wrapped(sb):
        canary.on_batch(sb, sch.kv)
        return inner(sb)

    runner.forward_batch_generation = wrapped
    engine.kv_canary = canary
    return canary


class TokenOracle:
    """sampling backend that returns predetermined tokens: f(rid, position) -> token (deterministic, model-free)"""

    def __init__(self, vocab_size: int, fn: Optional[Callable[[str, int], int]] = None):
        self.V = vocab_size
        self.fn = fn or (lambda rid, pos: (hash((rid, pos)) % (vocab_size - 300)) + 260)

    def install(self, engine: Any) -> None:
        import torch

        runner = engine.runner
        oracle = self

        def sample(fb, logits):
            reqs = runner._oracle_reqs
            toks = [oracle.fn(r.rid, len(r.origin_input_ids) + len(r.output_ids)) for r in reqs]
            t = torch.tensor(toks, dtype=torch.int32, device=logits.device)
            if fb.next_token_buf is not None and fb.sample_slots is not None:
                fb.next_token_buf[fb.sample_slots.long()] = t
            return t

        runner.token_oracle = sample
        runner.args.sampling_backend = "token_oracle"
        inner = runner._forward

        def fwd(sb):
            import numpy as np

            rows = np.arange(len(sb.reqs)) if sb.forward_mode.is_decode() else np.flatnonzero(sb.sample_mask)
            runner._oracle_reqs = [sb.reqs[i] for i in rows.tolist()]
            return inner(sb)

        runner._forward = fwd
        runner.graphs = {}     # the oracle replaces the captured sampler
