"""Per-request (Req) and per-batch (ScheduleBatch) CPU state.

Parity: reference ``python/sglang/srt/managers/schedule_batch.py`` (``Req:816``,
``init_next_round_input:1301``, ``update_finish_state:1632``, ``ScheduleBatch:2011``,
``prepare_for_extend:2378``, ``retract_decode:2825``, ``prepare_for_decode:3060``, ``filter_batch:3147``,
``merge_batch:3233``; finish reasons ``:225-292``).

B200-first differences: KV bookkeeping is page arrays on the host (see ``mem_cache/allocator.py``);
the batch keeps *vectorised numpy state* (``seq_lens``, ``req_slots``, ``page_rows``) that persists
across decode steps so a decode iteration costs a handful of numpy ops, and emits a page-table *delta*
(row, col, page) for the device instead of rebuilding indices.
"""
from __future__ import annotations

import enum
import time
from dataclasses import dataclass
from typing import Any, Dict, List, Optional, Tuple

import numpy as np

from ..sampling.sampling_params import SamplingParams


class ForwardMode(enum.IntEnum):
    EXTEND = 1
    DECODE = 2
    MIXED = 3
    IDLE = 4
    TARGET_VERIFY = 5
    DRAFT_EXTEND = 6

    def is_extend(self) -> bool:
        return self in (ForwardMode.EXTEND, ForwardMode.MIXED, ForwardMode.DRAFT_EXTEND, ForwardMode.TARGET_VERIFY)

    def is_decode(self) -> bool:
        return self == ForwardMode.DECODE


# --------------------------------------------------------------------------- finish reasons
class FinishReason:
    def __init__(self, kind: str, **kw: Any) -> None:
        self.kind, self.kw = kind, kw

    def to_json(self) -> Dict[str, Any]:
        return {"type": self.kind, **self.kw}

    def __repr__(self) -> str:
        return f"FinishReason({self.to_json()})"


def FINISH_MATCHED_TOKEN(tok: int) -> FinishReason:
    return FinishReason("stop", matched=int(tok))


def FINISH_MATCHED_STR(s: str) -> FinishReason:
    return FinishReason("stop", matched=s)


def FINISH_LENGTH(n: int) -> FinishReason:
    return FinishReason("length", length=int(n))


def FINISH_ABORT(msg: str = "aborted", status_code: int = 400) -> FinishReason:
    return FinishReason("abort", message=msg, status_code=status_code)


# --------------------------------------------------------------------------- Req
class Req:
    def __init__(self, rid: str, origin_input_ids: List[int], sampling_params: SamplingParams,
                 origin_input_text: Optional[str] = None, return_logprob: bool = False, top_logprobs_num: int = 0,
                 logprob_start_len: int = -1, stream: bool = False, lora_id: Optional[str] = None,
                 extra_key: Any = None, priority: Optional[int] = None, eos_token_ids: Optional[set] = None,
                 mm_inputs: Optional[Dict[str, Any]] = None, is_embedding: bool = False,
                 return_hidden_states: bool = False, session_id: Optional[str] = None):
        self.rid = rid
        self.origin_input_text = origin_input_text
        self.origin_input_ids = list(origin_input_ids)
        self.output_ids: List[int] = []
        self.sampling_params = sampling_params
        self.return_logprob = return_logprob
        self.top_logprobs_num = top_logprobs_num
        self.logprob_start_len = logprob_start_len
        self.stream = stream
        self.lora_id = lora_id
        self.extra_key = extra_key if lora_id is None else (extra_key, lora_id)
        self.priority = priority if priority is not None else 0
        self.eos_token_ids = eos_token_ids or set()
        self.mm_inputs = mm_inputs
        self.is_embedding = is_embedding
        self.return_hidden_states = return_hidden_states
        self.session_id = session_id

        # KV bookkeeping
        self.req_slot: Optional[int] = None
        self.page_ids: np.ndarray = np.empty(0, dtype=np.int32)  # pages covering tokens [0, kv_alloc_len)
        self.kv_len = 0                 # tokens whose KV is committed in the cache
        self.num_cached_pages = 0       # leading pages shared with (and locked in) the radix tree
        self.last_node: Any = None
        self.prefix_len = 0             # tokens served from the cache in the current prefill round
        self.extend_len = 0
        self.fill_len = 0               # tokens that will have KV after the current round
        self.is_chunked = 0
        self.cached_tokens = 0          # total prefix-cache hits (reported in meta_info)
        self.is_retracted = False
        self.retraction_count = 0

        # output / streaming state
        self.finished_reason: Optional[FinishReason] = None
        self.to_abort: Optional[FinishReason] = None
        self.host_state_flag: Optional[int] = None      # cached static part of ScheduleBatch.needs_host_state()
        self.samp4: Optional[tuple] = None               # cached (temperature, top_p, min_p, top_k) row for the packed step buffer
        self.sent_offset = 0            # number of output tokens already streamed
        self.decoded_text = ""          # scheduler-side incremental detokenisation (stop strings only)
        self.stop_check_offset = 0
        self.embedding: Any = None
        self.hidden_states: List[Any] = []

        # logprobs
        self.input_token_logprobs: Optional[List[Any]] = None
        self.output_token_logprobs: List[Tuple[float, int]] = []
        self.output_top_logprobs: List[Any] = []

        # grammar / jump-forward
        self.grammar: Any = None
        self.grammar_wait: Any = None
        self.jump_forward_tokens: List[int] = []

        # speculative decoding stats
        self.spec_verify_ct = 0
        self.spec_accepted_tokens = 0

        # timing
        self.created_time = time.perf_counter()
        self.queue_time_start = self.created_time
        self.first_token_time: Optional[float] = None
        self.finished_time: Optional[float] = None
        self.trace_context: Optional[Dict[str, str]] = None

    # ---- ids
    @property
    def seqlen(self) -> int:
        return len(self.origin_input_ids) + len(self.output_ids)

    def all_ids(self) -> List[int]:
        return self.origin_input_ids + self.output_ids

    def finished(self) -> bool:
        return self.finished_reason is not None

    # ---- finish checks (reference schedule_batch.py:1632 update_finish_state)
    def check_finished(self, tokenizer: Any = None, new_tokens: int = 1) -> None:
        if self.finished():
            return
        if self.to_abort is not None:
            self.finished_reason = self.to_abort
            return
        sp = self.sampling_params
        n_out = len(self.output_ids)
        if n_out == 0:
            return
        # check each of the newly appended tokens (spec decode / jump-forward append several)
        for i in range(max(0, n_out - new_tokens), n_out):
            tok = self.output_ids[i]
            if i + 1 >= sp.min_new_tokens:
                if not sp.ignore_eos and tok in self.eos_token_ids:
                    self.output_ids = self.output_ids[: i + 1]
                    self.finished_reason = FINISH_MATCHED_TOKEN(tok)
                    return
                if tok in sp.stop_token_ids:
                    self.output_ids = self.output_ids[: i + 1]
                    self.finished_reason = FINISH_MATCHED_TOKEN(tok)
                    return
            if i + 1 >= sp.max_new_tokens:
                self.output_ids = self.output_ids[: i + 1]
                self.finished_reason = FINISH_LENGTH(sp.max_new_tokens)
                return
        if (sp.stop or sp.stop_regex) and tokenizer is not None:
            # decode a bounded tail; long enough to contain any stop string
            max_stop = max([len(s) for s in sp.stop] + [32 if sp.stop_regex else 0])
            tail_tokens = self.output_ids[-(max_stop + 8 + new_tokens):]
            tail = tokenizer.decode(tail_tokens)
            for s in sp.stop:
                if s and s in tail:
                    self.finished_reason = FINISH_MATCHED_STR(s)
                    return
            if sp.stop_regex:
                import re

                for pat in sp.stop_regex:
                    m = re.search(pat, tail)
                    if m:
                        self.finished_reason = FINISH_MATCHED_STR(m.group(0))
                        return

    def reset_for_retract(self) -> None:
        self.page_ids = np.empty(0, dtype=np.int32)
        self.kv_len = 0
        self.num_cached_pages = 0
        self.last_node = None
        self.prefix_len = self.extend_len = self.fill_len = 0
        self.is_chunked = 0
        self.req_slot = None
        self.is_retracted = True
        self.retraction_count += 1

    def __repr__(self) -> str:
        return f"Req(rid={self.rid}, in={len(self.origin_input_ids)}, out={len(self.output_ids)}, kv={self.kv_len})"


# --------------------------------------------------------------------------- batch
@dataclass
class ScheduleBatch:
    reqs: List[Req]
    forward_mode: ForwardMode = ForwardMode.EXTEND
    # vectorised per-request state (decode fast path)
    seq_lens: Optional[np.ndarray] = None          # int32 [B]: kv length AFTER this step
    req_slots: Optional[np.ndarray] = None         # int32 [B]
    # per-step token-level arrays
    input_ids: Optional[np.ndarray] = None         # int32 [T] (extend) / None (decode: gathered on device)
    positions: Optional[np.ndarray] = None         # int32 [T]
    out_cache_loc: Optional[np.ndarray] = None     # int32 [T] slot ids
    mamba_snap: Optional[Dict[int, Any]] = None    # recurrent-state snapshots planned for this prefill step
    out_cache_loc_swa: Optional[np.ndarray] = None # int32 [T] slot ids in the sliding-window pool (hybrid SWA models)
    q_lens: Optional[np.ndarray] = None            # int32 [B]
    prefix_lens: Optional[np.ndarray] = None       # int32 [B]
    # page-table delta for the device (row=req_slot, col=page index, val=page id)
    pt_rows: Optional[np.ndarray] = None
    pt_cols: Optional[np.ndarray] = None
    pt_vals: Optional[np.ndarray] = None
    chunked_req: Optional[Req] = None
    mixed_ids: Any = None                          # ids of running requests riding along as one-token extends (--enable-mixed-chunk)
    return_logprob: bool = False
    has_grammar: bool = False
    spec_info: Any = None
    launch_time: float = 0.0
    # which requests produce a sampled token this step (chunked prefill reqs do not)
    sample_mask: Optional[np.ndarray] = None
    mm_embeds: Any = None
    decoding_reqs: Optional[List[Req]] = None      # mixed mode

    def batch_size(self) -> int:
        return len(self.reqs)

    def needs_host_state(self) -> bool:
        """True when sampling this batch reads per-request HOST state that must be current: penalties / logit bias (output-token counts),
        min_new_tokens still active (EOS suppression), per-request seeds (stream position), hidden-state return.  Such a batch cannot run
        one step ahead of result processing (overlap) and does not fit the captured decode graph, whose sampler sees only static tensors."""
        for r in self.reqs:
            f = r.host_state_flag
            if f is None:      # static part, computed once per request: 0 = never, 1 = always, 2 = while min_new_tokens is active
                sp = r.sampling_params
                if (sp.frequency_penalty or sp.presence_penalty or sp.repetition_penalty != 1.0 or sp.logit_bias
                        or sp.sampling_seed is not None or getattr(r, "return_hidden_states", False)):
                    f = 1
                else:
                    f = 2 if sp.min_new_tokens else 0
                r.host_state_flag = f
            if f == 1 or (f == 2 and len(r.output_ids) < r.sampling_params.min_new_tokens + 1):
                return True
        return False

    def is_empty(self) -> bool:
        return len(self.reqs) == 0

    # ---- decode-state maintenance
    def init_decode_state(self) -> None:
        self.seq_lens = np.asarray([r.kv_len for r in self.reqs], dtype=np.int32)
        self.req_slots = np.asarray([r.req_slot for r in self.reqs], dtype=np.int32)

    def filter_batch(self, keep: List[int]) -> None:
        """keep only the requests at indices ``keep`` (reference filter_batch:3147)"""
        if len(keep) == len(self.reqs):
            return
        self.reqs = [self.reqs[i] for i in keep]
        idx = np.asarray(keep, dtype=np.int64)
        if self.seq_lens is not None:
            self.seq_lens = self.seq_lens[idx]
            self.req_slots = self.req_slots[idx]

    def merge_batch(self, other: "ScheduleBatch") -> None:
        """append another batch's requests (reference merge_batch:3233)"""
        if self.seq_lens is None:
            self.init_decode_state()
        if other.seq_lens is None or len(other.seq_lens) != len(other.reqs):
            other.init_decode_state()
        self.reqs.extend(other.reqs)
        self.seq_lens = np.concatenate([self.seq_lens, other.seq_lens])
        self.req_slots = np.concatenate([self.req_slots, other.req_slots])
        self.return_logprob |= other.return_logprob
        self.has_grammar |= other.has_grammar
