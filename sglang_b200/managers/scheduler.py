"""Continuous-batching scheduler (one per GPU rank).

Parity: reference ``python/sglang/srt/managers/scheduler.py`` (``event_loop_normal:1748``,
``event_loop_overlap:1783``, ``get_next_batch_to_run:3080``, ``_get_new_batch_prefill_raw:3252``,
``update_running_batch:3566``, ``run_batch:3711``, ``process_batch_result:4035``, ``flush_cache:4376``,
``abort_request:4567``) and ``scheduler_components/batch_result_processor.py``.

Design (B200-first): the scheduler owns only host state.  A step (1) drains the inbox, (2) forms a batch
(prefill-first, chunked prefill, retract-on-OOM), (3) hands it to the worker which launches the GPU work
asynchronously, (4) processes the *previous* step's result while the GPU runs (one-iteration-lag overlap;
next-token ids never touch the host on the critical path: the sampler writes them to a device buffer
indexed by request slot and the next decode gathers from it).
"""
from __future__ import annotations

import logging
import os
import time
from collections import deque
from typing import Any, Callable, Deque, Dict, List, Optional, Tuple

import numpy as np

from ..environ import envs
from ..mem_cache.allocator import PagedAllocator
from ..mem_cache.kv_manager import KVManager
from ..mem_cache.radix_cache import create_radix_cache
from ..observability import trace as _trace
from ..sampling.sampling_params import SamplingParams
from ..server_args import ServerArgs
from . import io_struct as ios
from .schedule_batch import (FINISH_ABORT, FINISH_LENGTH, ForwardMode, Req, ScheduleBatch)
from .schedule_policy import AddResult, PrefillAdder, SchedulePolicy

logger = logging.getLogger(__name__)


class Scheduler:
    def __init__(self, server_args: ServerArgs, worker: Any, tokenizer: Any = None,
                 send_to_detokenizer: Optional[Callable[[Any], None]] = None, tp_rank: int = 0):
        self.args = server_args
        self.worker = worker
        self.tokenizer = tokenizer
        self.tp_rank = tp_rank
        self.send_to_detokenizer = send_to_detokenizer or (lambda m: None)
        self.page_size = server_args.page_size

        info = worker.get_memory_info()  # {"num_pages", "max_context_len", "max_running"}
        self.max_context_len = info["max_context_len"]
        self.max_running = info["max_running"]
        self.allocator = PagedAllocator(info["num_pages"], self.page_size)
        self.swa = info.get("swa")
        self.mamba = info.get("mamba")
        if self.mamba is not None and not server_args.disable_radix_cache:
            # state-space / hybrid model with snapshot rows: prefix caching through recurrent-state snapshots (mem_cache/mamba_radix.py)
            from ..mem_cache.mamba_radix import MambaKVManager, MambaRadixCache

            self.tree = MambaRadixCache(self.allocator, self.mamba["rows"], self.page_size, server_args.radix_eviction_policy,
                                        enable_kv_events=server_args.enable_kv_events)
            self.kv = MambaKVManager(self.allocator, self.tree, self.page_size, self.max_running, self.max_context_len,
                                     self.mamba["granularity"], worker.copy_mamba_state)
        elif self.swa is None:
            self.tree = create_radix_cache(self.allocator, self.page_size, server_args.radix_eviction_policy,
                                           disable=server_args.disable_radix_cache, impl=server_args.radix_cache_impl,
                                           enable_kv_events=server_args.enable_kv_events)
            self.kv = KVManager(self.allocator, self.tree, self.page_size, self.max_running, self.max_context_len)
        else:       # hybrid sliding-window model: second pool + page table, tombstoning prefix tree (mem_cache/swa.py)
            from ..mem_cache.swa import SWAKVManager, SWARadixCache

            self.swa_allocator = PagedAllocator(self.swa["num_pages"], self.page_size)
            self.tree = SWARadixCache(self.allocator, self.swa_allocator, self.swa["window"], self.page_size, server_args.radix_eviction_policy,
                                      disable=server_args.disable_radix_cache, enable_kv_events=server_args.enable_kv_events)
            self.kv = SWAKVManager(self.allocator, self.swa_allocator, self.tree, self.page_size, self.max_running, self.max_context_len,
                                   self.swa["window"], fetch_granularity=128 if getattr(worker, "is_cuda", False) else self.page_size)
        from .session_controller import SessionController

        self.sessions = SessionController(self.tree, self.page_size)
        worker.attach_kv_manager(self.kv)
        self.eplb = None
        if getattr(server_args, "enable_eplb", False) and getattr(worker, "model", None) is not None:
            from ..eplb.expert_distribution import EPLBManager, attach

            rec = getattr(worker, "expert_recorder", None) or attach(worker)
            if rec is not None:
                if rec.mode == "stat":
                    rec.mode = "window"
                self.eplb = EPLBManager(worker, rec, getattr(getattr(worker, "ep", None), "world_size", 1) or 1,
                                        server_args.eplb_rebalance_num_iterations, server_args.eplb_min_balancedness)
        self.policy = SchedulePolicy(server_args.schedule_policy, self.kv, server_args.enable_priority_scheduling,
                                     server_args.random_seed)

        self.waiting_queue: List[Req] = []
        self.running_batch = ScheduleBatch(reqs=[], forward_mode=ForwardMode.DECODE)
        self.running_batch.init_decode_state()
        self.last_batch: Optional[ScheduleBatch] = None
        self.step_wall_s = 0.0
        self.step_calls = 0
        self.chunked_req: Optional[Req] = None
        self.pending: Deque[Tuple[ScheduleBatch, Any]] = deque()
        self.enable_overlap = not server_args.disable_overlap_schedule and worker.supports_overlap()
        self.grammar_queue: List[Req] = []
        self.grammar_manager = None
        self.spec_worker = None
        self.paused = False

        # new-token-ratio estimator (reference scheduler.py:3580-3611)
        cons = server_args.schedule_conservativeness
        self.init_new_token_ratio = min(0.7 * cons, 1.0)
        self.min_new_token_ratio = min(self.init_new_token_ratio * 0.14, 1.0)
        self.new_token_ratio_decay = (self.init_new_token_ratio - self.min_new_token_ratio) / 600
        self.new_token_ratio = self.init_new_token_ratio

        self.forward_sleep_time: Optional[float] = None
        self.consensus_mode = os.environ.get("SGLB_RANK_CONSENSUS", "off")
        self.prefill_delayer = None
        self.min_free_slots_delayer = None
        self._init_delayers()

        self.forward_ct = 0
        self.decode_ct = 0
        self.num_retracted = 0
        self.num_generated_tokens = 0
        self.num_prefill_tokens = 0
        self.num_mixed_decode_tokens = self.num_mixed_fallbacks = 0
        self._has_timeouts = bool(server_args.request_waiting_timeout or server_args.request_running_timeout)
        a_ = server_args
        self._mixed_chunk = bool(a_.enable_mixed_chunk and not a_.speculative_algorithm and not a_.enable_dp_attention
                                 and getattr(a_, "disaggregation_mode", "null") in (None, "null") and not getattr(a_, "dllm_algorithm", None)
                                 and self.swa is None and self.mamba is None and a_.pp_size == 1)
        self.num_cached_tokens = 0
        self.last_decode_log = time.perf_counter()
        self.last_decode_tokens = 0
        self.gen_throughput = 0.0
        self.metrics = None
        self.watchdog_last_progress = time.perf_counter()
        self.batch_records: List[Dict[str, Any]] = []   # scripted-runtime hook (reference scheduler_hook.py:46-53)
        self.record_batches = server_args.scripted_scheduler
        self.eos_token_ids = set(worker.get_eos_token_ids())

    # ================================================================== request intake
    def handle_request(self, msg: Any) -> Optional[Any]:
        if isinstance(msg, ios.TokenizedGenerateReqInput):
            return self.handle_generate_request(msg)
        if isinstance(msg, ios.AbortReq):
            return self.abort_request(msg)
        if isinstance(msg, ios.FlushCacheReq):
            return ios.FlushCacheResp(success=self.flush_cache())
        if isinstance(msg, ios.GetInternalStateReq):
            return ios.GetInternalStateResp(state=self.get_internal_state())
        if isinstance(msg, ios.HealthCheckReq):
            return ios.HealthCheckResp(ok=True, forward_ct=self.forward_ct)
        if isinstance(msg, ios.GetKvEventsReq):
            ev = self.tree.take_kv_events() if hasattr(self.tree, "take_kv_events") else []
            return ios.GetKvEventsResp([list(e) for e in ev], self.page_size)
        if isinstance(msg, ios.GetLoadReq):
            return ios.GetLoadResp(num_running_reqs=len(self.running_batch.reqs), num_waiting_reqs=len(self.waiting_queue),
                                   num_tokens=int(self.kv.total_pages() - self.kv.available_pages()) * self.page_size,
                                   token_usage=self.kv.token_usage())
        if isinstance(msg, ios.RpcReq):
            self.drain()
            target = self if hasattr(self, msg.method) else self.worker
            if msg.method.startswith("_") or not hasattr(target, msg.method):
                return ios.RpcResp(False, None, f"no public method {msg.method!r} on the scheduler or the model runner")
            try:
                return ios.RpcResp(True, getattr(target, msg.method)(**(msg.kwargs or {})))
            except Exception as e:      # noqa: BLE001 - reported to the caller
                return ios.RpcResp(False, None, f"{type(e).__name__}: {e}")
        if isinstance(msg, ios.SlowDownReq):
            self.forward_sleep_time = msg.forward_sleep_time or None
            return ios.GenericResp(True, f"forward sleep = {self.forward_sleep_time}")
        if isinstance(msg, ios.PauseGenerationReq):
            return self.pause_generation(msg.mode)
        if isinstance(msg, ios.ContinueGenerationReq):
            self.paused = False
            return ios.GenericResp()
        if isinstance(msg, (ios.UpdateWeightsFromDiskReq, ios.UpdateWeightsFromTensorReq, ios.GetWeightsByNameReq,
                            ios.InitWeightsUpdateGroupReq, ios.DestroyWeightsUpdateGroupReq, ios.UpdateWeightsFromDistributedReq,
                            ios.ReleaseMemoryOccupationReq, ios.ResumeMemoryOccupationReq, ios.ProfileReq,
                            ios.LoadLoRAAdapterReq, ios.UnloadLoRAAdapterReq, ios.ExpertDistributionReq, ios.WeightsCheckerReq,
                            ios.SendWeightsToRemoteInstanceReq)):
            self.drain()
            if isinstance(msg, (ios.UpdateWeightsFromDiskReq, ios.UpdateWeightsFromTensorReq)) or \
                    (isinstance(msg, ios.UpdateWeightsFromDistributedReq) and msg.flush_cache):
                self.flush_cache()
            return self.worker.handle_control(msg)
        if isinstance(msg, ios.OpenSessionReq):
            sid, ok = self.sessions.open(msg.session_id, int(msg.capacity_of_str_len or 0))
            return ios.OpenSessionResp(session_id=sid, success=ok)
        if isinstance(msg, ios.CloseSessionReq):
            self.sessions.close(msg.session_id)
            return ios.GenericResp()
        raise TypeError(f"unknown message {type(msg)}")

    def handle_generate_request(self, m: ios.TokenizedGenerateReqInput) -> None:
        sp = SamplingParams.from_dict(m.sampling_params)
        session_err, session_id = None, None
        if m.session_params and m.session_params.get("id") is not None:
            # continue an earlier turn of the session: history (token ids) + this request's new input
            full, session_err = self.sessions.build_input(m.session_params, list(m.input_ids), m.rid)
            if full is not None:
                m.input_ids, session_id = full, m.session_params["id"]
        req = Req(m.rid, m.input_ids, sp, origin_input_text=m.input_text, return_logprob=m.return_logprob,
                  top_logprobs_num=m.top_logprobs_num, logprob_start_len=m.logprob_start_len, stream=m.stream,
                  lora_id=m.lora_id, extra_key=m.extra_key, priority=m.priority, eos_token_ids=self.eos_token_ids,
                  mm_inputs=m.mm_inputs, is_embedding=m.is_embedding, return_hidden_states=m.return_hidden_states)
        req.trace_context = m.trace_context
        if m.created_time:
            req.created_time = m.created_time
        req.session_id = session_id
        if session_err is not None:
            return self._reject(req, session_err)
        # validation (reference scheduler.py:2414-2783)
        n_in = len(req.origin_input_ids)
        if n_in == 0:
            return self._reject(req, "empty prompt")
        if n_in >= self.max_context_len:
            return self._reject(req, f"prompt ({n_in} tokens) exceeds the context length {self.max_context_len}")
        sp.max_new_tokens = min(sp.max_new_tokens, self.max_context_len - n_in)
        if self.args.max_queued_requests is not None and len(self.waiting_queue) >= self.args.max_queued_requests:
            return self._reject(req, "the request queue is full", 503)
        total_pages = (n_in + sp.max_new_tokens + self.page_size - 1) // self.page_size
        if (n_in + self.page_size) // self.page_size + 1 > self.kv.total_pages():
            return self._reject(req, "prompt does not fit in the KV cache")
        if sp.has_grammar and self.grammar_manager is not None:
            ready = self.grammar_manager.attach(req)
            if not ready:
                self.grammar_queue.append(req)
                return None
            if req.to_abort is not None:
                return self._reject(req, req.to_abort.kw.get("message", "invalid grammar"))
        elif sp.has_grammar:
            return self._reject(req, "grammar-constrained decoding requested but no grammar backend is enabled")
        self.waiting_queue.append(req)
        return None

    def _reject(self, req: Req, msg: str, code: int = 400) -> None:
        req.finished_reason = FINISH_ABORT(msg, code)
        self.stream_output([req])

    def abort_request(self, m: ios.AbortReq) -> None:
        hit = lambda r: m.abort_all or r.rid == m.rid or (m.rid and r.rid.startswith(m.rid))
        keep = []
        for r in self.waiting_queue:
            if hit(r):
                r.finished_reason = FINISH_ABORT()
                if r.req_slot is not None:
                    self.kv.release(r, insert=False)
                self.stream_output([r])
            else:
                keep.append(r)
        self.waiting_queue = keep
        # requests still compiling a grammar hold no KV yet: finish them here (they would otherwise be admitted after the abort)
        still = []
        for r in self.grammar_queue:
            if hit(r):
                r.finished_reason = FINISH_ABORT()
                self.stream_output([r])
            else:
                still.append(r)
        self.grammar_queue = still
        # everything that holds KV: the decode batch, the chunked request AND the prefill that has just run but is not merged yet
        # (self.last_batch) — the one-iteration window the first version missed, which silently dropped the abort
        for r in self.live_reqs():
            if hit(r) and not r.finished():
                r.to_abort = FINISH_ABORT()   # picked up at the next result processing
        return None

    def pause_generation(self, mode: str) -> Any:
        self.drain()
        self.paused = True
        if mode == "abort":
            self.abort_request(ios.AbortReq(abort_all=True))
            self.paused = False
        elif mode == "retract":
            self._retract(self.running_batch, list(range(len(self.running_batch.reqs))))
        return ios.GenericResp()

    def flush_cache(self) -> bool:
        self.drain()
        # live_reqs() covers running, chunked, waiting-with-KV and the just-run prefill (self.last_batch) that is not merged into the
        # running batch yet; flushing under the latter reset the tree / allocator beneath live requests (dec_lock_ref underflow)
        if any(not r.finished() for r in self.live_reqs()):
            logger.warning("flush_cache refused: requests in flight")
            return False
        self.last_batch = None
        self.sessions.unpin_all()
        self.tree.reset()
        self.allocator.clear()
        self.kv.slots.clear()
        self.kv.page_table[:] = 0
        self.kv.take_deltas()
        self.worker.flush()
        self.new_token_ratio = self.init_new_token_ratio
        return True

    # ================================================================== batch formation
    def get_next_batch_to_run(self) -> Optional[ScheduleBatch]:
        if self.paused:
            return None
        # merge the last prefill batch into the running (decode) batch (reference :3133-3157)
        if self.last_batch is not None and self.last_batch.forward_mode.is_extend():
            lb = self.last_batch
            keep = [i for i, r in enumerate(lb.reqs) if not r.finished() and r is not self.chunked_req
                    and not r.is_embedding and r.sampling_params.max_new_tokens > 0]
            if keep:
                lb.reqs = [lb.reqs[i] for i in keep]
                lb.init_decode_state()
                self.running_batch.merge_batch(lb)
            self.last_batch = None

        self._poll_grammar_queue()
        allow = self._prefill_allowed() if (self.prefill_delayer is not None or self.min_free_slots_delayer is not None) else True
        new_batch = self._get_new_batch_prefill() if allow else None
        if new_batch is not None:
            if self.prefill_delayer is not None:
                self.prefill_delayer.observe_prefill(len(new_batch.reqs))
            return new_batch
        if self.running_batch.reqs:
            b = self._update_running_batch()
            return b
        return None

    def _init_delayers(self) -> None:
        a = self.args
        from .prefill_delayer import MinFreeSlotsDelayer, PrefillDelayer, resolve_min_free_slots

        n = resolve_min_free_slots(a.min_free_slots, self.max_running)
        if n is not None:
            self.min_free_slots_delayer = MinFreeSlotsDelayer(n)
        if not a.enable_prefill_delayer:
            return
        gather = None
        if a.enable_dp_attention:
            from ..parallel import parallel_state as ps

            st = ps.get_state()
            stride = max(1, st.tp.world_size // max(1, st.attn_dp_size))
            gather = lambda row: st.tp.all_gather_pyobj(row)[::stride]       # one row per DP group (its attention-TP rank 0)
        self.prefill_delayer = PrefillDelayer(gather, a.dp_size, a.enable_dp_attention, a.prefill_delayer_max_delay_passes,
                                              a.prefill_delayer_token_usage_low_watermark, a.prefill_delayer_queue_min_ratio,
                                              a.prefill_delayer_max_delay_ms, a.prefill_max_requests)

    def _prefill_allowed(self) -> bool:
        """one negotiation per scheduler pass on every rank (also with nothing to prefill: keeps the gather matched)"""
        running = len(self.running_batch.reqs)
        if self.min_free_slots_delayer is not None and self.chunked_req is None and self.waiting_queue and \
                self.min_free_slots_delayer.should_delay(running, self.max_running - running):
            local_ok = False
        else:
            local_ok = True
        if self.prefill_delayer is None:
            return local_ok
        prefillable = local_ok and bool(self.waiting_queue) and self.chunked_req is None and running < self.max_running
        d = self.prefill_delayer.negotiate(prefillable, self.kv.token_usage(), running, len(self.waiting_queue), self.max_running)
        # a chunked request in flight always continues (its pages are already held)
        return local_ok and (d.allow or self.chunked_req is not None)

    def _poll_grammar_queue(self) -> None:
        if not self.grammar_queue:
            return
        still = []
        for r in self.grammar_queue:
            if self.grammar_manager.poll(r):
                if r.to_abort is not None:
                    self._reject(r, r.to_abort.kw.get("message", "invalid grammar"))
                else:
                    self.waiting_queue.append(r)
            else:
                still.append(r)
        self.grammar_queue = still

    def _get_new_batch_prefill(self) -> Optional[ScheduleBatch]:
        if not self.waiting_queue and self.chunked_req is None:
            return None
        running = self.running_batch.reqs
        if len(running) >= self.max_running and self.chunked_req is None:
            return None
        if envs.SGLB_TEST_RETRACT.get() and len(running) > 0 and self.chunked_req is None and \
                self.decode_ct % max(2, envs.SGLB_TEST_RETRACT_INTERVAL.get()) != 0:
            pass
        self.policy.calc_priority(self.waiting_queue)
        adder = PrefillAdder(self.kv, self.page_size, self.new_token_ratio, self.args.max_prefill_tokens,
                             self.args.chunked_prefill_size, running, self.max_running)
        if self.chunked_req is not None:
            ok = adder.add_chunked_req(self.chunked_req)
            if not ok:
                return None
        admitted = []
        for req in self.waiting_queue:
            res = adder.add_one_req(req)
            if res == AddResult.CONTINUE:
                admitted.append(req)
                if adder.new_chunked_req is not None:
                    break
                continue
            break
        if admitted:
            ids = set(map(id, admitted))
            self.waiting_queue = [r for r in self.waiting_queue if id(r) not in ids]
        if not adder.can_run:
            return None
        prev_chunked = self.chunked_req
        self.chunked_req = adder.new_chunked_req
        mixed = self._mixable_running() if self._mixed_chunk else []
        for r in mixed:          # a running request joins the extend step as a one-token extend over its own cached tokens
            r.prefix_len, r.fill_len = r.kv_len, r.kv_len + 1
        batch = ScheduleBatch(reqs=adder.can_run + mixed, forward_mode=ForwardMode.EXTEND, chunked_req=self.chunked_req)
        batch.mixed_ids = {id(r) for r in mixed}
        if mixed and not self._prepare_for_extend(batch):          # no room for the decode tokens on top of the prefill: plain prefill step
            self.num_mixed_fallbacks += 1
            for r in mixed:
                self.kv.trim(r, r.kv_len)          # (a page taken for the extra token must not be allocated a second time by the decode path)
            batch = ScheduleBatch(reqs=adder.can_run, forward_mode=ForwardMode.EXTEND, chunked_req=self.chunked_req)
            batch.mixed_ids, mixed = set(), []
            ok = self._prepare_for_extend(batch)
        else:
            ok = True if mixed else self._prepare_for_extend(batch)
        if ok and mixed:
            ids_ = batch.mixed_ids
            rb = self.running_batch
            rb.filter_batch([i for i, r in enumerate(rb.reqs) if id(r) not in ids_])      # they return through the last-batch merge
            self.num_mixed_decode_tokens += len(mixed)
        if not ok:
            # allocation failed despite the budget (fragmentation of reserves): put everything back
            for r in batch.reqs:
                if r is prev_chunked:
                    continue
                self.kv.release(r, insert=False)
                r.is_chunked = 0
                self.waiting_queue.insert(0, r)
            self.chunked_req = prev_chunked
            if not self.running_batch.reqs and not self.pending and prev_chunked is None:
                # nothing is running, nothing in flight, the tree had its chance to evict: this batch can never be placed.  Reject its largest
                # request instead of retrying forever (seen with a KV pool of a handful of pages: the engine span at 20 k iterations / s)
                victim = max((r for r in batch.reqs), key=lambda r: r.seqlen)
                self.waiting_queue = [r for r in self.waiting_queue if r is not victim]
                self._reject(victim, f"request ({victim.seqlen} tokens) does not fit in the KV cache "
                                     f"({self.kv.total_pages() * self.page_size} tokens" + (f", sliding-window pool {self.swa_allocator.total_pages * self.page_size}"
                                                                                              if self.swa is not None else "") + ")")
            return None
        return batch

    def _mixable_running(self) -> List[Req]:
        """``--enable-mixed-chunk`` (reference ``ScheduleBatch.mix_with_running``): the running requests ride along with this prefill step
        instead of stalling behind it.  The host must know their last token and finish state, so in-flight results are drained first;
        requests whose decode step needs more than a sampled token (grammar, logprobs, hidden states) keep the plain decode path."""
        if not self.running_batch.reqs:
            return []
        self.drain()
        return [r for r in self.running_batch.reqs if r.finished_reason is None and not r.finished() and r.grammar is None
                and not r.return_logprob and not getattr(r, "return_hidden_states", False) and r.to_abort is None]

    def _prepare_for_extend(self, batch: ScheduleBatch) -> bool:
        """allocate pages, build token-level arrays (reference prepare_for_extend:2378)"""
        ids, pos, loc, qlens, plens, slots, seqlens, sample = [], [], [], [], [], [], [], []
        loc_swa = [] if self.swa is not None else None
        for r in batch.reqs:
            if self.swa is not None:
                self.kv.slide(r)          # a long prompt's earlier chunks leave the window of the sliding layers
            if not self.kv.alloc_for_extend(r, r.fill_len):
                return False
        for r in batch.reqs:
            all_ids = r.all_ids()
            s, e = r.prefix_len, r.fill_len
            ids.append(np.asarray(all_ids[s:e], dtype=np.int32))
            pos.append(np.arange(s, e, dtype=np.int32))
            loc.append(self.kv.slots_for(r, s, e))
            if loc_swa is not None:
                loc_swa.append(self.kv.slots_for_swa(r, s, e))
            qlens.append(e - s)
            plens.append(s)
            slots.append(r.req_slot)
            seqlens.append(e)
            sample.append(e == len(all_ids))  # last chunk -> produces a token
            r.kv_len = e   # committed at launch: later steps are stream-ordered behind this one
            if r.queue_time_start is not None:
                r.queue_time_end = time.perf_counter()
            batch.return_logprob |= r.return_logprob
            batch.has_grammar |= r.grammar is not None
        batch.input_ids = np.concatenate(ids)
        batch.positions = np.concatenate(pos)
        batch.out_cache_loc = np.concatenate(loc)
        batch.out_cache_loc_swa = np.concatenate(loc_swa) if loc_swa is not None else None
        if self.mamba is not None and hasattr(self.kv, "plan_snapshot"):
            plan = {}
            for i, r in enumerate(batch.reqs):
                p = self.kv.plan_snapshot(r)
                if p is not None:
                    plan[i] = p
            batch.mamba_snap = plan or None
        batch.q_lens = np.asarray(qlens, dtype=np.int32)
        batch.prefix_lens = np.asarray(plens, dtype=np.int32)
        batch.req_slots = np.asarray(slots, dtype=np.int32)
        batch.seq_lens = np.asarray(seqlens, dtype=np.int32)
        batch.sample_mask = np.asarray(sample, dtype=bool)
        batch.pt_rows, batch.pt_cols, batch.pt_vals = self.kv.take_deltas()
        self.num_prefill_tokens += int(batch.q_lens.sum())
        self.num_cached_tokens += int(batch.prefix_lens.sum())
        return True

    def _update_running_batch(self) -> Optional[ScheduleBatch]:
        """decode step for the running batch; retract on OOM (reference update_running_batch:3566)"""
        rb = self.running_batch
        keep = [i for i, r in enumerate(rb.reqs) if r.finished_reason is None]
        rb.filter_batch(keep)
        if not rb.reqs:
            return None
        if self.swa is not None:
            for r in rb.reqs:
                self.kv.slide(r)          # sliding layers: give back pages no future token of this request can see
        force_retract = envs.SGLB_TEST_RETRACT.get() and self.decode_ct > 0 and \
            self.decode_ct % envs.SGLB_TEST_RETRACT_INTERVAL.get() == 0 and len(rb.reqs) > 1
        positions = rb.seq_lens.copy()   # index of the token written this step
        loc = None if force_retract else self.kv.alloc_for_decode(rb.reqs, rb.req_slots, positions)
        if loc is None:
            self.drain()   # never retract under an in-flight batch
            keep = [i for i, r in enumerate(rb.reqs) if not r.finished()]
            rb.filter_batch(keep)
            if not rb.reqs:
                return None
            positions = rb.seq_lens.copy()
            loc = None if force_retract else self.kv.alloc_for_decode(rb.reqs, rb.req_slots, positions)
            while loc is None and not force_retract and self.sessions.unpin_lru(8):
                # KV pressure: session pins are an optimisation — give their pages back before retracting / aborting live requests
                loc = self.kv.alloc_for_decode(rb.reqs, rb.req_slots, positions)
            while loc is None:
                if len(rb.reqs) <= 1 and not force_retract:
                    # cannot make progress: abort the last request
                    r = rb.reqs[0]
                    r.finished_reason = FINISH_ABORT("out of KV cache memory", 500)
                    self.kv.release(r, insert=False)
                    self.stream_output([r])
                    rb.filter_batch([])
                    return None
                victims = self._pick_retract_victims(rb, force_retract)
                force_retract = False
                self._retract(rb, victims)
                positions = rb.seq_lens.copy()
                loc = self.kv.alloc_for_decode(rb.reqs, rb.req_slots, positions)
            self.new_token_ratio = min(1.0, self.new_token_ratio + 0.2)   # be more conservative after a retract
        else:
            self.new_token_ratio = max(self.min_new_token_ratio, self.new_token_ratio - self.new_token_ratio_decay)
        rb.forward_mode = ForwardMode.DECODE
        rb.positions = positions.astype(np.int32)
        rb.out_cache_loc = loc
        rb.out_cache_loc_swa = self.kv.decode_slots_swa(rb.req_slots, positions) if self.swa is not None else None
        rb.seq_lens = rb.seq_lens + 1
        rb.q_lens = None
        rb.input_ids = None
        rb.sample_mask = None
        rb.pt_rows, rb.pt_cols, rb.pt_vals = self.kv.take_deltas()
        rb.has_grammar = rb.return_logprob = False
        for r in rb.reqs:
            r.kv_len += 1
            if r.grammar is not None:
                rb.has_grammar = True
            if r.return_logprob:
                rb.return_logprob = True
        # the worker gets a *snapshot* so later filter/merge of running_batch cannot race with it
        snap = ScheduleBatch(reqs=list(rb.reqs), forward_mode=ForwardMode.DECODE, seq_lens=rb.seq_lens.copy(),
                             req_slots=rb.req_slots.copy(), positions=rb.positions, out_cache_loc=rb.out_cache_loc,
                             out_cache_loc_swa=rb.out_cache_loc_swa, pt_rows=rb.pt_rows, pt_cols=rb.pt_cols, pt_vals=rb.pt_vals,
                             return_logprob=rb.return_logprob, has_grammar=rb.has_grammar)
        self.decode_ct += 1
        return snap

    def _pick_retract_victims(self, rb: ScheduleBatch, force: bool) -> List[int]:
        """reference retract_decode:2825/2889 — retract the requests with the fewest generated tokens first
        (cheapest to recompute), longest prompts breaking ties"""
        order = sorted(range(len(rb.reqs)), key=lambda i: (len(rb.reqs[i].output_ids), -len(rb.reqs[i].origin_input_ids)))
        need = self.kv.pages_needed_for_decode(rb.seq_lens)
        victims, freed = [], 0
        for i in order:
            if len(victims) >= len(rb.reqs) - 1:
                break
            victims.append(i)
            freed += rb.reqs[i].num_pages - rb.reqs[i].num_cached_pages
            if force or freed + self.kv.available_pages() >= need + len(rb.reqs):
                break
        return victims

    def _retract(self, rb: ScheduleBatch, victims: List[int]) -> None:
        vs = set(victims)
        for i in victims:
            r = rb.reqs[i]
            self.kv.release(r, insert=True)   # computed prefix stays reusable in the tree
            r.reset_for_retract()
            self.waiting_queue.insert(0, r)
            self.num_retracted += 1
        rb.filter_batch([i for i in range(len(rb.reqs)) if i not in vs])

    # ================================================================== run + results
    def run_batch(self, batch: ScheduleBatch) -> Any:
        self.forward_ct += 1
        if self.forward_sleep_time:
            time.sleep(self.forward_sleep_time)
        if self.consensus_mode != "off":        # SPMD: every TP rank must have formed the identical batch (utils/rank_consensus.py)
            from ..parallel import parallel_state as ps
            from ..utils.rank_consensus import check_consensus

            tp = ps.get_state().attn_tp if self.args.enable_dp_attention else ps.get_state().tp
            check_consensus(tp, (int(batch.forward_mode.value) if hasattr(batch.forward_mode, "value") else str(batch.forward_mode),
                                 [r.rid for r in batch.reqs], np.asarray(batch.seq_lens).tolist()), "Scheduler.run_batch", self.consensus_mode)
        batch.launch_time = time.perf_counter()
        if self.record_batches:
            self.batch_records.append({"mode": batch.forward_mode.name, "rids": [r.rid for r in batch.reqs],
                                       "chunked_rid": batch.chunked_req.rid if batch.chunked_req else None,
                                       "tokens": int(batch.q_lens.sum()) if batch.q_lens is not None else len(batch.reqs)})
        stuck = envs.SGLB_TEST_STUCK_SCHEDULER.get()
        if stuck > 0:
            time.sleep(stuck)
        if self.spec_worker is not None and batch.forward_mode.is_decode():
            return self.spec_worker.forward_batch_speculative(batch)
        return self.worker.forward_batch_generation(batch)

    def process_batch_result(self, batch: ScheduleBatch, result: Any) -> None:
        result.wait()
        self.watchdog_last_progress = time.perf_counter()
        if batch.forward_mode.is_extend():
            self._process_prefill_result(batch, result)
        else:
            self._process_decode_result(batch, result)

    def _process_prefill_result(self, batch: ScheduleBatch, result: Any) -> None:
        toks = result.next_token_ids
        out: List[Req] = []
        now = time.perf_counter()
        j = 0
        for i, r in enumerate(batch.reqs):
            produces = bool(batch.sample_mask[i])
            if r.is_embedding:
                r.embedding = result.embeddings[i] if result.embeddings is not None else None
                r.finished_reason = FINISH_LENGTH(0)
                self.kv.release(r)
                out.append(r)
                continue
            if r.return_logprob and result.input_logprobs is not None and result.input_logprobs[i] is not None and \
                    not (r.retraction_count > 0 and r.output_ids):       # a re-prefill after retraction reported its prompt already
                if r.input_token_logprobs is None:       # first computed chunk: the first reported token has no logprob of its own
                    first = max(int(batch.prefix_lens[i]), r.logprob_start_len)
                    r.input_token_logprobs = [(None, int(r.origin_input_ids[first]))] if first < len(r.origin_input_ids) else []
                r.input_token_logprobs = r.input_token_logprobs + result.input_logprobs[i]
            if not produces:
                # chunk of a longer prompt: publish the computed prefix (reference cache_unfinished_req)
                r.is_chunked = max(0, r.is_chunked - 1)
                self.kv.cache_unfinished(r)
                continue
            if r.finished():       # aborted while in flight
                continue
            if r.sampling_params.max_new_tokens == 0:   # prefill-only (cache warm-up / prompt scoring)
                r.finished_reason = FINISH_LENGTH(0)
                r.finished_time = now
                r.kv_len = min(r.kv_len, r.seqlen)
                self.kv.release(r)
                out.append(r)
                continue
            tok = int(toks[i])
            self._append_token(r, tok, result, i)
            if r.first_token_time is None:
                r.first_token_time = now
            r.check_finished(self.tokenizer)
            if r.to_abort is not None and not r.finished():
                r.finished_reason = r.to_abort
            if r.finished():
                r.finished_time = now
                r.kv_len = min(r.kv_len, r.seqlen - 1)
                self.kv.release(r)
            elif not (batch.mixed_ids and id(r) in batch.mixed_ids):
                self.kv.cache_unfinished(r)
            out.append(r)
        d = self.kv.take_deltas()
        if d[0].size:
            self._carry_deltas(d)
        self.stream_output(out)

    def _carry_deltas(self, d) -> None:
        # deltas produced while processing results ride with the next batch
        self.kv._delta_rows.insert(0, d[0]); self.kv._delta_cols.insert(0, d[1]); self.kv._delta_vals.insert(0, d[2])

    def _append_token(self, r: Req, tok: int, result: Any, i: int) -> None:
        r.output_ids.append(tok)
        self.num_generated_tokens += 1
        if r.return_logprob and result.next_token_logprobs is not None:
            r.output_token_logprobs.append((float(result.next_token_logprobs[i]), tok))
            if r.top_logprobs_num > 0 and result.top_logprobs is not None:
                r.output_top_logprobs.append(result.top_logprobs[i])
        if r.return_hidden_states and result.hidden_states is not None:
            r.hidden_states.append(result.hidden_states[i])
        if r.grammar is not None:
            try:
                r.grammar.accept_token(tok)
            except Exception as e:  # grammar rejected a sampled token: abort that request only
                r.to_abort = FINISH_ABORT(f"grammar error: {e}", 500)

    def _process_decode_result(self, batch: ScheduleBatch, result: Any) -> None:
        toks = result.next_token_ids
        out: List[Req] = []
        now = time.perf_counter()
        accept = getattr(result, "accept_lens", None)
        if accept is None and not batch.return_logprob and not batch.has_grammar and result.hidden_states is None:
            # plain decode (the hot case): one tight loop, the finish test inlined — at 1024 running requests the generic path below costs
            # ~1.8 ms of Python per iteration, more than a TP=8 decode step takes on the device (tools/host_step_bench.py)
            tl = toks.tolist()
            tokenizer = self.tokenizer
            n_app = 0
            for i, r in enumerate(batch.reqs):
                if r.finished_reason is not None or r.req_slot is None:
                    continue        # finished (or retracted) one step earlier; this lagging result is discarded
                tok = tl[i]
                oids = r.output_ids
                oids.append(tok)
                n_app += 1
                sp = r.sampling_params
                if (len(oids) >= sp.max_new_tokens or r.to_abort is not None or sp.stop or sp.stop_regex or sp.stop_token_ids
                        or (not sp.ignore_eos and tok in r.eos_token_ids)):
                    r.check_finished(tokenizer, 1)           # rare: let the full rule set decide
                    if r.to_abort is not None and r.finished_reason is None:
                        r.finished_reason = r.to_abort
                    if r.finished_reason is not None:
                        r.finished_time = now
                        r.kv_len = min(r.kv_len, r.seqlen - 1)
                        self.kv.release(r)
                out.append(r)
            self.num_generated_tokens += n_app
        else:
            for i, r in enumerate(batch.reqs):
                if r.finished() or r.req_slot is None:
                    continue        # finished (or retracted) one step earlier; this lagging result is discarded
                if accept is not None:
                    n_new = int(accept[i])
                    for t in result.accepted_tokens[i][:n_new]:
                        self._append_token(r, int(t), result, i)
                    r.spec_verify_ct += 1
                    r.spec_accepted_tokens += n_new - 1
                else:
                    n_new = 1
                    self._append_token(r, int(toks[i]), result, i)
                r.check_finished(self.tokenizer, n_new)
                if r.to_abort is not None and not r.finished():
                    r.finished_reason = r.to_abort
                if r.finished():
                    r.finished_time = now
                    # with overlap a lagging step may already have bumped kv_len; the KV of the final token is
                    # never needed, so clamp to what was really computed for tokens of all_ids()
                    r.kv_len = min(r.kv_len, r.seqlen - 1)
                    self.kv.release(r)
                elif r.grammar is not None and self.args.enable_jump_forward:
                    self._try_jump_forward(r)
                out.append(r)
        d = self.kv.take_deltas()
        if d[0].size:
            self._carry_deltas(d)
        self.stream_output(out)
        if self.decode_ct % max(1, self.args.decode_log_interval) == 0:
            self._log_decode_stats()

    def _try_jump_forward(self, r: Req) -> None:
        """Compressed-FSM jump-forward (reference v0.2 semantics; API ``base_grammar_backend.py:120-147``):
        when the grammar admits exactly one continuation string, append its tokens and re-prefill."""
        jf = r.grammar.try_jump_forward(self.tokenizer) if hasattr(r.grammar, "try_jump_forward") else None
        if not jf:
            return
        new_tokens = r.grammar.jump_forward_tokens(jf, self.tokenizer, r.output_ids)
        if not new_tokens:
            return
        room = r.sampling_params.max_new_tokens - len(r.output_ids)
        new_tokens = new_tokens[: max(0, room - 1)]
        if not new_tokens:
            return
        for t in new_tokens:
            r.output_ids.append(int(t))
            r.grammar.accept_token(int(t))
        self.num_generated_tokens += len(new_tokens)
        r.jump_forward_tokens.extend(new_tokens)
        # leave the decode batch and re-enter through the prefill path (KV of the prefix is reused via the tree)
        rb = self.running_batch
        if r in rb.reqs:
            rb.filter_batch([i for i, x in enumerate(rb.reqs) if x is not r])
        r.kv_len = min(r.kv_len, r.seqlen - 1 - len(new_tokens))
        self.kv.release(r, insert=True)
        r.reset_for_retract()
        r.is_retracted = False
        self.waiting_queue.insert(0, r)

    # ================================================================== output streaming
    def stream_output(self, reqs: List[Req]) -> None:
        """reference scheduler_components/output_streamer.py:113"""
        sel = []
        for r in reqs:
            fin = r.finished_reason is not None
            if fin and r.session_id is not None and not getattr(r, "_session_closed", False):
                r._session_closed = True
                if r.finished_reason.kind == "abort":
                    self.sessions.on_aborted(r.session_id, r.rid)
                else:
                    self.sessions.on_finished(r.session_id, r.rid, r.output_ids, r.extra_key)
            if fin:
                sel.append(r)
            elif r.stream:
                interval = r.sampling_params.stream_interval or self.args.stream_interval
                if len(r.output_ids) - r.sent_offset >= interval:
                    sel.append(r)
        if not sel or self.tp_rank != 0:
            for r in sel:
                r.sent_offset = len(r.output_ids)
            return
        msg = ios.BatchTokenIDOutput(
            rids=[r.rid for r in sel],
            finished_reasons=[r.finished_reason.to_json() if r.finished() else None for r in sel],
            output_ids=[r.output_ids[r.sent_offset:] for r in sel],
            read_offsets=[r.sent_offset for r in sel],
            prompt_tokens=[len(r.origin_input_ids) for r in sel],
            completion_tokens=[len(r.output_ids) for r in sel],
            cached_tokens=[r.cached_tokens for r in sel],
            skip_special_tokens=[r.sampling_params.skip_special_tokens for r in sel],
            no_stop_trim=[r.sampling_params.no_stop_trim for r in sel],
            stop_strs=[r.sampling_params.stop for r in sel],
        )
        if any(r.return_logprob for r in sel):
            msg.output_logprobs = [r.output_token_logprobs[r.sent_offset:] if r.return_logprob else None for r in sel]
            msg.input_logprobs = [r.input_token_logprobs if (r.return_logprob and r.sent_offset == 0) else None for r in sel]
            msg.top_logprobs = [r.output_top_logprobs[r.sent_offset:] if r.return_logprob else None for r in sel]
        if any(r.spec_verify_ct for r in sel):
            msg.spec_verify_ct = [r.spec_verify_ct for r in sel]
            msg.spec_accepted_tokens = [r.spec_accepted_tokens for r in sel]
        if any(r.is_embedding for r in sel):
            msg.embeddings = [r.embedding for r in sel]
        if any(r.return_hidden_states for r in sel):
            msg.hidden_states = [r.hidden_states for r in sel]
        msg.ttfts = [(r.first_token_time - r.created_time) if r.first_token_time is not None else None for r in sel]
        msg.queue_times = [(getattr(r, 'queue_time_end', None) or r.created_time) - r.created_time for r in sel]
        for r in sel:
            r.sent_offset = len(r.output_ids)
        self.send_to_detokenizer(msg)

    # ================================================================== loops
    def step(self) -> bool:
        """One scheduler iteration.  Returns False when there is nothing to do."""
        t0 = time.perf_counter()
        try:
            return self._step()
        finally:
            dt = time.perf_counter() - t0
            self.step_wall_s += dt      # minus BatchResult.wait_s = pure host time (bench.py reports the split)
            self.step_calls += 1
            if self.eplb is not None and self.last_batch is not None:
                self.eplb.on_forward_pass()
            if _trace.get_tracer() is not None and _trace.get_tracer().level >= 3:
                lb = self.last_batch
                now = time.time()
                _trace.get_tracer().step_span("scheduler_step", now - dt, now, mode=str(getattr(lb, "forward_mode", None)) if lb is not None else "idle",
                                              batch_size=len(lb.reqs) if lb is not None else 0, running=len(self.running_batch.reqs),
                                              waiting=len(self.waiting_queue))

    _stall_iters = 0

    def _check_starvation(self) -> None:
        """Nothing runs, nothing is in flight, yet requests wait: the head of the queue passed validation but can not be placed (its decode
        reserve or one of the pools is too small and there is nobody left to finish or be evicted).  After a few idle iterations reject it —
        without this the engine spins forever (observed at 20 k iterations / s with a KV pool of a few pages)."""
        if self.paused or not self.waiting_queue or self.running_batch.reqs or self.chunked_req is not None or self.grammar_queue:
            self._stall_iters = 0          # (a paused scheduler is idle on purpose: its queue is not starving)
            return
        if self.last_batch is not None and any(not r.finished() for r in self.last_batch.reqs):
            self._stall_iters = 0
            return
        self._stall_iters += 1
        if self._stall_iters < 64:
            return
        self._stall_iters = 0
        victim = self.waiting_queue.pop(0)
        if victim.req_slot is not None:
            self.kv.release(victim, insert=False)
        cap = f"{self.kv.total_pages() * self.page_size} tokens" + (f", sliding-window pool {self.swa_allocator.total_pages * self.page_size}" if self.swa is not None else "")
        self._reject(victim, f"request ({victim.seqlen} prompt tokens + reserve for {victim.sampling_params.max_new_tokens} new tokens) can not be "
                             f"scheduled in the KV cache ({cap})", 400)

    def _expire_requests(self) -> None:
        """``--request-waiting-timeout`` / ``--request-running-timeout`` (reference scheduler.py waiting / running timeouts): a queued request
        that has waited too long is finished with an abort (HTTP 408); a running one is flagged and ends at its next result"""
        now = time.perf_counter()
        wt, rt = self.args.request_waiting_timeout, self.args.request_running_timeout
        if wt and self.waiting_queue:
            keep = []
            for r in self.waiting_queue:
                if r is not self.chunked_req and now - r.created_time > wt and not r.output_ids:
                    if r.req_slot is not None:
                        self.kv.release(r, insert=False)
                    self._reject(r, f"request waited longer than {wt:g} s in the queue", 408)
                else:
                    keep.append(r)
            self.waiting_queue = keep
        if rt:
            for r in self.live_reqs():
                start = getattr(r, "queue_time_end", None) or r.created_time
                if r.to_abort is None and not r.finished() and now - start > rt:
                    r.to_abort = FINISH_ABORT(f"request ran longer than {rt:g} s", 408)

    def _step(self) -> bool:
        if self._has_timeouts and (self.forward_ct & 7) == 0 or (self._has_timeouts and not self.running_batch.reqs):
            self._expire_requests()
        batch = self.get_next_batch_to_run()
        if batch is None and self.args.enable_dp_attention:
            self.drain()
            self.worker.forward_idle()        # keep the other DP groups' collectives matched
            self._on_idle()
            return False
        if batch is None:
            if self.pending:
                self.drain()
                return True
            self._check_starvation()
            self._on_idle()
            return False
        self._stall_iters = 0
        # grammar state, penalties, min_new_tokens and per-request seeds read host-side request state: it must not lag by a step
        host_state = batch.needs_host_state()
        need_sync = (not self.enable_overlap) or batch.has_grammar or host_state or \
            (self.pending and (self.pending[-1][0].has_grammar or self.pending[-1][0].needs_host_state()))
        if need_sync:
            self.drain()
            if batch.has_grammar and batch.forward_mode.is_decode():
                # requests may have finished / jumped forward while draining
                alive = [i for i, r in enumerate(batch.reqs) if not r.finished() and r.req_slot is not None]
                if len(alive) != len(batch.reqs):
                    self._undo_decode_prepare(batch, alive)
                    if not alive:
                        return True
        result = self.run_batch(batch)
        self.last_batch = batch
        self.pending.append((batch, result))
        while len(self.pending) > (1 if self.enable_overlap and not need_sync else 0):
            b, res = self.pending.popleft()
            self.process_batch_result(b, res)
        if envs.SGLB_STRICT_MEM_CHECK.get():
            self.check_invariants()
        return True

    def _undo_decode_prepare(self, batch: ScheduleBatch, alive: List[int]) -> None:
        idx = np.asarray(alive, dtype=np.int64)
        batch.reqs = [batch.reqs[i] for i in alive]
        for name in ("seq_lens", "req_slots", "positions", "out_cache_loc"):
            setattr(batch, name, getattr(batch, name)[idx])

    def drain(self) -> None:
        while self.pending:
            b, res = self.pending.popleft()
            self.process_batch_result(b, res)

    def has_work(self) -> bool:
        return bool(self.waiting_queue or self.running_batch.reqs or self.chunked_req is not None or self.pending
                    or self.grammar_queue or (self.last_batch is not None and self.last_batch.forward_mode.is_extend()))

    def run_until_idle(self, max_steps: int = 1 << 60) -> int:
        n = 0
        while self.has_work() and n < max_steps:
            self.step()
            n += 1
        self.drain()
        return n

    def _on_idle(self) -> None:
        self.new_token_ratio = self.init_new_token_ratio
        if not self.has_work():
            self.check_invariants()

    # ================================================================== introspection
    def live_reqs(self) -> List[Req]:
        out = list(self.running_batch.reqs)
        if self.chunked_req is not None and self.chunked_req not in out:
            out.append(self.chunked_req)
        if self.last_batch is not None:
            out.extend(r for r in self.last_batch.reqs if r not in out)
        out.extend(r for r in self.waiting_queue if r.req_slot is not None and r not in out)
        return out

    def check_invariants(self) -> None:
        self.kv.check_invariants(self.live_reqs())

    def get_internal_state(self) -> Dict[str, Any]:
        return {
            "forward_ct": self.forward_ct, "num_running_reqs": len(self.running_batch.reqs),
            "num_waiting_reqs": len(self.waiting_queue), "token_usage": self.kv.token_usage(),
            "available_pages": self.kv.available_pages(), "total_pages": self.kv.total_pages(),
            "evictable_pages": self.tree.evictable_pages(), "protected_pages": self.tree.protected_pages(),
            "new_token_ratio": self.new_token_ratio, "num_retracted_reqs": self.num_retracted,
            "num_generated_tokens": self.num_generated_tokens, "num_prefill_tokens": self.num_prefill_tokens,
            "num_cached_tokens": self.num_cached_tokens, "gen_throughput": self.gen_throughput,
            "cache_hit_rate": self.num_cached_tokens / max(1, self.num_cached_tokens + self.num_prefill_tokens),
            "max_running_requests": self.max_running, "max_context_len": self.max_context_len,
            "page_size": self.page_size, "overlap": self.enable_overlap,
            **self._extra_state(),
        }

    def _extra_state(self) -> Dict[str, Any]:

        """everything else the metrics exporter publishes (cheap host counters only; never touches the device)"""
        w = self.worker
        out: Dict[str, Any] = {"decode_ct": self.decode_ct, "scheduler_iterations": self.step_calls, "scheduler_wall_s": self.step_wall_s,
                               "grammar_queue": len(getattr(self, "grammar_queue", ()) or ()), "num_sessions": len(getattr(self.sessions, "sessions", {}) or {}),
                               "chunked_req": int(self.chunked_req is not None), "pending_results": len(self.pending)}
        for k in ("graph_launches", "h2d_bytes", "d2h_bytes", "prefill_graph_hits"):
            v = getattr(w, k, None)
            if isinstance(v, (int, float)):
                out[k] = v
        if self.swa is not None:
            out.update(swa_total_pages=self.swa_allocator.total_pages, swa_available_pages=self.swa_allocator.available_pages(),
                       swa_slid_pages=self.kv.swa_slid_pages, swa_tombstoned_pages=self.tree.swa_tombstoned)
        sw = getattr(self, "spec_worker", None)
        if sw is not None:
            out.update(spec_accept_rate=sw.accept_rate(), spec_accept_length=sw.accept_length(), spec_verify_ct=sw.total_verify)
        lm = getattr(w, "lora_manager", None)
        if lm is not None and hasattr(lm, "pool"):
            out.update(lora_slots_used=len(getattr(lm.pool, "slot_of", {}) or {}), lora_adapters_loaded=len(getattr(lm, "adapters", {}) or {}))
        hc = getattr(self, "hicache", None)
        if hc is not None:
            for k in ("host_hits", "host_writes", "l3_hits", "l3_writes", "host_pages_used"):
                v = getattr(hc, k, None)
                if isinstance(v, (int, float)):
                    out["hicache_" + k] = v
        nvl = getattr(getattr(w, "tp", None), "nvl", None)
        if nvl is not None:
            out.update(nvl_collective_calls=nvl.calls, nvl_rs_calls=nvl.rs_calls)
        if self._mixed_chunk:
            out.update(num_mixed_decode_tokens=self.num_mixed_decode_tokens, num_mixed_fallbacks=self.num_mixed_fallbacks)
        return out

    def _log_decode_stats(self) -> None:
        now = time.perf_counter()
        dt = now - self.last_decode_log
        if dt <= 0:
            return
        self.gen_throughput = (self.num_generated_tokens - self.last_decode_tokens) / dt
        self.last_decode_log, self.last_decode_tokens = now, self.num_generated_tokens
        logger.info("Decode batch. #running-req: %d, #token: %d, token usage: %.2f, gen throughput (token/s): %.2f, "
                    "#queue-req: %d", len(self.running_batch.reqs),
                    (self.kv.total_pages() - self.kv.available_pages()) * self.page_size, self.kv.token_usage(),
                    self.gen_throughput, len(self.waiting_queue))
        if self.metrics is not None:
            self.metrics.log_stats(self.get_internal_state())
