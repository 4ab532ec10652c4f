"""Engine: the in-process (offline) API and the core that the HTTP server wraps.

Parity: reference ``python/sglang/srt/entrypoints/engine.py:207`` (``generate:360``, ``async_generate``, ``encode``,
``flush_cache``, sessions, profiling, weight-update family ``:1396-1509``, LoRA load/unload, memory release/resume,
``shutdown:1258``) and the front half of ``managers/tokenizer_manager.py:386`` (tokenise, request state,
``_handle_batch_output:2166`` -> meta_info).

Topology (B200-first): ONE process per GPU.  The front half (tokeniser, detokeniser, request states) and the
scheduler share that process; two drive modes exist:
* synchronous: ``generate`` steps the scheduler itself until its requests finish — zero IPC, deterministic, and the
  only mode in which every TP rank can run SPMD under ``torchrun`` (all ranks make identical scheduling decisions);
* background: ``start_background_loop`` runs the scheduler on a thread fed by a thread-safe inbox; used by the HTTP
  server and ``async_generate``.  With TP>1, rank 0 broadcasts each iteration's new requests over the Gloo group
  (reference ``request_receiver.py:153-216``).
"""
from __future__ import annotations

import os

import asyncio
import dataclasses
import logging
import queue
import threading
import time
import uuid
from typing import Any, AsyncIterator, Dict, Iterator, List, Optional, Union

import torch

from ..configs.model_config import load_model_config
from ..environ import envs
from ..managers import io_struct as ios
from ..managers.detokenizer import Detokenizer
from ..managers.scheduler import Scheduler
from ..model_executor.model_runner import ModelRunner
from ..parallel import parallel_state as ps
from ..sampling.sampling_params import SamplingParams
from ..server_args import ServerArgs
from ..utils.tokenizer import get_tokenizer

logger = logging.getLogger(__name__)


class ReqState:
    __slots__ = ("rid", "text", "output_ids", "finished", "meta", "out_q", "created", "first_token_time", "obj", "logprobs",
                 "input_logprobs", "top_logprobs", "embedding", "hidden_states", "last_emit", "done")

    def __init__(self, rid: str, obj: Any, stream_q: Optional[Any]):
        self.rid, self.obj, self.out_q = rid, obj, stream_q
        self.text = ""
        self.output_ids: List[int] = []
        self.finished = False
        self.meta: Dict[str, Any] = {}
        self.created = time.perf_counter()
        self.first_token_time: Optional[float] = None
        self.logprobs: List[Any] = []
        self.input_logprobs: Any = None
        self.top_logprobs: List[Any] = []
        self.embedding = None
        self.hidden_states = None
        self.last_emit = 0
        self.done = threading.Event()      # set when the final output has been applied (multi-process front end waits on it)

    def result(self) -> Dict[str, Any]:
        out = {"text": self.text, "output_ids": list(self.output_ids), "meta_info": dict(self.meta)}
        if self.embedding is not None:
            out = {"embedding": self.embedding, "meta_info": dict(self.meta)}
        return out


class Engine:
    def __init__(self, server_args: Optional[ServerArgs] = None, **kwargs: Any):
        self.server_args = server_args or ServerArgs(**kwargs)
        a = self.server_args
        logging.basicConfig(level=getattr(logging, a.log_level.upper(), logging.INFO),
                            format="[%(asctime)s %(name)s] %(message)s")
        world = a.tp_size * a.pp_size
        if world > 1:
            ps.init_distributed(a.tp_size, a.pp_size, a.ep_size, device=a.device, enable_nvl=a.comm_backend != "nccl",
                                attn_dp_size=a.dp_size if a.enable_dp_attention else 1, dcp_size=a.dcp_size,
                                cp_size=a.cp_size if a.enable_prefill_cp else 1)
        else:
            if a.device == "cuda":
                torch.cuda.set_device(a.base_gpu_id)
            ps.set_state(None)
            ps.get_state()
        st = ps.get_state()
        self.tp_rank, self.pp_rank = st.tp.rank, st.pp.rank
        self.is_driver = st.rank == 0
        self.model_config = load_model_config(a.model_path, __import__('json').loads(a.json_model_override_args or '{}') or None, a.context_length)
        self.runner = ModelRunner(a, self.model_config)
        self.tokenizer = None if a.skip_tokenizer_init else get_tokenizer(
            a.tokenizer_path, self.model_config.vocab_size, self.model_config.eos_token_id, self.model_config.bos_token_id,
            a.trust_remote_code)
        self.detok = Detokenizer(self.tokenizer) if self.tokenizer is not None else None
        emits_output = st.tp.rank == 0 and st.pp.is_last if st.pp.world_size > 1 else st.tp.rank == 0
        # DP-attention: every attention group serves its own requests (round-robin by arrival); the first rank of each group emits
        # that group's outputs and all of them are gathered to rank 0 once per iteration
        self._dp = st.attn_dp_size if a.enable_dp_attention else 1
        self._dp_rank = st.attn_dp_rank
        self._dp_ctr = 0
        if self._dp > 1:
            emits_output = st.attn_tp.rank == 0
        # pipeline parallel: the LAST stage samples, rank 0 talks to the client -> the last stage's tp-rank-0 relays every step's
        # output messages to rank 0 over the Gloo twin of the world group (both run the same SPMD loop, one hand-off per step)
        self._relay_src = (st.pp.world_size - 1) * st.tp.world_size if st.pp.world_size > 1 else None
        self._relay_buf: List[Any] = []
        self._world = st.world
        if getattr(self.runner.model, "disable_chunked_prefill", False):
            a.chunked_prefill_size = None       # encoder-only models attend bidirectionally: a request is always encoded in one pass
        if getattr(self.runner.model, "disable_radix_cache", False) and not a.disable_radix_cache and not self.runner.mamba_snapshot_rows:
            logger.info("%s keeps recurrent state per request: prefix caching disabled", type(self.runner.model).__name__)
            a.disable_radix_cache = True
            if a.schedule_policy in ("lpm", "dfs-weight"):
                a.schedule_policy = "fcfs"
        self.scheduler = Scheduler(a, self.runner, self.tokenizer, self._on_scheduler_output, 0 if emits_output else 1)
        self._install_optional_components()
        self.states: Dict[str, ReqState] = {}
        self._lock = threading.RLock()
        self._inbox: "queue.Queue[Any]" = queue.Queue()
        self._bg_thread: Optional[threading.Thread] = None
        self._bg_stop = threading.Event()
        self._wake = threading.Event()
        self.metrics = None
        if a.enable_metrics:
            from ..observability.metrics import MetricsCollector

            self.metrics = MetricsCollector(a.served_model_name)
            self.scheduler.metrics = self.metrics
        self.weight_version = "default"
        self.start_time = time.time()
        self.tracer = None
        if a.enable_trace:
            from ..observability.trace import get_tracer, init_tracing

            ep = a.otlp_traces_endpoint if str(a.otlp_traces_endpoint or "").startswith("http") else None    # OTLP/HTTP JSON; a bare host:port is gRPC (no SDK here)
            self.tracer = get_tracer() or init_tracing(level=int(getattr(a, "trace_level", 2) or 2), file=getattr(a, "trace_file", None), otlp_endpoint=ep)

    def _install_optional_components(self) -> None:
        a = self.server_args
        if a.grammar_backend != "none" and self.tokenizer is not None:
            try:
                from ..constrained.grammar_manager import GrammarManager

                self.scheduler.grammar_manager = GrammarManager(self.tokenizer, self.model_config.vocab_size, a.grammar_backend,
                                                                 reasoning_parser=a.reasoning_parser)
            except Exception as e:  # grammar backends are optional at import time
                logger.warning("grammar backend unavailable: %s", e)
        if a.speculative_algorithm or a.dllm_algorithm:
            from ..speculative.spec_worker import create_spec_worker

            self.scheduler.spec_worker = create_spec_worker(a, self.runner, self.scheduler)
        if a.enable_lora:
            from ..lora.lora_manager import LoRAManager

            self.runner.lora_manager = LoRAManager(self.runner, a.max_loras_per_batch, a.max_lora_rank)
            for p in a.lora_paths or []:
                name, _, path = p.partition("=")
                self.runner.lora_manager.load_adapter(name, path or name)
            self.runner.recapture_graphs()        # decode graphs must contain the SGMV kernels the manager just wrapped in
        if a.enable_hierarchical_cache:
            from ..mem_cache.hicache import attach_hicache

            attach_hicache(self.scheduler, self.runner, a)

    # ================================================================== request construction
    def _tokenize(self, g: ios.GenerateReqInput) -> ios.TokenizedGenerateReqInput:
        sp = dict(g.sampling_params or {})
        if g.input_ids is not None:
            ids, text = list(g.input_ids), g.text
        else:
            if self.tokenizer is None:
                raise ValueError("text input requires a tokenizer (skip_tokenizer_init is set)")
            text = g.text
            ids = self.tokenizer.encode(text)
            bos = getattr(self.tokenizer, "bos_token_id", None)
            if isinstance(self.tokenizer, object) and type(self.tokenizer).__name__ == "ByteTokenizer" and bos is not None:
                ids = [bos] + ids
        mm = None
        if g.image_data is not None and self.model_config.is_multimodal:
            from ..multimodal.processor import process_images

            ids, mm = process_images(ids, g.image_data, self.model_config, self.tokenizer, self._encoder_client())
        if g.audio_data is not None and self.model_config.audio is not None:
            from ..multimodal.audio import process_audio

            if self.model_config.audio.get("llm"):          # audio clips spliced into a decoder-only prompt (Qwen2-Audio)
                from ..multimodal.audio import process_audio_llm

                ids, mm = process_audio_llm(ids, g.audio_data, self.model_config)
            else:
                mm = process_audio(g.audio_data, self.model_config)
        SamplingParams.from_dict(sp).verify(self.model_config.vocab_size)
        return ios.TokenizedGenerateReqInput(
            rid=g.rid or self._new_rid(), input_text=text, input_ids=ids, sampling_params=sp, return_logprob=bool(g.return_logprob),
            logprob_start_len=int(g.logprob_start_len), top_logprobs_num=int(g.top_logprobs_num), stream=bool(g.stream),
            return_hidden_states=g.return_hidden_states, lora_id=g.lora_path, session_params=g.session_params, mm_inputs=mm,
            custom_logit_processor=g.custom_logit_processor, bootstrap_host=g.bootstrap_host, bootstrap_port=g.bootstrap_port,
            bootstrap_room=g.bootstrap_room, priority=g.priority, extra_key=g.extra_key, created_time=time.perf_counter())

    def _encoder_client(self):
        """EPD: client of the vision-encoder fleet (``--encoder-urls``), None when images are encoded locally"""
        if not self.server_args.encoder_urls:
            return None
        c = getattr(self, "_enc_client", None)
        if c is None:
            from ..disaggregation.encoder import EncoderClient

            c = self._enc_client = EncoderClient(self.server_args.encoder_urls)
        return c

    def _submit(self, g: ios.GenerateReqInput, stream_q: Optional[Any] = None, is_embedding: bool = False) -> ReqState:
        t = self._tokenize(g)
        t.is_embedding = is_embedding
        stt = ReqState(t.rid, g, stream_q)
        stt.meta = {"id": t.rid, "prompt_tokens": len(t.input_ids), "completion_tokens": 0, "cached_tokens": 0,
                    "finish_reason": None, "weight_version": self.weight_version}
        with self._lock:
            self.states[t.rid] = stt
        if self._dp > 1:
            owner = self._dp_ctr % self._dp            # deterministic on every rank (SPMD): arrival order round-robin
            self._dp_ctr += 1
            stt.meta["dp_rank"] = owner
            if owner != self._dp_rank:
                return stt                              # another attention group serves it; rank 0 still tracks its state
        self._dispatch(t)
        return stt

    def _dispatch(self, t: ios.TokenizedGenerateReqInput) -> None:
        """hand a tokenised request to the scheduler: its inbox (background loop), a direct call (offline), or — in the multi-process front
        end (managers/multiprocess.py) — a ZMQ frame"""
        if self._bg_thread is not None:
            self._inbox.put(t)
            self._wake.set()
        else:
            self.scheduler.handle_request(t)

    # ================================================================== output path
    def _any_rank_has_work(self) -> bool:
        """DP-attention: groups hold different requests, so the loop condition must be global (one tiny Gloo all-reduce per iteration)"""
        mine = self.scheduler.has_work()
        if self._dp <= 1:
            return mine
        import torch.distributed as dist

        t = torch.tensor([1 if mine else 0], dtype=torch.int32)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self._world.cpu_group)
        return bool(t.item())

    def _new_rid(self) -> str:
        """request ids must agree on every rank of an SPMD job (results are matched by id when a later pipeline stage relays them)"""
        if self._world.world_size > 1:
            self._rid_ctr = getattr(self, "_rid_ctr", 0) + 1
            return f"req-{self._rid_ctr:08d}"
        return uuid.uuid4().hex

    def _relay_sync(self) -> None:
        """PP only: ship this step's output messages from the sampling stage to rank 0 (no-op otherwise)"""
        if self._dp > 1:
            # every group's emitter contributes this iteration's messages; rank 0 consumes all of them (its own come back too)
            allm = self._world.all_gather_pyobj(self._relay_buf)
            self._relay_buf = []
            if ps.get_state().rank == 0:
                for msgs in allm:
                    for m in msgs:
                        self._on_scheduler_output(m, relayed=True)
            return
        if self._relay_src is None:
            return
        rank = ps.get_state().rank
        if rank == self._relay_src:
            self._world.send_pyobj(self._relay_buf, 0)
            self._relay_buf = []
        elif rank == 0:
            for m in self._world.recv_pyobj(self._relay_src):
                self._on_scheduler_output(m, relayed=True)

    def _on_scheduler_output(self, msg: ios.BatchTokenIDOutput, relayed: bool = False) -> None:
        if not relayed and (self._dp > 1 or (self._relay_src is not None and ps.get_state().rank == self._relay_src)):
            self._relay_buf.append(msg)
            return
        out = self.detok.process(msg) if self.detok is not None else ios.BatchStrOutput(
            rids=msg.rids, finished_reasons=msg.finished_reasons, output_strs=[""] * len(msg.rids), output_ids=msg.output_ids,
            prompt_tokens=msg.prompt_tokens, completion_tokens=msg.completion_tokens, cached_tokens=msg.cached_tokens,
            output_logprobs=msg.output_logprobs, input_logprobs=msg.input_logprobs, top_logprobs=msg.top_logprobs,
            spec_verify_ct=msg.spec_verify_ct, spec_accepted_tokens=msg.spec_accepted_tokens, embeddings=msg.embeddings,
            hidden_states=msg.hidden_states, ttfts=msg.ttfts, queue_times=msg.queue_times)
        self._apply_output(out)

    def _apply_output(self, out: ios.BatchStrOutput) -> None:
        """fold one detokenised batch into the request states and wake their waiters (called in-process, or by the front end's receive thread)"""
        now = time.perf_counter()
        for i, rid in enumerate(out.rids):
            st = self.states.get(rid)
            if st is None:
                continue
            st.text += out.output_strs[i]
            st.output_ids.extend(out.output_ids[i])
            if st.first_token_time is None and out.output_ids[i]:
                st.first_token_time = now
            m = st.meta
            m["completion_tokens"] = out.completion_tokens[i]
            m["cached_tokens"] = out.cached_tokens[i]
            if out.prompt_tokens is not None and out.prompt_tokens[i]:
                m["prompt_tokens"] = out.prompt_tokens[i]      # authoritative (sessions prepend history inside the scheduler)
            if out.output_logprobs is not None and out.output_logprobs[i] is not None:
                st.logprobs.extend(out.output_logprobs[i])
                m["output_token_logprobs"] = [(lp, tid, None) for lp, tid in st.logprobs]
                if out.input_logprobs is not None and out.input_logprobs[i] is not None:
                    m["input_token_logprobs"] = [(lp, tid, None) for lp, tid in out.input_logprobs[i]]
                if out.top_logprobs is not None and out.top_logprobs[i]:
                    st.top_logprobs.extend(out.top_logprobs[i])
                    m["output_top_logprobs"] = st.top_logprobs
            if out.spec_verify_ct is not None:
                m["spec_verify_ct"] = out.spec_verify_ct[i]
                v = max(1, out.spec_verify_ct[i])
                m["spec_accept_length"] = 1.0 + out.spec_accepted_tokens[i] / v
                m["spec_accepted_drafts"] = out.spec_accepted_tokens[i]
            if out.embeddings is not None and out.embeddings[i] is not None:
                st.embedding = out.embeddings[i]
            if out.hidden_states is not None and out.hidden_states[i]:
                m["hidden_states"] = out.hidden_states[i]
            fr = out.finished_reasons[i]
            if fr is not None:
                st.finished = True
                m["finish_reason"] = fr
                m["e2e_latency"] = now - st.created
                if out.ttfts is not None and out.ttfts[i] is not None:
                    m["ttft"] = out.ttfts[i]          # measured by the scheduler at the step that produced token 1
                    m["queue_time"] = out.queue_times[i] if out.queue_times else None
                elif st.first_token_time is not None:
                    m["ttft"] = st.first_token_time - st.created
                if self.metrics is not None:
                    self.metrics.observe_request(m)
                if getattr(self, "tracer", None) is not None:
                    tc = getattr(st.obj, "trace_context", None) if st.obj is not None else None
                    self.tracer.trace_request(rid, m, time.time() - m["e2e_latency"], tc)
            if st.out_q is not None:
                st.out_q.put(st.result() if not st.finished else {**st.result(), "_done": True})
            if st.finished:
                with self._lock:
                    self.states.pop(rid, None)
                st.done.set()

    # ================================================================== public API
    def generate(self, prompt: Optional[Union[str, List[str]]] = None, sampling_params: Optional[Union[Dict, List[Dict]]] = None,
                 input_ids: Optional[Union[List[int], List[List[int]]]] = None, image_data: Any = None,
                 return_logprob: Union[bool, List[bool]] = False, logprob_start_len: Union[int, List[int]] = -1,
                 top_logprobs_num: Union[int, List[int]] = 0, lora_path: Any = None, stream: bool = False,
                 return_hidden_states: bool = False, rid: Any = None, priority: Optional[int] = None,
                 session_params: Optional[Dict] = None, extra_key: Optional[str] = None, audio_data: Any = None):
        g = ios.GenerateReqInput(text=prompt, input_ids=input_ids, image_data=image_data, audio_data=audio_data, sampling_params=sampling_params,
                                 return_logprob=return_logprob, logprob_start_len=logprob_start_len,
                                 top_logprobs_num=top_logprobs_num, lora_path=lora_path, stream=stream,
                                 return_hidden_states=return_hidden_states, rid=rid, priority=priority,
                                 session_params=session_params, extra_key=extra_key)
        is_batch = g.is_batch()
        singles = g.normalize()
        if stream:
            if is_batch or len(singles) != 1:
                raise ValueError("stream=True supports a single request")
            return self._stream_one(singles[0])
        states = [self._submit(s, None) for s in singles]
        self._drive(states)
        res = [s.result() for s in states]
        return res if (is_batch or len(res) > 1) else res[0]

    def _drive(self, states: List[ReqState]) -> None:
        if self._bg_thread is not None:
            while not all(s.finished for s in states):
                time.sleep(0.0005)
            return
        sch = self.scheduler
        while self._any_rank_has_work():
            sch.step()
            self._relay_sync()
        sch.drain()
        self._relay_sync()
        if self.is_driver:
            pending = [s for s in states if not s.finished]
            if pending:
                raise RuntimeError(f"{len(pending)} requests did not finish (scheduler idle)")

    def _stream_one(self, g: ios.GenerateReqInput) -> Iterator[Dict[str, Any]]:
        q: "queue.Queue[Dict[str, Any]]" = queue.Queue()
        g.stream = True
        st = self._submit(g, q)
        sch = self.scheduler
        while True:
            if self._bg_thread is None:
                while q.empty() and sch.has_work():
                    sch.step()
                if q.empty():
                    sch.drain()
            try:
                item = q.get(timeout=60 if self._bg_thread is not None else 0.001)
            except queue.Empty:
                if self._bg_thread is None and not sch.has_work():
                    return
                continue
            done = item.pop("_done", False)
            yield item
            if done:
                return

    async def async_generate(self, prompt=None, sampling_params=None, input_ids=None, image_data=None, stream: bool = False,
                             **kw) -> Union[Dict[str, Any], AsyncIterator[Dict[str, Any]]]:
        self.start_background_loop()
        loop = asyncio.get_running_loop()
        g = ios.GenerateReqInput(text=prompt, input_ids=input_ids, image_data=image_data, sampling_params=sampling_params,
                                 stream=stream, **kw)
        singles = g.normalize()
        if self.server_args.enable_dynamic_batch_tokenizer and self.tokenizer is not None:
            # text prompts are encoded off the event loop, batched with whatever else arrived in the same couple of milliseconds
            dt = getattr(self, "_dyn_tok", None)
            if dt is None:
                from ..managers.async_batch_tokenizer import AsyncDynamicBatchTokenizer

                dt = self._dyn_tok = AsyncDynamicBatchTokenizer(self.tokenizer, self.server_args.dynamic_batch_tokenizer_batch_size,
                                                                self.server_args.dynamic_batch_tokenizer_batch_timeout)
            todo = [s for s in singles if s.input_ids is None and isinstance(s.text, str) and s.image_data is None and s.audio_data is None]
            if todo and type(self.tokenizer).__name__ != "ByteTokenizer":
                for s_, ids_ in zip(todo, await asyncio.gather(*[dt.encode(s.text) for s in todo])):
                    s_.input_ids = ids_
        aq: "asyncio.Queue[Dict[str, Any]]" = asyncio.Queue()

        class _Bridge:
            def put(self_inner, item):
                loop.call_soon_threadsafe(aq.put_nowait, item)

        states = [self._submit(s, _Bridge()) for s in singles]

        async def gen():
            remaining = len(states)
            while remaining:
                item = await aq.get()
                if item.pop("_done", False):
                    remaining -= 1
                    yield item
                elif stream:
                    yield item

        if stream:
            return gen()
        outs = [o async for o in gen()]
        by_id = {o["meta_info"]["id"]: o for o in outs}
        res = [by_id[s.rid] for s in states]
        return res if g.is_batch() or len(res) > 1 else res[0]

    def encode(self, prompt: Union[str, List[str], None] = None, image_data: Any = None, input_ids: Any = None) -> Any:
        g = ios.GenerateReqInput(text=prompt, input_ids=input_ids, image_data=image_data, sampling_params={"max_new_tokens": 1})
        states = [self._submit(s, None, is_embedding=True) for s in g.normalize()]
        self._drive(states)
        res = [s.result() for s in states]
        return res if g.is_batch() else res[0]

    def score(self, query: str, items: List[str], label_token_ids: List[int], apply_softmax: bool = True) -> List[List[float]]:
        """P(label token | query + item) for each item (reference tokenizer_manager_score_mixin)"""
        outs = self.generate([query + it for it in items], {"max_new_tokens": 1, "temperature": 0}, return_logprob=True,
                             top_logprobs_num=max(20, len(label_token_ids)))
        scores = []
        for o in outs:
            top = dict((tid, lp) for lp, tid in (o["meta_info"].get("output_top_logprobs") or [[]])[0])
            row = [top.get(t, -1e9) for t in label_token_ids]
            if apply_softmax:
                row = torch.softmax(torch.tensor(row), 0).tolist()
            scores.append(row)
        return scores

    def rerank(self, query: str, documents: List[str]) -> List[Dict[str, Any]]:
        if type(self.runner.model).__name__.endswith("ForSequenceClassification") and self.tokenizer is not None:
            # cross-encoder (bge-reranker, ms-marco-MiniLM, ...): ONE pass over "<cls> query <sep> document <sep>", the head's logit is the score
            # (reference CrossEncodingPooler / serving_rerank)
            tk = self.tokenizer
            try:
                pairs = [tk(query, d, truncation=True, max_length=self.model_config.max_position_embeddings - 2)["input_ids"] for d in documents]
            except TypeError:          # tokenizers without pair encoding: join with the separator ourselves
                sep = getattr(tk, "sep_token_id", None)
                pairs = [tk.encode(query) + ([sep] if sep is not None else []) + tk.encode(d)[1:] for d in documents]
            outs = self.encode(input_ids=pairs)
            sc = [float(o["embedding"][0]) for o in (outs if isinstance(outs, list) else [outs])]
            order = sorted(range(len(documents)), key=lambda i: -sc[i])
            return [{"index": i, "score": sc[i], "document": documents[i]} for i in order]
        embs = self.encode([query] + documents)
        q = torch.tensor(embs[0]["embedding"])
        sc = [float(torch.dot(q, torch.tensor(e["embedding"]))) for e in embs[1:]]
        order = sorted(range(len(documents)), key=lambda i: -sc[i])
        return [{"index": i, "score": sc[i], "document": documents[i]} for i in order]

    # ---- control plane
    def _control(self, msg: Any) -> Any:
        if self._bg_thread is not None:
            box: "queue.Queue[Any]" = queue.Queue()
            self._inbox.put((msg, box))
            self._wake.set()
            return box.get()
        return self.scheduler.handle_request(msg)

    def abort_request(self, rid: str = "", abort_all: bool = False) -> None:
        self._control(ios.AbortReq(rid=rid, abort_all=abort_all))

    def flush_cache(self) -> bool:
        return self._control(ios.FlushCacheReq()).success

    def take_kv_events(self) -> Dict[str, Any]:
        """KV-cache events since the last call (needs ``enable_kv_events``): feed for a KV-aware router / indexer"""
        r = self._control(ios.GetKvEventsReq())
        return {"events": r.events or [], "page_size": r.page_size}

    def get_server_info(self) -> Dict[str, Any]:
        st = self._control(ios.GetInternalStateReq()).state
        return {**self.server_args.to_dict(), "internal_states": [st], "version": "0.1.0", "weight_version": self.weight_version,
                "model_config": dataclasses.asdict(self.model_config)}

    def get_load(self) -> Dict[str, Any]:
        return dataclasses.asdict(self._control(ios.GetLoadReq()))

    def start_profile(self, output_dir: Optional[str] = None, num_steps: Optional[int] = None, activities=None,
                      profile_by_stage: bool = False) -> Any:
        return self._control(ios.ProfileReq("start", output_dir, num_steps, activities, profile_by_stage))

    def stop_profile(self) -> Any:
        return self._control(ios.ProfileReq("stop"))

    def update_weights_from_disk(self, model_path: str, load_format: Optional[str] = None) -> Any:
        r = self._control(ios.UpdateWeightsFromDiskReq(model_path, load_format))
        if r.success:
            self.weight_version = model_path
        return r

    def init_weights_update_group(self, master_address: str, master_port: int, rank_offset: int, world_size: int,
                                  group_name: str = "weight_update_group", backend: str = "nccl") -> Any:
        return self._control(ios.InitWeightsUpdateGroupReq(master_address, master_port, rank_offset, world_size, group_name, backend))

    def destroy_weights_update_group(self, group_name: str = "weight_update_group") -> Any:
        return self._control(ios.DestroyWeightsUpdateGroupReq(group_name))

    def update_weights_from_distributed(self, names: List[str], dtypes: List[str], shapes: List[List[int]],
                                        group_name: str = "weight_update_group", flush_cache: bool = True) -> Any:
        r = self._control(ios.UpdateWeightsFromDistributedReq(names, dtypes, shapes, group_name, flush_cache))
        if r.success:
            self.weight_version = f"distributed:{group_name}"
        return r

    def update_weights_from_tensor(self, named_tensors: Any, flush_cache: bool = True) -> Any:
        return self._control(ios.UpdateWeightsFromTensorReq(named_tensors, flush_cache))

    def get_weights_by_name(self, name: str, truncate_size: int = 100) -> Any:
        return self._control(ios.GetWeightsByNameReq(name, truncate_size)).parameter

    def send_weights_to_remote_instance(self, master_address: str, master_port: int, group_name: str = "remote_instance_load",
                                        backend: str = "nccl") -> Any:
        return self._control(ios.SendWeightsToRemoteInstanceReq(master_address, int(master_port), group_name, backend))

    def collective_rpc(self, method: str, **kwargs: Any) -> Any:
        """call ``method(**kwargs)`` on the scheduler (or its model runner) of EVERY rank (reference Engine.collective_rpc); returns this
        rank's result"""
        return self._control(ios.RpcReq(method, kwargs)).result

    def weights_checker(self, action: str = "checksum") -> Any:
        return self._control(ios.WeightsCheckerReq(action))

    def slow_down(self, forward_sleep_time: Optional[float]) -> Any:
        return self._control(ios.SlowDownReq(forward_sleep_time))

    def release_memory_occupation(self, tags: Optional[List[str]] = None) -> Any:
        return self._control(ios.ReleaseMemoryOccupationReq(tags))

    def resume_memory_occupation(self, tags: Optional[List[str]] = None) -> Any:
        return self._control(ios.ResumeMemoryOccupationReq(tags))

    def load_lora_adapter(self, lora_name: str, lora_path: str) -> Any:
        return self._control(ios.LoadLoRAAdapterReq(lora_name, lora_path))

    def unload_lora_adapter(self, lora_name: str) -> Any:
        return self._control(ios.UnloadLoRAAdapterReq(lora_name))

    def open_session(self, capacity_of_str_len: int = 0, session_id: Optional[str] = None) -> Optional[str]:
        return self._control(ios.OpenSessionReq(capacity_of_str_len, session_id)).session_id

    def close_session(self, session_id: str) -> None:
        self._control(ios.CloseSessionReq(session_id))

    def pause_generation(self, mode: str = "abort") -> Any:
        return self._control(ios.PauseGenerationReq(mode))

    def continue_generation(self) -> Any:
        return self._control(ios.ContinueGenerationReq())

    def save_sharded_model(self, path: str) -> None:
        from ..model_loader.loader import save_sharded_state

        save_sharded_state(self.runner.model, path, ps.get_state().rank)

    # ================================================================== background loop
    def start_background_loop(self) -> None:
        if self._bg_thread is not None:
            return
        self._bg_stop.clear()
        self._shm_bcast = None
        if self._world.world_size > 1 and self.server_args.nnodes == 1 and envs.SGLB_SHM_BROADCAST.get():
            # single node: per-iteration request fan-out over a shared-memory ring instead of a Gloo collective
            from ..parallel.shm_broadcast import ShmBroadcaster

            try:
                if self.is_driver:
                    self._shm_bcast = ShmBroadcaster(self._world.world_size - 1)
                    self._world.broadcast_pyobj(self._shm_bcast.handle(), src=0)
                else:
                    h = self._world.broadcast_pyobj(None, src=0)
                    self._shm_bcast = ShmBroadcaster.attach(h, ps.get_state().rank - 1) if h is not None else None
            except Exception as e:      # /dev/shm unavailable: keep the Gloo path
                logger.warning("shared-memory broadcast unavailable (%s); using Gloo", e)
                if self.is_driver:
                    self._world.broadcast_pyobj(None, src=0)
                self._shm_bcast = None
        self._bg_thread = threading.Thread(target=self._bg_loop, name="sglb-scheduler", daemon=True)
        self._bg_thread.start()
        a = self.server_args
        if getattr(self, "_watchdog", None) is None and a.watchdog_timeout and a.watchdog_timeout > 0:
            from ..utils.watchdog import SchedulerWatchdog

            # reference utils/watchdog.py: forward_ct must advance while work is pending; soft timeout logs, hard timeout dumps and stops
            self._watchdog = SchedulerWatchdog(self.scheduler, float(a.watchdog_timeout), a.soft_watchdog_timeout,
                                               interval=min(1.0, max(0.05, float(a.watchdog_timeout) / 10)), on_hard=self._on_watchdog_timeout).start()

    wedged = False
    exit_on_watchdog = False          # servers set this: a wedged scheduler thread cannot be recovered, the process tree goes down

    def _on_watchdog_timeout(self) -> None:
        self.wedged = True
        try:
            path = self.dump_requests_before_crash()
            logger.error("watchdog: in-flight requests dumped to %s", path)
        except Exception as e:
            logger.error("watchdog: crash dump failed: %s", e)
        if self.exit_on_watchdog:
            from ..utils.process_tree import kill_process_tree

            kids = [p for p in (getattr(self, "_p_sched", None), getattr(self, "_p_detok", None)) if p is not None] + list(getattr(self, "_peers", []) or [])
            kill_process_tree(kids, grace_s=1.0)
            os._exit(1)

    def _bg_loop(self) -> None:
        sch = self.scheduler
        tp = ps.get_state().tp
        if self.server_args.device == "cuda":
            torch.cuda.set_device(self.runner.device)
        while not self._bg_stop.is_set():
            new: List[Any] = []
            if self.is_driver:
                while True:
                    try:
                        new.append(self._inbox.get_nowait())
                    except queue.Empty:
                        break
            if self._world.world_size > 1:      # every rank of the TP x PP world runs the same loop on the same requests
                payload = [m if not isinstance(m, tuple) else m[0] for m in new] if self.is_driver else None
                if self._shm_bcast is not None:
                    payload = self._shm_bcast.broadcast(payload, timeout=None)
                else:
                    payload = self._world.broadcast_pyobj(payload, src=0)
                if not self.is_driver:
                    new = payload
            for m in new:
                try:
                    if isinstance(m, tuple):
                        msg, box = m
                        box.put(sch.handle_request(msg))
                    else:
                        sch.handle_request(m)
                except Exception as e:  # a bad request must not kill the loop
                    logger.exception("request handling failed: %s", e)
                    if isinstance(m, tuple):
                        m[1].put(ios.GenericResp(False, str(e)))
            if sch.has_work():
                try:
                    sch.step()
                    self._relay_sync()
                except Exception:
                    logger.exception("scheduler step failed; aborting in-flight requests")
                    self.dump_requests_before_crash()
                    self._fail_all()
                    raise
            else:
                sch.drain()
                self._relay_sync()
                self._wake.wait(timeout=0.002)
                self._wake.clear()

    def dump_requests_before_crash(self) -> Optional[str]:
        """``--crash-dump-folder``: write the in-flight and queued requests (ids, token ids, sampling params, progress) plus the scheduler's
        internal state to a JSON file for offline replay (reference tokenizer_manager.dump_requests_before_crash:2986)"""
        folder = self.server_args.crash_dump_folder
        if not folder:
            return None
        import json
        import os

        try:
            os.makedirs(folder, exist_ok=True)
            sch = self.scheduler
            reqs = []
            live = sch.live_reqs()          # running + chunked + the just-run prefill batch + waiting requests that hold KV
            seen = set()
            for where, lst in (("running", live), ("waiting", sch.waiting_queue), ("grammar", sch.grammar_queue)):
                for r in lst:
                    if id(r) in seen or r.finished():
                        continue
                    seen.add(id(r))
                    reqs.append({"rid": r.rid, "state": where, "input_ids": list(r.origin_input_ids), "output_ids": list(r.output_ids),
                                 "sampling_params": getattr(r.sampling_params, "__dict__", {}), "kv_len": int(r.kv_len),
                                 "retractions": int(r.retraction_count)})
            path = os.path.join(folder, f"crash_dump_{int(time.time())}_rank{ps.get_state().rank}.json")
            with open(path, "w") as f:
                json.dump({"time": time.time(), "server_args": self.server_args.to_dict(), "internal_state": sch.get_internal_state(),
                           "requests": reqs}, f, default=str)
            logger.error("crash dump with %d requests written to %s", len(reqs), path)
            return path
        except Exception:
            logger.exception("writing the crash dump failed")
            return None

    def _fail_all(self) -> None:
        with self._lock:
            for st in list(self.states.values()):
                st.finished = True
                st.meta["finish_reason"] = {"type": "abort", "message": "engine failure"}
                if st.out_q is not None:
                    st.out_q.put({**st.result(), "_done": True})
            self.states.clear()

    def shutdown(self) -> None:
        wd = getattr(self, "_watchdog", None)
        if wd is not None:
            wd.stop()
            self._watchdog = None
        self._bg_stop.set()
        self._wake.set()
        if self._bg_thread is not None:
            self._bg_thread.join(timeout=10)
            self._bg_thread = None
        self._release_device_memory()

    def _release_device_memory(self) -> None:
        """give the device memory back (KV pools, captured graphs, weights): a second engine created in the same process sizes its pools from
        the FREE memory and would otherwise end up with a pool of a handful of pages"""
        r = getattr(self, "runner", None)
        if r is None or not getattr(r, "is_cuda", False) or getattr(self, "_released", False):
            return
        self._released = True
        try:
            import gc

            import torch

            torch.cuda.synchronize(r.device)
            for name in ("graphs", "graphs_greedy", "prefill_graphs"):
                d = getattr(r, name, None)
                if isinstance(d, dict):
                    d.clear()
            for name in ("kv_pool", "page_table", "next_token_buf", "model", "lora_manager", "_stage_ring", "_res_ring"):
                if hasattr(r, name):
                    setattr(r, name, None)
            sch = getattr(self, "scheduler", None)
            if sch is not None:
                sch.worker = None
            gc.collect()
            torch.cuda.empty_cache()
        except Exception:      # noqa: BLE001 — best effort: shutdown must not raise
            pass

    def __enter__(self) -> "Engine":
        return self

    def __exit__(self, *a: Any) -> None:
        self.shutdown()
