"""Interpreter for sgl programs.

Parity: reference ``python/sglang/lang/interpreter.py`` (``run_program:57``, ``run_program_batch:93``,
``StreamExecutor:274`` with its worker thread ``:422`` and ``_execute :461`` dispatch — fill ``:505``, gen ``:593``,
select ``:647``, image ``:524``, fork ``:370``; ``ProgramState:852`` with ``fork:888``, text / messages / variable
access and streaming iterators).
"""
from __future__ import annotations

import queue
import threading
import uuid
from concurrent.futures import ThreadPoolExecutor
from typing import Any, Dict, Iterator, List, Optional, Union

from .ir import (SglArgument, SglCommitLazy, SglConcateAndAppend, SglConstantText, SglExpr, SglExprList, SglFunction, SglGen,
                 SglImage, SglRoleBegin, SglRoleEnd, SglSamplingParams, SglSelect, SglSeparateReasoning, SglVarScopeBegin,
                 SglVarScopeEnd, SglVideo)


def run_program(program: SglFunction, backend: Any, func_args, func_kwargs: Dict[str, Any], default_sp: SglSamplingParams,
                stream: bool, sync: bool = False, use_thread: bool = True) -> "ProgramState":
    if hasattr(backend, "endpoint"):
        backend = backend.endpoint
    assert backend is not None, "set a backend with sgl.set_default_backend(...) or pass backend="
    stream_executor = StreamExecutor(backend, func_kwargs, default_sp, chat_template=None, stream=stream, use_thread=use_thread)
    state = ProgramState(stream_executor)
    if stream:
        t = threading.Thread(target=_run_internal, args=(state, program, func_args, func_kwargs, sync), daemon=True)
        t.start()
        return state
    _run_internal(state, program, func_args, func_kwargs, sync)
    return state


def _run_internal(state: "ProgramState", program: SglFunction, func_args, func_kwargs, sync: bool) -> None:
    try:
        state.ret_value = program.func(state, *func_args, **func_kwargs)
    except Exception as e:
        state.stream_executor.error_ = e
        raise
    finally:
        state.stream_executor.end()
    if sync:
        state.stream_executor.sync()


def run_program_batch(program: SglFunction, backend: Any, batch_arguments: List[Dict[str, Any]], default_sp: SglSamplingParams,
                      num_threads: Union[str, int], progress_bar: bool) -> List["ProgramState"]:
    if hasattr(backend, "endpoint"):
        backend = backend.endpoint
    if len(batch_arguments) > 1:
        cache_program(program, backend, batch_arguments[0])
    n = max(16, 64) if num_threads == "auto" else int(num_threads)
    n = min(n, len(batch_arguments))
    if n <= 1:
        return [run_program(program, backend, (), a, default_sp, False, True) for a in batch_arguments]
    with ThreadPoolExecutor(n) as ex:
        futs = [ex.submit(run_program, program, backend, (), a, default_sp, False, True) for a in batch_arguments]
        return [f.result() for f in futs]


def cache_program(program: SglFunction, backend: Any, example_args: Optional[Dict[str, Any]] = None) -> None:
    """warm the server's radix cache with the program's constant prefix (reference ``tracer.extract_prefix_by_tracing``)"""
    from .tracer import extract_prefix_by_tracing

    try:
        prefix = extract_prefix_by_tracing(program, backend, example_args)
    except Exception:
        prefix = ""
    if prefix and len(prefix) > 64 and hasattr(backend, "cache_prefix"):
        backend.cache_prefix(prefix)


class StreamExecutor:
    def __init__(self, backend: Any, arguments: Dict[str, Any], default_sampling_para: SglSamplingParams, chat_template: Any,
                 stream: bool, use_thread: bool = True):
        self.sid = uuid.uuid4().hex
        self.backend = backend
        self.arguments = arguments
        self.default_sampling_para = default_sampling_para
        self.stream = stream
        self.variables: Dict[str, Any] = {}
        self.variable_event: Dict[str, threading.Event] = {}
        self.meta_info: Dict[str, Any] = {}
        self.is_finished = False
        self.error_: Optional[Exception] = None
        self.text_ = ""
        self.messages_: List[Dict[str, Any]] = []
        self.chat_template = chat_template or backend.get_chat_template()
        self.cur_role: Optional[str] = None
        self.cur_role_begin_pos = 0
        self.images_: List[Any] = []
        self.cur_images: List[Any] = []
        self.fork_start_text_pos: Optional[int] = None
        self.stream_text_event = threading.Event() if stream else None
        self.stream_var_event: Dict[str, threading.Event] = {}
        self.use_thread = use_thread
        if use_thread:
            self.queue: "queue.Queue[Optional[SglExpr]]" = queue.Queue()
            self.worker = threading.Thread(target=self._thread_worker_func, daemon=True)
            self.worker.start()

    # ---- submission
    def submit(self, expr: SglExpr) -> None:
        self._init_var_event(expr)
        if self.use_thread:
            self.queue.put(expr)
        else:
            self._execute(expr)

    def sync(self) -> None:
        if self.use_thread:
            self.queue.join()

    def get_var(self, name: str) -> Any:
        if name in self.variable_event:
            self.variable_event[name].wait()
        if self.error_ is not None:
            raise self.error_
        return self.variables[name]

    def set_var(self, name: str, value: Any) -> None:
        self.variables[name] = value

    def get_meta_info(self, name: str, timeout: Optional[float] = None) -> Any:
        if name in self.variable_event and not self.variable_event[name].wait(timeout):
            return None
        return self.meta_info.get(name)

    def fork(self, size: int = 1, position_ids_offset: Optional[List[int]] = None) -> List["StreamExecutor"]:
        if size > 1:
            self.submit(SglCommitLazy())
        self.sync()
        exes = [StreamExecutor(self.backend, self.arguments, self.default_sampling_para, self.chat_template, self.stream,
                               self.use_thread) for _ in range(size)]
        for e in exes:
            e.variables = dict(self.variables)
            e.text_ = str(self.text_)
            e.messages_ = list(self.messages_)
            e.cur_role, e.cur_role_begin_pos = self.cur_role, self.cur_role_begin_pos
            e.fork_start_text_pos = len(self.text_)
            e.images_ = list(self.images_)
        return exes

    def text(self) -> str:
        self.sync()
        return self.text_

    def messages(self) -> List[Dict[str, Any]]:
        self.sync()
        return self.messages_

    def end(self) -> None:
        if self.use_thread:
            if self.worker.is_alive():
                self.queue.put(None)
        self.backend.end_program(self)

    def _thread_worker_func(self) -> None:
        error = None
        while True:
            expr = self.queue.get()
            if expr is None:
                self.queue.task_done()
                break
            try:
                if error is None:
                    self._execute(expr)
            except Exception as e:
                error = self.error_ = e
            self.queue.task_done()
            if self.stream_text_event:
                self.stream_text_event.set()
        if error is not None:
            for ev in self.variable_event.values():
                ev.set()
            for ev in self.stream_var_event.values():
                ev.set()
        if self.stream_text_event:
            self.stream_text_event.set()
        self.is_finished = True

    # ---- execution
    def _init_var_event(self, expr: SglExpr) -> None:
        if isinstance(expr, (SglGen, SglSelect, SglVarScopeBegin)):
            if expr.name is not None:
                self.variable_event[expr.name] = threading.Event()
                if self.stream:
                    self.stream_var_event[expr.name] = threading.Event()
        elif isinstance(expr, SglExprList):
            for e in expr.expr_list:
                self._init_var_event(e)
        elif isinstance(expr, SglSeparateReasoning):
            self._init_var_event(expr.expr)

    def _execute(self, other: Any) -> None:
        if isinstance(other, str):
            other = SglConstantText(other)
        if isinstance(other, SglConstantText):
            self._execute_fill(other.value)
        elif isinstance(other, SglGen):
            self._execute_gen(other)
        elif isinstance(other, SglSelect):
            self._execute_select(other)
        elif isinstance(other, SglExprList):
            for x in other.expr_list:
                self._execute(x)
        elif isinstance(other, SglRoleBegin):
            self._execute_role_begin(other)
        elif isinstance(other, SglRoleEnd):
            self._execute_role_end(other)
        elif isinstance(other, SglImage):
            self._execute_image(other)
        elif isinstance(other, SglVideo):
            self._execute_image(SglImage(other.path))
        elif isinstance(other, SglVarScopeBegin):
            self.variables[f"__scope_{other.name}"] = len(self.text_)
        elif isinstance(other, SglVarScopeEnd):
            start = self.variables.pop(f"__scope_{other.name}")
            self.variables[other.name] = self.text_[start:]
            self.variable_event[other.name].set()
        elif isinstance(other, SglCommitLazy):
            self.backend.commit_lazy_operations(self)
        elif isinstance(other, SglConcateAndAppend):
            for st in other.states:
                self._execute_fill(st.text()[st.stream_executor.fork_start_text_pos or 0:])
        elif isinstance(other, SglSeparateReasoning):
            self._execute_separate_reasoning(other)
        elif isinstance(other, SglArgument):
            self._execute_fill(str(other.value))
        else:
            raise ValueError(f"unknown expression type {type(other)}")

    def _execute_fill(self, value: str) -> None:
        self.text_ += str(value)

    def _execute_image(self, expr: SglImage) -> None:
        self.images_.append(expr.path)
        self.cur_images.append(expr.path)
        self.text_ += self.chat_template.image_token

    def _resolve_sp(self, sp: SglSamplingParams) -> SglSamplingParams:
        out = self.default_sampling_para.clone()
        defaults = SglSamplingParams()
        for f in sp.__dataclass_fields__:
            v = getattr(sp, f)
            if v != getattr(defaults, f):
                setattr(out, f, v)
        return out

    def _execute_gen(self, expr: SglGen) -> None:
        sp = self._resolve_sp(expr.sampling_params)
        name = expr.name
        if not self.stream:
            comp, meta = self.backend.generate(self, sampling_params=sp)
            if isinstance(comp, list):   # n > 1
                self.text_ += comp[0]
            else:
                self.text_ += comp
            val = comp
            if sp.dtype in ("int", int) and isinstance(comp, str):
                try:
                    val = int(comp.strip().split()[0])
                except Exception:
                    val = comp
            self.variables[name] = val
            self.meta_info[name] = meta
            if name in self.variable_event:
                self.variable_event[name].set()
        else:
            self.variables[name] = ""
            self.stream_var_event[name].set()
            for comp, meta in self.backend.generate_stream(self, sampling_params=sp):
                self.text_ += comp
                self.variables[name] += comp
                self.meta_info[name] = meta
                self.stream_var_event[name].set()
                self.stream_text_event.set()
            self.variable_event[name].set()
            self.stream_var_event[name].set()

    def _execute_select(self, expr: SglSelect) -> None:
        decision = self.backend.select(self, expr.choices, expr.temperature, expr.choices_method)
        if expr.name is not None:
            self.variables[expr.name] = decision.decision
            self.meta_info[expr.name] = decision.meta_info
            self.variable_event[expr.name].set()
            if self.stream_var_event:
                self.stream_var_event[expr.name].set()
        self.text_ += decision.decision

    def _execute_role_begin(self, expr: SglRoleBegin) -> None:
        assert self.cur_role is None, "nested roles are not allowed"
        if len(self.messages_) == 0 and expr.role != "system":
            default_system = self.chat_template.default_system_prompt
            if default_system:
                self._execute_role_begin(SglRoleBegin("system"))
                self._execute_fill(default_system)
                self._execute_role_end(SglRoleEnd("system"))
        self.cur_role = expr.role
        prefix, _ = self.chat_template.get_prefix_and_suffix(expr.role, self.messages_)
        self._execute_fill(prefix)
        self.cur_role_begin_pos = len(self.text_)

    def _execute_role_end(self, expr: SglRoleEnd) -> None:
        new_text = self.text_[self.cur_role_begin_pos:].lstrip()
        _, suffix = self.chat_template.get_prefix_and_suffix(expr.role, self.messages_)
        self._execute_fill(suffix)
        if self.cur_images:
            content: Any = [{"type": "text", "text": new_text}] + [{"type": "image_url", "image_url": {"url": p}} for p in self.cur_images]
            self.cur_images = []
        else:
            content = new_text
        self.messages_.append({"role": expr.role, "content": content})
        self.cur_role = None

    def _execute_separate_reasoning(self, expr: SglSeparateReasoning) -> None:
        from ..parser.reasoning_parser import ReasoningParser

        inner = expr.expr
        self._execute(inner)
        names = [inner.name] if isinstance(inner, (SglGen, SglSelect)) else [e.name for e in getattr(inner, "expr_list", []) if hasattr(e, "name")]
        for n in names:
            if n is None:
                continue
            r, c = ReasoningParser(expr.model_type).parse_non_stream(self.variables[n])
            self.variables[f"{n}_reasoning_content"] = r or ""
            self.variables[n] = c


class ProgramState:
    def __init__(self, stream_executor: StreamExecutor):
        self.stream_executor = stream_executor
        self.ret_value: Any = None

    # ---- role context managers
    class _Role:
        def __init__(self, state: "ProgramState", role: str):
            self.s, self.role = state, role

        def __enter__(self):
            self.s.stream_executor.submit(SglRoleBegin(self.role))
            return self.s

        def __exit__(self, *a):
            self.s.stream_executor.submit(SglRoleEnd(self.role))

    def system(self, expr: Any = None):
        return self._role_common("system", expr)

    def user(self, expr: Any = None):
        return self._role_common("user", expr)

    def assistant(self, expr: Any = None):
        return self._role_common("assistant", expr)

    def _role_common(self, role: str, expr: Any):
        if expr is None:
            return ProgramState._Role(self, role)
        self.stream_executor.submit(SglExprList([SglRoleBegin(role), expr if isinstance(expr, SglExpr) else SglConstantText(str(expr)),
                                                 SglRoleEnd(role)]))
        return self

    def var_scope(self, name: str):
        s = self

        class _Scope:
            def __enter__(self_inner):
                s.stream_executor.submit(SglVarScopeBegin(name))

            def __exit__(self_inner, *a):
                s.stream_executor.submit(SglVarScopeEnd(name))

        return _Scope()

    def fork(self, size: int = 1, position_ids_offset: Optional[List[int]] = None) -> "ProgramStateGroup":
        return ProgramStateGroup([ProgramState(x) for x in self.stream_executor.fork(size, position_ids_offset)], self)

    def copy(self, position_ids_offset=None) -> "ProgramState":
        return self.fork(1, position_ids_offset)[0]

    def text(self) -> str:
        return self.stream_executor.text()

    def messages(self) -> List[Dict[str, Any]]:
        return self.stream_executor.messages()

    def sync(self) -> None:
        self.stream_executor.sync()

    def error(self) -> Optional[Exception]:
        return self.stream_executor.error_

    def text_iter(self, var_name: Optional[str] = None) -> Iterator[str]:
        se = self.stream_executor
        assert se.stream, "text_iter requires run(..., stream=True)"
        prev = 0
        if var_name is None:
            while True:
                se.stream_text_event.wait()
                se.stream_text_event.clear()
                out = str(se.text_[prev:])
                prev += len(out)
                if out:
                    yield out
                if se.is_finished:
                    out = str(se.text_[prev:])
                    if out:
                        yield out
                    break
        else:
            while var_name not in se.stream_var_event:
                if se.is_finished:
                    return
                threading.Event().wait(0.001)
            ev = se.stream_var_event[var_name]
            while True:
                ev.wait()
                ev.clear()
                out = str(se.variables.get(var_name, "")[prev:])
                prev += len(out)
                if out:
                    yield out
                if se.variable_event[var_name].is_set():
                    out = str(se.variables.get(var_name, "")[prev:])
                    if out:
                        yield out
                    break

    def get_var(self, name: str) -> Any:
        return self.stream_executor.get_var(name)

    def set_var(self, name: str, value: Any) -> None:
        self.stream_executor.set_var(name, value)

    def get_meta_info(self, name: str) -> Any:
        return self.stream_executor.get_meta_info(name)

    def __iadd__(self, other: Any) -> "ProgramState":
        if other is None:
            raise ValueError("tried to append None to the state")
        self.stream_executor.submit(other if isinstance(other, SglExpr) else SglConstantText(str(other)))
        return self

    def __getitem__(self, name: str) -> Any:
        return self.get_var(name)

    def __setitem__(self, name: str, value: Any) -> None:
        self.set_var(name, value)

    def __contains__(self, name: str) -> bool:
        return name in self.stream_executor.variables

    def __repr__(self) -> str:
        return f"ProgramState({self.text()})"


class ProgramStateGroup:
    def __init__(self, states: List[ProgramState], src_state: Optional[ProgramState] = None):
        self.states, self.src_state = states, src_state

    def join(self, mode: str = "gather_variable") -> None:
        if mode == "gather_variable":
            src_vars = self.src_state.stream_executor.variables
            src_events = self.src_state.stream_executor.variable_event
            for s in self.states:
                s.stream_executor.sync()
            keys = set().union(*[set(s.stream_executor.variables) for s in self.states])
            for k in keys:
                vals = [s.stream_executor.variables.get(k) for s in self.states]
                src_vars[k] = vals
                if k in src_events:
                    src_events[k].set()
        elif mode == "concate_and_append":
            self.src_state.stream_executor.submit(SglConcateAndAppend(self.states))
        else:
            raise ValueError(mode)
        for s in self.states:
            s.stream_executor.end()

    def __getitem__(self, i: int) -> ProgramState:
        return self.states[i]

    def __setitem__(self, i: int, v: ProgramState) -> None:
        pass

    def __iadd__(self, other: Any) -> "ProgramStateGroup":
        if isinstance(other, (list, tuple)):
            for s, o in zip(self.states, other):
                s += o
        elif callable(other) and not isinstance(other, SglExpr):
            for i, s in enumerate(self.states):
                s += other(i)
        else:
            for s in self.states:
                s += other
        return self

    def __len__(self) -> int:
        return len(self.states)
