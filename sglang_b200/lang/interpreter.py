"""Executor for ``sgl`` programs: a deferred op log that is evaluated on demand.

Public surface (what programs and ``SglFunction.run / run_batch`` see) matches the reference frontend — ``ProgramState`` with ``+=``,
role helpers, ``fork / join``, ``text() / messages()``, ``state["var"]``, ``text_iter`` — see reference ``python/sglang/lang/interpreter.py``
(``run_program:57``, ``run_program_batch:93``, ``ProgramState:852``).

The machinery underneath is different by design.  The reference gives every program state a worker thread, a queue and one
``threading.Event`` per variable.  Here a state is a plain value object:

  * ``state += expr`` only APPENDS to the state's op log (``_Pending``) — nothing runs;
  * a *demand* (reading a variable, ``text()``, ``fork``, the end of a non-streaming ``run``) drains the log in order on the calling
    thread (``_drain``): constant text is concatenated into the ``Transcript``, ``gen`` / ``select`` call the backend;
  * ``stream=True`` needs no thread either: evaluation is written as generators (``_steps`` -> ``_eval`` -> the backend's chunk stream), and
    ``text_iter`` simply iterates them — every ``next()`` advances the program until the next piece of text exists (pull-based streaming);
  * concurrency exists exactly where requests can overlap on the server: the branches of a ``fork`` (and the programs of ``run_batch``) are
    drained together on one shared pool (``_POOL``) when they are joined, so N branches put N requests in flight at once.

Backends receive the state itself and read ``state.prompt`` / ``state.images`` (``lang/backend/runtime_endpoint.py``).
"""
from __future__ import annotations

import itertools
import threading
from collections import deque
from concurrent.futures import ThreadPoolExecutor
from typing import Any, Deque, Dict, Iterator, List, Optional, Tuple, Union

from .ir import (SglArgument, SglCommitLazy, SglConcateAndAppend, SglConstantText, SglExpr, SglExprList, SglFunction, SglGen,
                 SglImage, SglRoleBegin, SglRoleEnd, SglSamplingParams, SglSelect, SglSeparateReasoning, SglVarScopeBegin,
                 SglVarScopeEnd, SglVideo)

_POOL = ThreadPoolExecutor(max_workers=64, thread_name_prefix="sgl-prog")
_ids = itertools.count()


class Transcript:
    """what the model will see: text pieces, image inputs and the chat-message view of the same content"""

    __slots__ = ("pieces", "images", "turn_images", "messages", "open_role", "role_mark")

    def __init__(self) -> None:
        self.pieces: List[str] = []
        self.images: List[Any] = []
        self.turn_images: List[Any] = []
        self.messages: List[Dict[str, Any]] = []
        self.open_role: Optional[str] = None
        self.role_mark = 0              # len(text) where the open role's own content starts

    def text(self) -> str:
        if len(self.pieces) > 1:        # collapse so repeated reads stay O(total)
            self.pieces = ["".join(self.pieces)]
        return self.pieces[0] if self.pieces else ""

    def __len__(self) -> int:
        return sum(len(p) for p in self.pieces)

    def write(self, s: str) -> None:
        if s:
            self.pieces.append(s)

    def branch(self) -> "Transcript":
        t = Transcript()
        t.pieces = [self.text()]
        t.images, t.turn_images = list(self.images), list(self.turn_images)
        t.messages = [dict(m) for m in self.messages]
        t.open_role, t.role_mark = self.open_role, self.role_mark
        return t


class _Ctx:
    """per-run constants shared by a program state and all of its forks"""

    __slots__ = ("backend", "arguments", "defaults", "template", "stream")

    def __init__(self, backend: Any, arguments: Dict[str, Any], defaults: SglSamplingParams, stream: bool):
        self.backend, self.arguments, self.defaults, self.stream = backend, arguments, defaults, stream
        self.template = backend.get_chat_template()


def _unwrap_backend(backend: Any) -> Any:
    backend = getattr(backend, "endpoint", backend)
    assert backend is not None, "set a backend with sgl.set_default_backend(...) or pass backend="
    return backend


# ------------------------------------------------------------------------------------------------------------------ entry points
def run_program(program: SglFunction, backend: Any, func_args, func_kwargs: Dict[str, Any], default_sp: SglSamplingParams,
                stream: bool, sync: bool = False, use_thread: bool = True) -> "ProgramState":
    """build the op log by running the user function, then (unless streaming) evaluate it.  ``use_thread`` is accepted for API
    compatibility; there is no per-program thread to turn off."""
    state = ProgramState(_Ctx(_unwrap_backend(backend), func_kwargs, default_sp, stream), Transcript())
    try:
        state.ret_value = program.func(state, *func_args, **func_kwargs)
        if not stream:
            state._drain()
    except Exception as e:
        state._fail(e)
        raise
    finally:
        if not stream:
            state._close()
    return state


def run_program_batch(program: SglFunction, backend: Any, batch_arguments: List[Dict[str, Any]], default_sp: SglSamplingParams,
                      num_threads: Union[str, int], progress_bar: bool) -> List["ProgramState"]:
    backend = _unwrap_backend(backend)
    if len(batch_arguments) > 1:
        cache_program(program, backend, batch_arguments[0])
    width = len(batch_arguments) if num_threads == "auto" else max(1, min(int(num_threads), len(batch_arguments)))
    one = lambda a: run_program(program, backend, (), a, default_sp, False, True)      # noqa: E731
    if width <= 1:
        return [one(a) for a in batch_arguments]
    gate = threading.Semaphore(width)

    def gated(a):
        with gate:
            return one(a)

    return [f.result() for f in [_POOL.submit(gated, a) for a in batch_arguments]]


def cache_program(program: SglFunction, backend: Any, example_args: Optional[Dict[str, Any]] = None) -> None:
    """warm the server's radix cache with the program's constant prefix (reference ``tracer.extract_prefix_by_tracing``)"""
    from .tracer import extract_prefix_by_tracing

    try:
        prefix = extract_prefix_by_tracing(program, backend, example_args)
    except Exception:
        prefix = ""
    if prefix and len(prefix) > 64 and hasattr(backend, "cache_prefix"):
        backend.cache_prefix(prefix)


# ------------------------------------------------------------------------------------------------------------------ the state
class _Pending:
    """the deferred op log of one state (appends from the program, pops from whoever demands a value)"""

    __slots__ = ("ops", "lock")

    def __init__(self) -> None:
        self.ops: Deque[Any] = deque()
        self.lock = threading.RLock()


def _flatten(expr: Any) -> Iterator[Any]:
    if isinstance(expr, SglExprList):
        for e in expr.expr_list:
            yield from _flatten(e)
    else:
        yield expr


class ProgramState:
    def __init__(self, ctx: _Ctx, transcript: Transcript, variables: Optional[Dict[str, Any]] = None, fork_mark: Optional[int] = None):
        self.sid = f"prog-{next(_ids)}"
        self._ctx = ctx
        self._tr = transcript
        self._vars: Dict[str, Any] = dict(variables or {})
        self._meta: Dict[str, Any] = {}
        self._log = _Pending()
        self._declared: set = set()          # variable names that WILL exist once the log is drained
        self._scopes: Dict[str, int] = {}
        self._err: Optional[BaseException] = None
        self._closed = False
        self._fork_mark = fork_mark          # len(text) at the fork point (join(mode="concate_and_append") appends what came after)
        self.ret_value: Any = None

    # ---- what a backend reads
    @property
    def prompt(self) -> str:
        return self._tr.text()

    @property
    def images(self) -> List[Any]:
        return self._tr.images

    # ---- building the log
    def __iadd__(self, other: Any) -> "ProgramState":
        if other is None:
            raise ValueError("tried to append None to the state")
        expr = other if isinstance(other, SglExpr) else SglConstantText(str(other))
        with self._log.lock:
            for e in _flatten(expr):
                inner = e.expr if isinstance(e, SglSeparateReasoning) else e
                for leaf in _flatten(inner):
                    name = getattr(leaf, "name", None)
                    if name is not None and isinstance(leaf, (SglGen, SglSelect, SglVarScopeBegin)):
                        self._declared.add(name)
                self._log.ops.append(e)
        return self

    def _role_ops(self, role: str, expr: Any) -> Any:
        if expr is None:
            return _RoleBlock(self, role)
        self += SglExprList([SglRoleBegin(role), expr if isinstance(expr, SglExpr) else SglConstantText(str(expr)), SglRoleEnd(role)])
        return self

    def system(self, expr: Any = None):
        return self._role_ops("system", expr)

    def user(self, expr: Any = None):
        return self._role_ops("user", expr)

    def assistant(self, expr: Any = None):
        return self._role_ops("assistant", expr)

    def var_scope(self, name: str) -> "_ScopeBlock":
        return _ScopeBlock(self, name)

    # ---- evaluation
    def _fail(self, e: BaseException) -> None:
        self._err = e

    def _close(self) -> None:
        if not self._closed:
            self._closed = True
            try:
                self._ctx.backend.end_program(self)
            except Exception:
                pass

    def _steps(self, until: Optional[str] = None, live: bool = False) -> Iterator[Tuple[str, Optional[str]]]:
        """evaluate pending ops in order, yielding ``(piece, variable)`` for every piece of text the moment it is appended (``variable``
        names the gen / select it belongs to).  Stops early once variable ``until`` exists.  ``live``: ask the backend for chunk streams."""
        while True:
            if self._err is not None:
                raise self._err
            if until is not None and until in self._vars:
                return
            with self._log.lock:
                if not self._log.ops:
                    return
                op = self._log.ops.popleft()
            try:
                yield from self._eval(op, live)
            except Exception as e:
                self._fail(e)
                raise

    def _drain(self, until: Optional[str] = None) -> None:
        for _ in self._steps(until):
            pass

    def _emit(self, piece: str, var: Optional[str] = None) -> Iterator[Tuple[str, Optional[str]]]:
        self._tr.write(piece)
        if piece:
            yield piece, var

    def _eval(self, op: Any, live: bool) -> Iterator[Tuple[str, Optional[str]]]:
        tr, tpl = self._tr, self._ctx.template
        if isinstance(op, SglConstantText):
            yield from self._emit(op.value)
        elif isinstance(op, SglArgument):
            yield from self._emit(str(op.value))
        elif isinstance(op, SglGen):
            yield from self._eval_gen(op, live)
        elif isinstance(op, SglSelect):
            d = self._ctx.backend.select(self, op.choices, op.temperature, op.choices_method)
            if op.name is not None:
                self._vars[op.name], self._meta[op.name] = d.decision, d.meta_info
            yield from self._emit(d.decision, op.name)
        elif isinstance(op, SglRoleBegin):
            if tr.open_role is not None:
                raise ValueError("nested roles are not allowed")
            if not tr.messages and op.role != "system" and tpl.default_system_prompt:
                for sub in (SglRoleBegin("system"), SglConstantText(tpl.default_system_prompt), SglRoleEnd("system")):
                    yield from self._eval(sub, live)
            yield from self._emit(tpl.get_prefix_and_suffix(op.role, tr.messages)[0])
            tr.open_role, tr.role_mark = op.role, len(tr)
        elif isinstance(op, SglRoleEnd):
            body = tr.text()[tr.role_mark:].lstrip()
            yield from self._emit(tpl.get_prefix_and_suffix(op.role, tr.messages)[1])
            content: Any = body
            if tr.turn_images:
                content = [{"type": "text", "text": body}] + [{"type": "image_url", "image_url": {"url": p}} for p in tr.turn_images]
                tr.turn_images = []
            tr.messages.append({"role": op.role, "content": content})
            tr.open_role = None
        elif isinstance(op, (SglImage, SglVideo)):
            tr.images.append(op.path)
            tr.turn_images.append(op.path)
            yield from self._emit(tpl.image_token)
        elif isinstance(op, SglVarScopeBegin):
            self._scopes[op.name] = len(tr)
        elif isinstance(op, SglVarScopeEnd):
            self._vars[op.name] = tr.text()[self._scopes.pop(op.name):]
        elif isinstance(op, SglCommitLazy):
            self._ctx.backend.commit_lazy_operations(self)
        elif isinstance(op, SglConcateAndAppend):
            for child in op.states:
                child._drain()
                yield from self._emit(child._tr.text()[child._fork_mark or 0:])
        elif isinstance(op, SglSeparateReasoning):
            yield from self._eval_reasoning(op, live)
        else:
            raise ValueError(f"unknown expression type {type(op)}")

    def _sampling(self, sp: SglSamplingParams) -> SglSamplingParams:
        """per-call parameters override the run's defaults field by field (a field left at its dataclass default does not)"""
        merged, blank = self._ctx.defaults.clone(), SglSamplingParams()
        for f in sp.__dataclass_fields__:
            if getattr(sp, f) != getattr(blank, f):
                setattr(merged, f, getattr(sp, f))
        return merged

    def _eval_gen(self, op: SglGen, live: bool) -> Iterator[Tuple[str, Optional[str]]]:
        sp, name, be = self._sampling(op.sampling_params), op.name, self._ctx.backend
        if live:
            acc: List[str] = []
            meta: Any = None
            for piece, meta in be.generate_stream(self, sampling_params=sp):
                acc.append(piece)
                yield from self._emit(piece, name)
            self._vars[name], self._meta[name] = "".join(acc), meta       # the variable exists once its stream is complete
            return
        out, meta = be.generate(self, sampling_params=sp)
        yield from self._emit(out[0] if isinstance(out, list) else out, name)      # n > 1: the first sample continues the text
        val: Any = out
        if sp.dtype in ("int", int) and isinstance(out, str):
            try:
                val = int(out.strip().split()[0])
            except Exception:
                val = out
        self._vars[name], self._meta[name] = val, meta

    def _eval_reasoning(self, op: SglSeparateReasoning, live: bool) -> Iterator[Tuple[str, Optional[str]]]:
        from ..parser.reasoning_parser import ReasoningParser

        leaves = list(_flatten(op.expr))
        for leaf in leaves:
            yield from self._eval(leaf, live)
        for leaf in leaves:
            n = getattr(leaf, "name", None)
            if n is not None and isinstance(leaf, (SglGen, SglSelect)):
                thought, answer = ReasoningParser(op.model_type).parse_non_stream(self._vars[n])
                self._vars[f"{n}_reasoning_content"] = thought or ""
                self._vars[n] = answer

    # ---- reading results (each read is a demand)
    def sync(self) -> None:
        self._drain()

    def text(self) -> str:
        self._drain()
        return self._tr.text()

    def messages(self) -> List[Dict[str, Any]]:
        self._drain()
        return self._tr.messages

    def error(self) -> Optional[BaseException]:
        return self._err

    def get_var(self, name: str) -> Any:
        if name not in self._vars:
            self._drain(until=name)
        if self._err is not None:
            raise self._err
        return self._vars[name]

    def set_var(self, name: str, value: Any) -> None:
        self._vars[name] = value

    def get_meta_info(self, name: str, timeout: Optional[float] = None) -> Any:
        if name not in self._meta and name in self._declared:
            self._drain(until=name)
        return self._meta.get(name)

    def __getitem__(self, name: str) -> Any:
        return self.get_var(name)

    def __setitem__(self, name: str, value: Any) -> None:
        self.set_var(name, value)

    def __contains__(self, name: str) -> bool:
        return name in self._vars or name in self._declared

    def __repr__(self) -> str:
        return f"ProgramState({self.text()})"

    def text_iter(self, var_name: Optional[str] = None) -> Iterator[str]:
        """pull-based streaming: iterating this generator IS what runs the program's remaining ops — each ``next()`` advances the op log
        until the next piece of text exists (for a ``gen`` that is the next chunk of the backend's stream).  Text that was already
        produced, because the program itself demanded a value earlier, is yielded first."""
        assert self._ctx.stream, "text_iter requires run(..., stream=True)"
        if var_name is None:
            done = self._tr.text()
            if done:
                yield done
        elif var_name in self._vars:
            if self._vars[var_name]:
                yield str(self._vars[var_name])
            return
        try:
            for piece, var in self._steps(until=var_name, live=True):
                if var_name is None or var == var_name:
                    yield piece
        finally:
            if var_name is None and not self._log.ops:
                self._close()

    # ---- fork / join
    def fork(self, size: int = 1, position_ids_offset: Optional[List[int]] = None) -> "ProgramStateGroup":
        if size > 1:
            self += SglCommitLazy()
        self._drain()
        mark = len(self._tr)
        kids = [ProgramState(self._ctx, self._tr.branch(), self._vars, fork_mark=mark) for _ in range(size)]
        return ProgramStateGroup(kids, self)

    def copy(self, position_ids_offset=None) -> "ProgramState":
        return self.fork(1, position_ids_offset)[0]


class _RoleBlock:
    """``with s.user(): ...``"""

    def __init__(self, state: ProgramState, role: str):
        self.state, self.role = state, role

    def __enter__(self) -> ProgramState:
        self.state += SglRoleBegin(self.role)
        return self.state

    def __exit__(self, *exc) -> None:
        self.state += SglRoleEnd(self.role)


class _ScopeBlock:
    """``with s.var_scope("name"): ...`` captures the text produced inside the block as a variable"""

    def __init__(self, state: ProgramState, name: str):
        self.state, self.name = state, name

    def __enter__(self) -> None:
        self.state += SglVarScopeBegin(self.name)

    def __exit__(self, *exc) -> None:
        self.state += SglVarScopeEnd(self.name)


class ProgramStateGroup:
    def __init__(self, states: List[ProgramState], src_state: Optional[ProgramState] = None):
        self.states, self.src_state = states, src_state

    def _drain_all(self) -> None:
        """all branches at once: each pending ``gen`` becomes an in-flight request, the server batches them"""
        if len(self.states) == 1:
            self.states[0]._drain()
            return
        for f in [_POOL.submit(s._drain) for s in self.states]:
            f.result()

    def join(self, mode: str = "gather_variable") -> None:
        if mode == "gather_variable":
            self._drain_all()
            names = set().union(*[set(s._vars) for s in self.states]) if self.states else set()
            for k in names:
                self.src_state._vars[k] = [s._vars.get(k) for s in self.states]
        elif mode == "concate_and_append":
            self._drain_all()
            self.src_state += SglConcateAndAppend(self.states)
        else:
            raise ValueError(mode)
        for s in self.states:
            s._close()

    def __getitem__(self, i: int) -> ProgramState:
        return self.states[i]

    def __setitem__(self, i: int, v: ProgramState) -> None:       # ``forks[i] += x`` re-assigns the same object
        pass

    def __iter__(self) -> Iterator[ProgramState]:
        return iter(self.states)

    def __len__(self) -> int:
        return len(self.states)

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
