"""Program tracing: run an sgl function symbolically to extract its constant prefix (reference
``python/sglang/lang/tracer.py:29 extract_prefix_by_tracing``) so run_batch can pre-cache it in the radix tree."""
from __future__ import annotations

from typing import Any, Dict, Optional

from .ir import SglArgument, SglConstantText, SglExpr, SglExprList, SglFunction, SglRoleBegin, SglRoleEnd


class _StopTracing(Exception):
    pass


class TracingScope:
    cur_scope: Optional["TracingScope"] = None

    def __init__(self, tracer_state: Any):
        self.tracer_state = tracer_state
        self.last_scope = TracingScope.cur_scope

    def __enter__(self):
        TracingScope.cur_scope = self
        return self

    def __exit__(self, *a):
        TracingScope.cur_scope = self.last_scope

    @staticmethod
    def get_current_scope() -> Optional["TracingScope"]:
        return TracingScope.cur_scope


class _PrefixTracer:
    """collects leading constant text; the first non-constant node stops the trace"""

    def __init__(self, backend: Any):
        self.backend = backend
        self.text = ""
        self.template = backend.get_chat_template() if backend is not None else None
        self.messages = []

    def _emit(self, e: Any) -> None:
        if isinstance(e, str):
            self.text += e
        elif isinstance(e, SglConstantText):
            self.text += e.value
        elif isinstance(e, SglExprList):
            for x in e.expr_list:
                self._emit(x)
        elif isinstance(e, SglRoleBegin) and self.template is not None:
            self.text += self.template.get_prefix_and_suffix(e.role, self.messages)[0]
        elif isinstance(e, SglRoleEnd) and self.template is not None:
            self.text += self.template.get_prefix_and_suffix(e.role, self.messages)[1]
        else:
            raise _StopTracing()

    def __iadd__(self, other: Any):
        self._emit(other)
        return self

    def system(self, expr=None):
        return self._role("system", expr)

    def user(self, expr=None):
        return self._role("user", expr)

    def assistant(self, expr=None):
        return self._role("assistant", expr)

    def _role(self, role, expr):
        if expr is None:
            raise _StopTracing()
        self._emit(SglExprList([SglRoleBegin(role), expr if isinstance(expr, SglExpr) else SglConstantText(str(expr)), SglRoleEnd(role)]))
        return self

    def __getattr__(self, name):
        raise _StopTracing()


class _Symbol:
    """placeholder for a runtime argument: any use of it ends the constant prefix"""

    def __str__(self):
        raise _StopTracing()

    __format__ = lambda self, spec: (_ for _ in ()).throw(_StopTracing())


def extract_prefix_by_tracing(program: SglFunction, backend: Any, example_args: Optional[Dict[str, Any]] = None) -> str:
    tracer = _PrefixTracer(backend)
    kwargs = {n: _Symbol() for n in program.arg_names}
    kwargs.update(program.bind_arguments)
    try:
        with TracingScope(tracer):
            program.func(tracer, **kwargs)
    except _StopTracing:
        pass
    except Exception:
        pass
    return tracer.text


def trace_program(program: SglFunction, arguments: Dict[str, Any], backend: Any):
    t = _PrefixTracer(backend)
    try:
        program.func(t, **arguments)
    except _StopTracing:
        pass
    return t
