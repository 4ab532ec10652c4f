"""Frontend-language IR.  Parity: reference ``python/sglang/lang/ir.py:141-619`` (SglFunction, SglGen, SglSelect,
SglImage, SglVideo, SglRoleBegin/End, SglFork, SglConcateAndAppend, SglExprList; ``SglSamplingParams``)."""
from __future__ import annotations

import dataclasses
import inspect
from typing import Any, Callable, Dict, List, Optional, Union


@dataclasses.dataclass
class SglSamplingParams:
    max_new_tokens: int = 128
    min_new_tokens: int = 0
    n: int = 1
    stop: Union[str, List[str]] = ()
    stop_token_ids: Optional[List[int]] = None
    temperature: float = 1.0
    top_p: float = 1.0
    top_k: int = -1
    min_p: float = 0.0
    frequency_penalty: float = 0.0
    presence_penalty: float = 0.0
    ignore_eos: bool = False
    return_logprob: Optional[bool] = None
    logprob_start_len: Optional[int] = None
    top_logprobs_num: Optional[int] = None
    json_schema: Optional[str] = None
    regex: Optional[str] = None
    ebnf: Optional[str] = None
    dtype: Optional[str] = None

    def clone(self) -> "SglSamplingParams":
        return dataclasses.replace(self)

    def to_srt_kwargs(self) -> Dict[str, Any]:
        d = {"max_new_tokens": self.max_new_tokens, "min_new_tokens": self.min_new_tokens, "n": self.n,
             "stop": list(self.stop) if not isinstance(self.stop, str) else self.stop, "stop_token_ids": self.stop_token_ids,
             "temperature": self.temperature, "top_p": self.top_p, "top_k": self.top_k, "min_p": self.min_p,
             "frequency_penalty": self.frequency_penalty, "presence_penalty": self.presence_penalty, "ignore_eos": self.ignore_eos}
        for k in ("json_schema", "regex", "ebnf"):
            if getattr(self, k) is not None:
                d[k] = getattr(self, k)
        return d


class SglExpr:
    node_ct = 0

    def __init__(self) -> None:
        self.node_id = SglExpr.node_ct
        SglExpr.node_ct += 1
        self.prev_node: Optional["SglExpr"] = None
        self.pid: Optional[int] = None

    def __add__(self, other: Any) -> "SglExprList":
        return self.concatenate_ir(self, other if isinstance(other, SglExpr) else SglConstantText(str(other)))

    def __radd__(self, other: Any) -> "SglExprList":
        return self.concatenate_ir(other if isinstance(other, SglExpr) else SglConstantText(str(other)), self)

    @staticmethod
    def concatenate_ir(a: "SglExpr", b: "SglExpr") -> "SglExprList":
        al = a.expr_list if isinstance(a, SglExprList) else [a]
        bl = b.expr_list if isinstance(b, SglExprList) else [b]
        return SglExprList(al + bl)


class SglExprList(SglExpr):
    def __init__(self, expr_list: List[SglExpr]):
        super().__init__()
        self.expr_list = expr_list


class SglConstantText(SglExpr):
    def __init__(self, value: str):
        super().__init__()
        self.value = value


class SglArgument(SglExpr):
    def __init__(self, name: str, value: Any):
        super().__init__()
        self.name, self.value = name, value


class SglGen(SglExpr):
    def __init__(self, name: Optional[str] = None, **kw: Any):
        super().__init__()
        self.name = name
        self.choices_hint = None
        self.sampling_params = SglSamplingParams(**{k: v for k, v in kw.items() if v is not None})


class SglSelect(SglExpr):
    def __init__(self, name: Optional[str], choices: List[str], temperature: float = 0.0, choices_method: Any = None):
        super().__init__()
        self.name, self.choices, self.temperature, self.choices_method = name, choices, temperature, choices_method


class SglImage(SglExpr):
    def __init__(self, path: Any):
        super().__init__()
        self.path = path


class SglVideo(SglExpr):
    def __init__(self, path: Any, num_frames: int):
        super().__init__()
        self.path, self.num_frames = path, num_frames


class SglRoleBegin(SglExpr):
    def __init__(self, role: str):
        super().__init__()
        self.role = role


class SglRoleEnd(SglExpr):
    def __init__(self, role: str):
        super().__init__()
        self.role = role


class SglVarScopeBegin(SglExpr):
    def __init__(self, name: str):
        super().__init__()
        self.name = name


class SglVarScopeEnd(SglExpr):
    def __init__(self, name: str):
        super().__init__()
        self.name = name


class SglConcateAndAppend(SglExpr):
    def __init__(self, states: List[Any]):
        super().__init__()
        self.states = states


class SglCommitLazy(SglExpr):
    pass


class SglSeparateReasoning(SglExpr):
    def __init__(self, model_type: str, expr: SglExpr):
        super().__init__()
        self.model_type, self.expr = model_type, expr


class SglFunction:
    def __init__(self, func: Callable, num_api_spec_tokens: Optional[int] = None, bind_arguments: Optional[Dict] = None):
        self.func = func
        self.num_api_spec_tokens = num_api_spec_tokens
        self.bind_arguments = bind_arguments or {}
        self.pin_prefix_rid = None
        spec = inspect.getfullargspec(func)
        assert spec.args and spec.args[0] == "s", 'the first argument of an sgl.function must be "s"'
        self.arg_names = spec.args[1:]
        self.arg_defaults = spec.defaults or ()

    def bind(self, **kwargs: Any) -> "SglFunction":
        return SglFunction(self.func, bind_arguments={**self.bind_arguments, **kwargs})

    def run(self, *args: Any, max_new_tokens: int = 128, stop=(), temperature: float = 1.0, top_p: float = 1.0, top_k: int = -1,
            min_p: float = 0.0, frequency_penalty: float = 0.0, presence_penalty: float = 0.0, ignore_eos: bool = False,
            return_logprob=None, logprob_start_len=None, top_logprobs_num=None, stream: bool = False, backend: Any = None,
            use_thread: bool = True, **kwargs: Any):
        from .interpreter import run_program

        sp = SglSamplingParams(max_new_tokens=max_new_tokens, stop=stop, temperature=temperature, top_p=top_p, top_k=top_k,
                               min_p=min_p, frequency_penalty=frequency_penalty, presence_penalty=presence_penalty,
                               ignore_eos=ignore_eos, return_logprob=return_logprob, logprob_start_len=logprob_start_len,
                               top_logprobs_num=top_logprobs_num)
        from .api import get_default_backend

        backend = backend or get_default_backend()
        kwargs = {**self.bind_arguments, **kwargs}
        return run_program(self, backend, args, kwargs, sp, stream, use_thread=use_thread)

    def run_batch(self, batch_kwargs: List[Dict[str, Any]], *, max_new_tokens: int = 128, stop=(), temperature: float = 1.0,
                  top_p: float = 1.0, top_k: int = -1, min_p: float = 0.0, ignore_eos: bool = False, backend: Any = None,
                  num_threads: Union[str, int] = "auto", progress_bar: bool = False, **_):
        from .api import get_default_backend
        from .interpreter import run_program_batch

        assert isinstance(batch_kwargs, (list, tuple))
        if not batch_kwargs:
            return []
        sp = SglSamplingParams(max_new_tokens=max_new_tokens, stop=stop, temperature=temperature, top_p=top_p, top_k=top_k,
                               min_p=min_p, ignore_eos=ignore_eos)
        backend = backend or get_default_backend()
        batch_kwargs = [{**self.bind_arguments, **kw} for kw in batch_kwargs]
        return run_program_batch(self, backend, batch_kwargs, sp, num_threads, progress_bar)

    def trace(self, *, backend: Any = None, **kwargs: Any):
        from .tracer import trace_program

        return trace_program(self, kwargs, backend)

    def cache(self, backend: Any = None):
        from .api import get_default_backend
        from .interpreter import cache_program

        return cache_program(self, backend or get_default_backend())

    def __call__(self, *args: Any, **kwargs: Any):
        from .tracer import TracingScope

        scope = TracingScope.get_current_scope()
        if scope is None:
            return self.run(*args, **kwargs)
        kwargs["backend"] = scope.tracer_state.backend
        return self.trace(*args, **kwargs)
