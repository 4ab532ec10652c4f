"""Public frontend API (reference ``python/sglang/lang/api.py:23-292`` / ``python/sglang/__init__.py:25-108``)."""
from __future__ import annotations

import re
from typing import Any, Callable, List, Optional

from .choices import (ChoicesSamplingMethod, greedy_token_selection, token_length_normalized,  # noqa: F401
                      unconditional_likelihood_normalized)
from .ir import (SglExpr, SglExprList, SglFunction, SglGen, SglImage, SglRoleBegin, SglRoleEnd, SglSelect, SglSeparateReasoning,
                 SglVideo)

_DEFAULT_BACKEND = {"v": None}


def function(func: Optional[Callable] = None, num_api_spec_tokens: Optional[int] = None):
    if func is not None:
        return SglFunction(func, num_api_spec_tokens=num_api_spec_tokens)
    return lambda f: SglFunction(f, num_api_spec_tokens=num_api_spec_tokens)


def set_default_backend(backend: Any) -> None:
    _DEFAULT_BACKEND["v"] = backend


def get_default_backend() -> Any:
    return _DEFAULT_BACKEND["v"]


def flush_cache(backend: Any = None) -> bool:
    backend = backend or get_default_backend()
    return bool(backend and backend.flush_cache())


def get_server_info(backend: Any = None) -> Any:
    backend = backend or get_default_backend()
    return backend.get_server_info() if backend else None


def gen(name: Optional[str] = None, max_tokens: Optional[int] = None, min_tokens: Optional[int] = None, n: Optional[int] = None,
        stop=None, stop_token_ids=None, temperature: Optional[float] = None, top_p: Optional[float] = None,
        top_k: Optional[int] = None, min_p: Optional[float] = None, frequency_penalty: Optional[float] = None,
        presence_penalty: Optional[float] = None, ignore_eos: Optional[bool] = None, return_logprob: Optional[bool] = None,
        logprob_start_len: Optional[int] = None, top_logprobs_num: Optional[int] = None, dtype: Any = None,
        choices: Optional[List[str]] = None, choices_method: Optional[ChoicesSamplingMethod] = None, regex: Optional[str] = None,
        json_schema: Optional[str] = None, ebnf: Optional[str] = None):
    if choices:
        return SglSelect(name, choices, 0.0 if temperature is None else temperature, choices_method or token_length_normalized)
    if regex is not None:
        re.compile(regex)   # fail early on an invalid pattern
    return SglGen(name, max_new_tokens=max_tokens, min_new_tokens=min_tokens, n=n, stop=stop, stop_token_ids=stop_token_ids,
                  temperature=temperature, top_p=top_p, top_k=top_k, min_p=min_p, frequency_penalty=frequency_penalty,
                  presence_penalty=presence_penalty, ignore_eos=ignore_eos, return_logprob=return_logprob,
                  logprob_start_len=logprob_start_len, top_logprobs_num=top_logprobs_num, dtype=dtype, regex=regex,
                  json_schema=json_schema, ebnf=ebnf)


def gen_int(name: Optional[str] = None, max_tokens: Optional[int] = None, **kw: Any):
    return gen(name, max_tokens=max_tokens, dtype="int", regex=r"[-+]?[0-9]+", **kw)


def gen_string(name: Optional[str] = None, max_tokens: Optional[int] = None, **kw: Any):
    return gen(name, max_tokens=max_tokens, dtype="str", **kw)


def select(name: Optional[str] = None, choices: Optional[List[str]] = None, temperature: float = 0.0,
           choices_method: ChoicesSamplingMethod = token_length_normalized):
    assert choices is not None
    return SglSelect(name, choices, temperature, choices_method)


def image(expr: Any):
    return SglImage(expr)


def video(path: str, num_frames: int):
    return SglVideo(path, num_frames)


def separate_reasoning(expr: Optional[SglExpr] = None, model_type: Optional[str] = None):
    return SglSeparateReasoning(model_type or "qwen3", expr)


def _role_common(name: str, expr: Optional[Any] = None):
    if expr is None:
        return SglExprList([SglRoleBegin(name), SglRoleEnd(name)])
    from .ir import SglConstantText

    return SglExprList([SglRoleBegin(name), expr if isinstance(expr, SglExpr) else SglConstantText(str(expr)), SglRoleEnd(name)])


def system(expr: Optional[Any] = None):
    return _role_common("system", expr)


def user(expr: Optional[Any] = None):
    return _role_common("user", expr)


def assistant(expr: Optional[Any] = None):
    return _role_common("assistant", expr)


def system_begin():
    return SglRoleBegin("system")


def system_end():
    return SglRoleEnd("system")


def user_begin():
    return SglRoleBegin("user")


def user_end():
    return SglRoleEnd("user")


def assistant_begin():
    return SglRoleBegin("assistant")


def assistant_end():
    return SglRoleEnd("assistant")
