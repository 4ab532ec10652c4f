"""sglang_b200 — a B200-native (sm_100a, NVLink 5 / NVSwitch) LLM serving engine with SGLang's API surface.

Public API mirrors ``python/sglang/__init__.py:25-108`` of the reference: ``Engine``, ``Runtime``,
``RuntimeEndpoint``, the ``sgl`` frontend primitives (``function``, ``gen``, ``select``, ``image``, role helpers)."""
__version__ = "0.1.0"

_LAZY = {
    "Engine": ("sglang_b200.entrypoints.engine", "Engine"),
    "ServerArgs": ("sglang_b200.server_args", "ServerArgs"),
    "SamplingParams": ("sglang_b200.sampling.sampling_params", "SamplingParams"),
    "Runtime": ("sglang_b200.lang.backend.runtime_endpoint", "Runtime"),
    "RuntimeEndpoint": ("sglang_b200.lang.backend.runtime_endpoint", "RuntimeEndpoint"),
    "launch_server": ("sglang_b200.entrypoints.http_server", "launch_server"),
    "OpenAI": ("sglang_b200.lang.backend.openai", "OpenAI"),
    "Anthropic": ("sglang_b200.lang.backend.openai", "Anthropic"),
    "LiteLLM": ("sglang_b200.lang.backend.openai", "LiteLLM"),
    "VertexAI": ("sglang_b200.lang.backend.openai", "VertexAI"),
}
_LANG = ["function", "gen", "gen_int", "gen_string", "select", "image", "video", "system", "user", "assistant",
         "system_begin", "system_end", "user_begin", "user_end", "assistant_begin", "assistant_end", "set_default_backend",
         "get_default_backend", "flush_cache", "get_server_info", "greedy_token_selection", "token_length_normalized",
         "unconditional_likelihood_normalized", "separate_reasoning"]


def __getattr__(name):
    import importlib

    if name in _LAZY:
        mod, attr = _LAZY[name]
        return getattr(importlib.import_module(mod), attr)
    if name in _LANG:
        return getattr(importlib.import_module("sglang_b200.lang.api"), name)
    raise AttributeError(name)
