"""Frontend-language backends for hosted / third-party APIs: any OpenAI-compatible endpoint and the Anthropic Messages API.

Parity: reference ``python/sglang/lang/backend/{openai,anthropic,litellm,vertexai}.py`` — ``sgl.function`` programs run unchanged against a
model that is NOT served by this engine.  Differences from ``RuntimeEndpoint``: the conversation is sent as role-tagged messages (the remote
side applies its own chat template), there is no prefix cache to warm and no token-level ``select`` on chat models; completion-style models
(``is_chat_model=False``) get the flat prompt.  Plain HTTP through ``requests`` (no SDK dependency; ``session`` can be injected for tests).
``LiteLLM`` / ``VertexAI`` names resolve to the OpenAI-compatible class pointed at the respective gateway.
"""
from __future__ import annotations

import json
from typing import Any, Dict, Iterator, List, Optional, Tuple

from ..chat_template import get_chat_template
from ..choices import ChoicesDecision, ChoicesSamplingMethod
from ..ir import SglSamplingParams
from .runtime_endpoint import BaseBackend


def _conversation(s: Any) -> Tuple[List[Dict[str, Any]], str]:
    """(closed turns, text already written into the currently open turn)"""
    tr = s._tr
    partial = tr.text()[tr.role_mark:].lstrip() if tr.open_role is not None else ""
    return [dict(m) for m in tr.messages], partial


class OpenAI(BaseBackend):
    """``sgl.set_default_backend(OpenAI("gpt-4o-mini"))`` or any OpenAI-compatible server: ``OpenAI("m", base_url="http://host:30000/v1")``"""

    def __init__(self, model_name: str, base_url: str = "https://api.openai.com/v1", api_key: Optional[str] = None, is_chat_model: Optional[bool] = None,
                 session: Any = None, timeout: float = 600.0, **default_params: Any):
        super().__init__()
        import os

        self.model_name = model_name
        self.base_url = base_url.rstrip("/")
        self.timeout = timeout
        self.default_params = default_params
        self.is_chat_model = is_chat_model if is_chat_model is not None else not any(t in model_name for t in ("instruct-", "davinci", "babbage"))
        self.chat_template = get_chat_template("default")
        if session is None:
            import requests

            session = requests.Session()
        self.s = session
        key = api_key or os.environ.get("OPENAI_API_KEY")
        self.headers = {"Authorization": f"Bearer {key}"} if key else {}
        self.token_usage = {"prompt_tokens": 0, "completion_tokens": 0}

    def get_model_name(self) -> str:
        return self.model_name

    # ---- request building
    def _params(self, sp: SglSamplingParams) -> Dict[str, Any]:
        p: Dict[str, Any] = dict(self.default_params)
        p.update(model=self.model_name, max_tokens=sp.max_new_tokens, temperature=sp.temperature, top_p=sp.top_p)
        if sp.stop:
            p["stop"] = sp.stop if isinstance(sp.stop, list) else [sp.stop]
        if sp.frequency_penalty:
            p["frequency_penalty"] = sp.frequency_penalty
        if sp.presence_penalty:
            p["presence_penalty"] = sp.presence_penalty
        if getattr(sp, "json_schema", None):
            p["response_format"] = {"type": "json_schema", "json_schema": {"name": "out", "schema": json.loads(sp.json_schema)}}
        elif getattr(sp, "regex", None):
            p["regex"] = sp.regex            # OpenAI-compatible servers that support it (this engine, vLLM) honour it; others ignore it
        return p

    def _body(self, s: Any, sp: SglSamplingParams) -> Tuple[str, Dict[str, Any]]:
        body = self._params(sp)
        if self.is_chat_model:
            msgs, partial = _conversation(s)
            if s._tr.open_role not in (None, "assistant"):
                raise ValueError("a chat model can only generate inside an assistant turn")
            if partial:
                # continuing a half-written assistant turn: supported by this engine / vLLM (continue_final_message), not by api.openai.com
                msgs.append({"role": "assistant", "content": partial})
                body["continue_final_message"] = True
                body["add_generation_prompt"] = False
            if not msgs:
                msgs = [{"role": "user", "content": s._tr.text()}]
            body["messages"] = msgs
            return "/chat/completions", body
        body["prompt"] = s.prompt
        return "/completions", body

    def _post(self, path: str, body: Dict[str, Any], stream: bool = False):
        r = self.s.post(self.base_url + path, json=body, headers=self.headers, timeout=self.timeout, **({"stream": True} if stream else {}))
        if r.status_code != 200:
            raise RuntimeError(f"{path} failed ({r.status_code}): {getattr(r, 'text', '')[:500]}")
        return r

    @staticmethod
    def _meta(obj: Dict[str, Any], finish: Any) -> Dict[str, Any]:
        u = obj.get("usage") or {}
        return {"prompt_tokens": u.get("prompt_tokens", 0), "completion_tokens": u.get("completion_tokens", 0),
                "finish_reason": {"type": finish or "stop"}, "id": obj.get("id")}

    # ---- BaseBackend surface
    def generate(self, s: Any, sampling_params: SglSamplingParams) -> Tuple[str, Dict[str, Any]]:
        path, body = self._body(s, sampling_params)
        obj = self._post(path, body).json()
        ch = obj["choices"][0]
        text = ch["message"]["content"] if "message" in ch else ch["text"]
        meta = self._meta(obj, ch.get("finish_reason"))
        self.token_usage["prompt_tokens"] += meta["prompt_tokens"]
        self.token_usage["completion_tokens"] += meta["completion_tokens"]
        return text or "", meta

    def generate_stream(self, s: Any, sampling_params: SglSamplingParams) -> Iterator[Tuple[str, Dict[str, Any]]]:
        path, body = self._body(s, sampling_params)
        body["stream"] = True
        r = self._post(path, body, stream=True)
        finish = None
        for line in r.iter_lines():
            if isinstance(line, bytes):
                line = line.decode("utf-8", "replace")
            if not line or not line.startswith("data:") or "[DONE]" in line:
                continue
            obj = json.loads(line[5:])
            if not obj.get("choices"):
                continue
            ch = obj["choices"][0]
            piece = (ch.get("delta") or {}).get("content") if "delta" in ch else ch.get("text")
            finish = ch.get("finish_reason") or finish
            if piece:
                yield piece, self._meta(obj, finish)
        yield "", self._meta({}, finish)

    def select(self, s: Any, choices: List[str], temperature: float, choices_method: Optional[ChoicesSamplingMethod]) -> ChoicesDecision:
        """chat APIs expose no prompt logprobs: ask for the answer constrained to the choices (guided_choice where the server supports it) and
        map the reply onto the closest choice"""
        sp = SglSamplingParams(max_new_tokens=max(8, max(len(c) for c in choices) // 2 + 4), temperature=0.0)
        path, body = self._body(s, sp)
        body["guided_choice"] = list(choices)
        obj = self._post(path, body).json()
        ch = obj["choices"][0]
        text = (ch["message"]["content"] if "message" in ch else ch["text"]) or ""
        t = text.strip()
        best = next((c for c in choices if t == c), None) or next((c for c in choices if t.startswith(c) or c.startswith(t) and t), None) \
            or max(choices, key=lambda c: sum(a == b for a, b in zip(c, t)))
        return ChoicesDecision(decision=best, meta_info={"raw": text, "method": "constrained-generation"})


class Anthropic(BaseBackend):
    """Anthropic Messages API (``/v1/messages``): system prompt as a top-level field, assistant prefill supported natively"""

    def __init__(self, model_name: str, base_url: str = "https://api.anthropic.com", api_key: Optional[str] = None, session: Any = None,
                 timeout: float = 600.0, version: str = "2023-06-01"):
        super().__init__()
        import os

        self.model_name = model_name
        self.base_url = base_url.rstrip("/")
        self.timeout = timeout
        if session is None:
            import requests

            session = requests.Session()
        self.s = session
        key = api_key or os.environ.get("ANTHROPIC_API_KEY")
        self.headers = {"anthropic-version": version, **({"x-api-key": key} if key else {})}

    def get_model_name(self) -> str:
        return self.model_name

    def _body(self, s: Any, sp: SglSamplingParams) -> Dict[str, Any]:
        msgs, partial = _conversation(s)
        system = "\n".join(m["content"] for m in msgs if m["role"] == "system" and isinstance(m["content"], str))
        turns = [m for m in msgs if m["role"] != "system"]
        if partial:
            turns.append({"role": "assistant", "content": partial})
        if not turns:
            turns = [{"role": "user", "content": s._tr.text()}]
        body: Dict[str, Any] = {"model": self.model_name, "messages": turns, "max_tokens": sp.max_new_tokens, "temperature": sp.temperature,
                                "top_p": sp.top_p}
        if system:
            body["system"] = system
        if sp.stop:
            body["stop_sequences"] = sp.stop if isinstance(sp.stop, list) else [sp.stop]
        return body

    def generate(self, s: Any, sampling_params: SglSamplingParams) -> Tuple[str, Dict[str, Any]]:
        r = self.s.post(self.base_url + "/v1/messages", json=self._body(s, sampling_params), headers=self.headers, timeout=self.timeout)
        if r.status_code != 200:
            raise RuntimeError(f"/v1/messages failed ({r.status_code}): {getattr(r, 'text', '')[:500]}")
        obj = r.json()
        text = "".join(b.get("text", "") for b in obj.get("content", []) if b.get("type") == "text")
        u = obj.get("usage") or {}
        return text, {"prompt_tokens": u.get("input_tokens", 0), "completion_tokens": u.get("output_tokens", 0),
                      "finish_reason": {"type": obj.get("stop_reason") or "stop"}, "id": obj.get("id")}

    def generate_stream(self, s: Any, sampling_params: SglSamplingParams) -> Iterator[Tuple[str, Dict[str, Any]]]:
        body = self._body(s, sampling_params)
        body["stream"] = True
        r = self.s.post(self.base_url + "/v1/messages", json=body, headers=self.headers, timeout=self.timeout, stream=True)
        if r.status_code != 200:
            raise RuntimeError(f"/v1/messages failed ({r.status_code})")
        for line in r.iter_lines():
            if isinstance(line, bytes):
                line = line.decode("utf-8", "replace")
            if line and line.startswith("data:"):
                ev = json.loads(line[5:])
                if ev.get("type") == "content_block_delta" and ev.get("delta", {}).get("type") == "text_delta":
                    yield ev["delta"]["text"], {}

    def select(self, s: Any, choices: List[str], temperature: float, choices_method: Optional[ChoicesSamplingMethod]) -> ChoicesDecision:
        text, _ = self.generate(s, SglSamplingParams(max_new_tokens=max(8, max(len(c) for c in choices) // 2 + 4), temperature=0.0))
        t = text.strip()
        best = next((c for c in choices if t.startswith(c)), None) or max(choices, key=lambda c: sum(a == b for a, b in zip(c, t)))
        return ChoicesDecision(decision=best, meta_info={"raw": text, "method": "generation"})


class LiteLLM(OpenAI):
    """LiteLLM proxy: an OpenAI-compatible gateway (``litellm --port 4000``)"""

    def __init__(self, model_name: str, base_url: str = "http://127.0.0.1:4000/v1", **kw: Any):
        super().__init__(model_name, base_url=base_url, **kw)


class VertexAI(OpenAI):
    """Vertex AI's OpenAI-compatible surface (``.../endpoints/openapi``); pass the access token as ``api_key``"""

    def __init__(self, model_name: str, base_url: str, **kw: Any):
        super().__init__(model_name, base_url=base_url, **kw)
