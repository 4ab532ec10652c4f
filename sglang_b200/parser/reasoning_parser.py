"""Reasoning-content separation (``<think>…</think>`` families), non-streaming and streaming.

Parity: reference ``python/sglang/srt/parser/reasoning_parser.py:313-489`` (detectors for deepseek-r1, qwen3,
kimi, gpt-oss...; ``ReasoningParser.parse_non_stream`` / ``parse_stream_chunk``)."""
from __future__ import annotations

from typing import Optional, Tuple


class BaseReasoningDetector:
    def __init__(self, start: str, end: str, force_reasoning: bool = False, stream_reasoning: bool = True):
        self.start, self.end = start, end
        self.in_reasoning = force_reasoning
        self.stream_reasoning = stream_reasoning
        self._buf = ""
        self._stripped_start = False

    def detect_and_parse(self, text: str) -> Tuple[Optional[str], str]:
        in_r = self.in_reasoning or self.start in text
        if not in_r:
            return None, text
        t = text.replace(self.start, "", 1) if self.start in text else text
        if self.end not in t:
            return t.strip(), ""          # truncated while still thinking
        r, _, rest = t.partition(self.end)
        return r.strip(), rest.lstrip("\n")

    def parse_stream_chunk(self, chunk: str) -> Tuple[str, str]:
        """returns (reasoning_delta, content_delta)"""
        self._buf += chunk
        buf = self._buf
        # hold back the longest tail that is a proper prefix of a marker (it may complete in the next chunk)
        hold = 0
        for tok in (self.start, self.end):
            for k in range(min(len(tok) - 1, len(buf)), 0, -1):
                if buf.endswith(tok[:k]):
                    hold = max(hold, k)
                    break
        if hold:
            return self._emit(buf[:-hold], keep=buf[-hold:])
        return self._emit(buf, keep="")

    def _emit(self, text: str, keep: str) -> Tuple[str, str]:
        self._buf = keep
        r_out, c_out = "", ""
        if not self._stripped_start and self.start in text:
            text = text.replace(self.start, "", 1)
            self._stripped_start = True
            self.in_reasoning = True
        if self.in_reasoning:
            if self.end in text:
                r, _, rest = text.partition(self.end)
                self.in_reasoning = False
                r_out, c_out = r, rest.lstrip("\n")
            else:
                r_out = text
        else:
            c_out = text
        if not self.stream_reasoning and self.in_reasoning:
            self._buf = r_out + self._buf
            r_out = ""
        return r_out, c_out


_FAMILIES = {
    "deepseek-r1": dict(start="<think>", end="</think>", force_reasoning=True),
    "qwen3": dict(start="<think>", end="</think>", force_reasoning=False),
    "qwen3-thinking": dict(start="<think>", end="</think>", force_reasoning=True),
    "kimi": dict(start="◁think▷", end="◁/think▷", force_reasoning=False),
    "glm45": dict(start="<think>", end="</think>", force_reasoning=False),
    "step3": dict(start="<think>", end="</think>", force_reasoning=True),
    "gpt-oss": dict(start="<|channel|>analysis<|message|>", end="<|end|>", force_reasoning=False),
}


class ReasoningParser:
    def __init__(self, model_type: str = "qwen3", stream_reasoning: bool = True, force_reasoning: Optional[bool] = None):
        if model_type not in _FAMILIES:
            raise ValueError(f"unknown reasoning parser {model_type!r}; available {sorted(_FAMILIES)}")
        kw = dict(_FAMILIES[model_type])
        if force_reasoning is not None:
            kw["force_reasoning"] = force_reasoning
        self.detector = BaseReasoningDetector(stream_reasoning=stream_reasoning, **kw)

    def parse_non_stream(self, text: str) -> Tuple[Optional[str], str]:
        return self.detector.detect_and_parse(text)

    def parse_stream_chunk(self, chunk: str) -> Tuple[str, str]:
        return self.detector.parse_stream_chunk(chunk)
