"""Tool-call extraction from model text (non-streaming and streaming).

Parity: reference ``python/sglang/srt/function_call/`` (``FunctionCallParser:56``, ``parse_stream_chunk:156``, one
``*_detector.py`` per model family).  Families here: ``llama3`` (``<|python_tag|>`` / bare JSON), ``qwen25`` / ``hermes``
(``<tool_call>{json}</tool_call>``), ``mistral`` (``[TOOL_CALLS] [...]``), ``deepseekv3``
(``<｜tool▁call▁begin｜>function<｜tool▁sep｜>name\\n```json ... ```<｜tool▁call▁end｜>``), ``pythonic`` (``[f(a=1), g()]``),
``json`` (any bare ``{"name":..., "arguments"|"parameters": ...}``).  Each detector also produces the *structural tag*
used to constrain generation when ``tool_choice`` is ``required`` / a named function.
"""
from __future__ import annotations

import ast
import json
import re
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Tuple


@dataclass
class ToolCallItem:
    tool_index: int
    name: Optional[str]
    parameters: str   # JSON string (possibly a partial delta when streaming)


@dataclass
class ParseResult:
    normal_text: str = ""
    calls: List[ToolCallItem] = field(default_factory=list)


def _tool_names(tools: List[Any]) -> List[str]:
    out = []
    for t in tools or []:
        f = t.function if hasattr(t, "function") else t.get("function", t)
        out.append(f.name if hasattr(f, "name") else f["name"])
    return out


def _norm_call(obj: Dict[str, Any], idx: int, names: List[str]) -> Optional[ToolCallItem]:
    if not isinstance(obj, dict) or "name" not in obj:
        return None
    if names and obj["name"] not in names:
        return None
    args = obj.get("arguments", obj.get("parameters", {}))
    if isinstance(args, str):
        try:
            args = json.loads(args)
        except Exception:
            pass
    return ToolCallItem(idx, obj["name"], json.dumps(args, ensure_ascii=False))


def _find_json_objects(text: str) -> List[Tuple[int, int, Any]]:
    """all top-level JSON objects/arrays embedded in text: (start, end, value)"""
    dec = json.JSONDecoder()
    out, i = [], 0
    while i < len(text):
        if text[i] in "{[":
            try:
                val, end = dec.raw_decode(text, i)
                out.append((i, end, val))
                i = end
                continue
            except Exception:
                pass
        i += 1
    return out


class BaseDetector:
    bot_token = ""
    eot_token = ""

    def has_tool_call(self, text: str) -> bool:
        return bool(self.bot_token) and self.bot_token in text

    def detect_and_parse(self, text: str, tools: List[Any]) -> ParseResult:
        raise NotImplementedError

    def structure_info(self, name: str) -> Tuple[str, str]:
        """(begin, end) literal text wrapping the JSON arguments of a call to ``name``"""
        raise NotImplementedError


class JsonDetector(BaseDetector):
    def has_tool_call(self, text: str) -> bool:
        return '"name"' in text

    def detect_and_parse(self, text, tools):
        names = _tool_names(tools)
        res = ParseResult()
        last = 0
        for s, e, val in _find_json_objects(text):
            vals = val if isinstance(val, list) else [val]
            calls = [c for c in (_norm_call(v, len(res.calls) + i, names) for i, v in enumerate(vals)) if c]
            if calls:
                res.normal_text += text[last:s]
                res.calls.extend(calls)
                last = e
        res.normal_text += text[last:]
        for i, c in enumerate(res.calls):
            c.tool_index = i
        return res

    def structure_info(self, name):
        return '{"name": "' + name + '", "arguments": ', "}"


class Llama3Detector(JsonDetector):
    bot_token = "<|python_tag|>"

    def has_tool_call(self, text):
        return self.bot_token in text or text.lstrip().startswith("{")

    def detect_and_parse(self, text, tools):
        if self.bot_token in text:
            head, _, tail = text.partition(self.bot_token)
            r = super().detect_and_parse(tail.replace(";", " "), tools)
            r.normal_text = head + ("" if r.calls else tail)
            return r
        return super().detect_and_parse(text, tools)

    def structure_info(self, name):
        return '<|python_tag|>{"name": "' + name + '", "parameters": ', "}"


class Qwen25Detector(BaseDetector):
    bot_token, eot_token = "<tool_call>", "</tool_call>"

    def detect_and_parse(self, text, tools):
        names = _tool_names(tools)
        res = ParseResult()
        pos = 0
        for m in re.finditer(r"<tool_call>\s*(.*?)\s*</tool_call>", text, re.S):
            res.normal_text += text[pos:m.start()]
            pos = m.end()
            try:
                c = _norm_call(json.loads(m.group(1)), len(res.calls), names)
                if c:
                    res.calls.append(c)
            except Exception:
                res.normal_text += m.group(0)
        res.normal_text += text[pos:]
        return res

    def structure_info(self, name):
        return '<tool_call>\n{"name": "' + name + '", "arguments": ', "}\n</tool_call>"


class MistralDetector(BaseDetector):
    bot_token = "[TOOL_CALLS]"

    def detect_and_parse(self, text, tools):
        names = _tool_names(tools)
        head, _, tail = text.partition(self.bot_token)
        res = ParseResult(normal_text=head)
        objs = _find_json_objects(tail)
        for _, _, val in objs[:1]:
            for v in (val if isinstance(val, list) else [val]):
                c = _norm_call(v, len(res.calls), names)
                if c:
                    res.calls.append(c)
        if not res.calls:
            res.normal_text = text
        return res

    def structure_info(self, name):
        return '[TOOL_CALLS] [{"name": "' + name + '", "arguments": ', "}]"


class DeepSeekV3Detector(BaseDetector):
    bot_token = "<｜tool▁calls▁begin｜>"

    def detect_and_parse(self, text, tools):
        names = _tool_names(tools)
        head, _, tail = text.partition(self.bot_token)
        res = ParseResult(normal_text=head)
        pat = r"<｜tool▁call▁begin｜>(?:function)?<｜tool▁sep｜>(.*?)\n```json\n(.*?)\n```<｜tool▁call▁end｜>"
        for m in re.finditer(pat, tail, re.S):
            name = m.group(1).strip()
            if names and name not in names:
                continue
            try:
                args = json.loads(m.group(2))
            except Exception:
                continue
            res.calls.append(ToolCallItem(len(res.calls), name, json.dumps(args, ensure_ascii=False)))
        if not res.calls:
            res.normal_text = text
        return res

    def structure_info(self, name):
        return f"<｜tool▁calls▁begin｜><｜tool▁call▁begin｜>function<｜tool▁sep｜>{name}\n```json\n", "\n```<｜tool▁call▁end｜><｜tool▁calls▁end｜>"


class PythonicDetector(BaseDetector):
    def has_tool_call(self, text):
        return bool(re.search(r"\[\s*[A-Za-z_][\w\.]*\(", text))

    def detect_and_parse(self, text, tools):
        names = _tool_names(tools)
        m = re.search(r"\[\s*[A-Za-z_][\w\.]*\(.*\)\s*\]", text, re.S)
        if not m:
            return ParseResult(normal_text=text)
        res = ParseResult(normal_text=text[:m.start()] + text[m.end():])
        try:
            tree = ast.parse(m.group(0).strip(), mode="eval")
            for el in tree.body.elts:
                if not isinstance(el, ast.Call):
                    continue
                name = ast.unparse(el.func)
                if names and name not in names:
                    continue
                args = {kw.arg: ast.literal_eval(kw.value) for kw in el.keywords}
                res.calls.append(ToolCallItem(len(res.calls), name, json.dumps(args, ensure_ascii=False)))
        except Exception:
            return ParseResult(normal_text=text)
        return res

    def structure_info(self, name):
        return f"[{name}(", ")]"


DETECTORS = {"llama3": Llama3Detector, "qwen25": Qwen25Detector, "qwen": Qwen25Detector, "hermes": Qwen25Detector,
             "mistral": MistralDetector, "deepseekv3": DeepSeekV3Detector, "pythonic": PythonicDetector, "json": JsonDetector}


class FunctionCallParser:
    def __init__(self, tools: List[Any], tool_call_parser: str = "json"):
        if tool_call_parser not in DETECTORS:
            raise ValueError(f"unknown tool_call_parser {tool_call_parser!r}; available {sorted(DETECTORS)}")
        self.detector: BaseDetector = DETECTORS[tool_call_parser]()
        self.tools = tools
        self._buf = ""
        self._sent_text = 0
        self._sent_calls = 0

    def has_tool_call(self, text: str) -> bool:
        return self.detector.has_tool_call(text)

    def parse_non_stream(self, text: str) -> Tuple[str, List[ToolCallItem]]:
        r = self.detector.detect_and_parse(text, self.tools)
        return r.normal_text, r.calls

    def parse_stream_chunk(self, chunk: str) -> Tuple[str, List[ToolCallItem]]:
        """Incremental: returns newly-safe normal text and newly-completed calls.  Text that could still be the start
        of a tool-call marker is held back until it is disambiguated."""
        self._buf += chunk
        bot = self.detector.bot_token
        text = self._buf
        if not self.detector.has_tool_call(text):
            hold = 0
            if bot:
                for k in range(min(len(bot) - 1, len(text)), 0, -1):
                    if text.endswith(bot[:k]):
                        hold = k
                        break
            safe = text[: len(text) - hold]
            delta = safe[self._sent_text:]
            self._sent_text = len(safe)
            return delta, []
        r = self.detector.detect_and_parse(text, self.tools)
        new_calls = r.calls[self._sent_calls:]
        self._sent_calls = max(self._sent_calls, len(r.calls))     # a call, once emitted, stays emitted (partial closers may un-match)
        # only text before the first marker is safe to emit
        cut = text.find(bot) if bot and bot in text else len(r.normal_text)
        safe = text[:cut] if bot else r.normal_text
        delta = safe[self._sent_text:] if len(safe) > self._sent_text else ""
        self._sent_text = max(self._sent_text, len(safe))
        return delta, new_calls

    def get_structural_constraint(self, tool_choice: Any) -> Optional[Dict[str, Any]]:
        """regex-style constraint forcing a call (``required``) or a specific function (named choice)"""
        names = _tool_names(self.tools)
        if hasattr(tool_choice, "function"):
            names = [tool_choice.function.name]
        elif tool_choice not in ("required",):
            return None
        alts = []
        for t in self.tools:
            f = t.function if hasattr(t, "function") else t["function"]
            name = f.name if hasattr(f, "name") else f["name"]
            if name not in names:
                continue
            begin, end = self.detector.structure_info(name)
            params = (f.parameters if hasattr(f, "parameters") else f.get("parameters")) or {"type": "object"}
            alts.append({"begin": begin, "schema": params, "end": end})
        return {"type": "structural_tag", "structures": alts, "triggers": [a["begin"][:8] for a in alts]}
