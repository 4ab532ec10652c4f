"""Further tool-call formats, built from three small grammars instead of one hand-written class per model:

* ``WrappedJsonDetector``   ``<open> {json | [json,...]} <close>`` with configurable name/argument keys
* ``NameThenJsonDetector``  ``<call-open> NAME <sep> {json} <call-close>`` (DeepSeek-V3.1, Kimi-K2, GPT-OSS/harmony, Step-3 json form)
* ``XmlParamDetector``      ``<open> <function-open NAME> (<param-open KEY> value <param-close>)* <function-close> <close>`` (Qwen3-Coder, GLM-4.x-MoE,
  MiniMax-M2, MiMo, DeepSeek-V3.2 DSML, Step-3, dots) — parameter values are typed with the tool's JSON schema when one is known

Parity: reference ``python/sglang/srt/function_call/{hermes,internlm,cohere_command4,llama32,lfm2,deepseekv31,kimik2,gpt_oss,
qwen3_coder,glm4_moe,glm47_moe,minimax_m2,mimo,deepseekv32,step3,dots,hunyuan,apertus2509}_detector.py`` — the formats (marker
tokens and nesting) are taken from those files' documented examples; the parsing code is ours.
"""
from __future__ import annotations

import json
import re
from typing import Any, Dict, List, Optional, Tuple

from .function_call_parser import (DETECTORS, BaseDetector, ParseResult, PythonicDetector, ToolCallItem, _find_json_objects, _norm_call,
                                   _tool_names)


def _schema_of(tools: List[Any], name: str) -> Dict[str, Any]:
    for t in tools or []:
        f = t.function if hasattr(t, "function") else t.get("function", t)
        n = f.name if hasattr(f, "name") else f["name"]
        if n == name:
            p = f.parameters if hasattr(f, "parameters") else f.get("parameters")
            return (p or {}).get("properties", {}) if isinstance(p, dict) else {}
    return {}


def _typed(value: str, spec: Optional[Dict[str, Any]]) -> Any:
    """XML-ish formats carry untyped text: use the schema type when present, otherwise try JSON, otherwise keep the string"""
    v = value.strip("\n") if value.startswith("\n") or value.endswith("\n") else value
    typ = (spec or {}).get("type")
    if typ == "string":
        return v
    if typ in ("integer", "number", "boolean", "object", "array", "null"):
        try:
            return json.loads(v if typ != "boolean" else v.strip().lower())
        except Exception:
            return v
    try:
        return json.loads(v)
    except Exception:
        return v


class WrappedJsonDetector(BaseDetector):
    name_key, args_keys = "name", ("arguments", "parameters")

    def __init__(self, bot: str, eot: str, name_key: str = "name", args_keys: Tuple[str, ...] = ("arguments", "parameters")):
        self.bot_token, self.eot_token, self.name_key, self.args_keys = bot, eot, name_key, args_keys

    def detect_and_parse(self, text, tools):
        names = _tool_names(tools)
        res = ParseResult()
        pat = re.escape(self.bot_token) + r"\s*(.*?)\s*" + (re.escape(self.eot_token) if self.eot_token else r"$")
        pos = 0
        for m in re.finditer(pat, text, re.S):
            res.normal_text += text[pos:m.start()]
            pos = m.end()
            got = False
            for _, _, val in _find_json_objects(m.group(1)):
                for v in (val if isinstance(val, list) else [val]):
                    if isinstance(v, dict) and self.name_key in v:
                        args = next((v[k] for k in self.args_keys if k in v), {})
                        c = _norm_call({"name": v[self.name_key], "arguments": args}, len(res.calls), names)
                        if c:
                            res.calls.append(c)
                            got = True
            if not got:
                res.normal_text += m.group(0)
        res.normal_text += text[pos:]
        return res

    def structure_info(self, name):
        return f'{self.bot_token}{{"{self.name_key}": "{name}", "{self.args_keys[0]}": ', "}" + self.eot_token


class NameThenJsonDetector(BaseDetector):
    """``call_re`` must expose groups ``name`` and ``args`` (a JSON object)"""

    def __init__(self, bot: str, call_re: str, begin_fmt: str, end: str, strip_prefix: str = ""):
        self.bot_token, self.call_re, self.begin_fmt, self.end, self.strip_prefix = bot, re.compile(call_re, re.S), begin_fmt, end, strip_prefix

    def detect_and_parse(self, text, tools):
        names = _tool_names(tools)
        idx = text.find(self.bot_token)
        if idx < 0:
            return ParseResult(normal_text=text)
        res = ParseResult(normal_text=text[:idx])
        for m in self.call_re.finditer(text, idx):
            name = m.group("name").strip()
            if self.strip_prefix and name.startswith(self.strip_prefix):
                name = name[len(self.strip_prefix):]
            name = name.split(":")[0] if ":" in name and name.rsplit(":", 1)[1].isdigit() else name      # kimi "functions.f:0"
            if names and name not in names:
                continue
            try:
                args = json.loads(m.group("args"))
            except Exception:
                continue
            res.calls.append(ToolCallItem(len(res.calls), name, json.dumps(args, ensure_ascii=False)))
        if not res.calls:
            res.normal_text = text
        return res

    def structure_info(self, name):
        return self.begin_fmt.format(name=name), self.end


class XmlParamDetector(BaseDetector):
    """``func_re``: groups ``name`` and ``body``; ``param_re``: groups ``key`` and ``value`` (matched inside ``body``)"""

    def __init__(self, bot: str, eot: str, func_re: str, param_re: str, begin_fmt: str, end: str):
        self.bot_token, self.eot_token = bot, eot
        self.func_re, self.param_re = re.compile(func_re, re.S), re.compile(param_re, re.S)
        self.begin_fmt, self.end = begin_fmt, end

    def detect_and_parse(self, text, tools):
        names = _tool_names(tools)
        idx = text.find(self.bot_token)
        if idx < 0:
            return ParseResult(normal_text=text)
        res = ParseResult(normal_text=text[:idx])
        for m in self.func_re.finditer(text, idx):
            name = m.group("name").strip()
            if names and name not in names:
                continue
            props = _schema_of(tools, name)
            args = {pm.group("key").strip(): _typed(pm.group("value"), props.get(pm.group("key").strip())) for pm in self.param_re.finditer(m.group("body"))}
            res.calls.append(ToolCallItem(len(res.calls), name, json.dumps(args, ensure_ascii=False)))
        if not res.calls:
            res.normal_text = text
        return res

    def structure_info(self, name):
        return self.begin_fmt.format(name=name), self.end


class Lfm2Detector(PythonicDetector):
    bot_token, eot_token = "<|tool_call_start|>", "<|tool_call_end|>"

    def has_tool_call(self, text):
        return self.bot_token in text

    def detect_and_parse(self, text, tools):
        m = re.search(re.escape(self.bot_token) + r"(.*?)" + re.escape(self.eot_token), text, re.S)
        if not m:
            return ParseResult(normal_text=text)
        r = super().detect_and_parse(m.group(1), tools)
        r.normal_text = text[:m.start()] + text[m.end():]
        return r

    def structure_info(self, name):
        return f"{self.bot_token}[{name}(", f")]{self.eot_token}"


def _register() -> None:
    D = DETECTORS
    D["hermes"] = lambda: WrappedJsonDetector("<tool_call>", "</tool_call>")
    D["mimo-json"] = D["hermes"]
    D["internlm"] = lambda: WrappedJsonDetector("<|action_start|><|plugin|>", "<|action_end|>", "name", ("parameters", "arguments"))
    D["cohere_command4"] = lambda: WrappedJsonDetector("<|START_ACTION|>", "<|END_ACTION|>", "tool_name", ("parameters", "arguments"))
    D["llama32"] = D["llama3"]
    D["apertus2509"] = lambda: WrappedJsonDetector("<|tools_prefix|>", "<|tools_suffix|>")
    D["lfm2"] = Lfm2Detector
    D["deepseekv31"] = lambda: NameThenJsonDetector(
        "<｜tool▁calls▁begin｜>", r"<｜tool▁call▁begin｜>(?P<name>.*?)<｜tool▁sep｜>(?P<args>\{.*?\})\s*<｜tool▁call▁end｜>",
        "<｜tool▁calls▁begin｜><｜tool▁call▁begin｜>{name}<｜tool▁sep｜>", "<｜tool▁call▁end｜><｜tool▁calls▁end｜>")
    D["kimi_k2"] = lambda: NameThenJsonDetector(
        "<|tool_calls_section_begin|>", r"<\|tool_call_begin\|>\s*(?P<name>[\w\.\-]+(?::\d+)?)\s*<\|tool_call_argument_begin\|>\s*(?P<args>\{.*?\})\s*<\|tool_call_end\|>",
        "<|tool_calls_section_begin|><|tool_call_begin|>functions.{name}:0<|tool_call_argument_begin|>", "<|tool_call_end|><|tool_calls_section_end|>",
        strip_prefix="functions.")
    D["gpt-oss"] = lambda: NameThenJsonDetector(
        "<|channel|>commentary", r"<\|channel\|>commentary to=(?P<name>[\w\.\-]+)\s*(?:<\|constrain\|>\s*json)?\s*<\|message\|>(?P<args>\{.*?\})\s*(?:<\|call\|>|$)",
        "<|start|>assistant<|channel|>commentary to=functions.{name} <|constrain|>json<|message|>", "<|call|>", strip_prefix="functions.")
    D["qwen3_coder"] = lambda: XmlParamDetector(
        "<tool_call>", "</tool_call>", r"<function=(?P<name>[^>\n]+)>(?P<body>.*?)</function>", r"<parameter=(?P<key>[^>\n]+)>\n?(?P<value>.*?)\n?</parameter>",
        "<tool_call>\n<function={name}>\n", "</function>\n</tool_call>")
    D["mimo"] = D["qwen3_coder"]
    D["glm45"] = lambda: XmlParamDetector(
        "<tool_call>", "</tool_call>", r"<tool_call>\s*(?P<name>[^\n<]+)\s*(?P<body>.*?)</tool_call>",
        r"<arg_key>(?P<key>.*?)</arg_key>\s*<arg_value>(?P<value>.*?)</arg_value>", "<tool_call>{name}\n", "</tool_call>")
    D["glm47"] = D["glm45"]
    D["minimax-m2"] = lambda: XmlParamDetector(
        "<minimax:tool_call>", "</minimax:tool_call>", r'<invoke name="(?P<name>[^"]+)">(?P<body>.*?)</invoke>',
        r'<parameter name="(?P<key>[^"]+)">(?P<value>.*?)</parameter>', '<minimax:tool_call>\n<invoke name="{name}">\n', "</invoke>\n</minimax:tool_call>")
    D["deepseekv32"] = lambda: XmlParamDetector(
        "<｜DSML｜function_calls>", "</｜DSML｜function_calls>", r'<｜DSML｜invoke name="(?P<name>[^"]+)">(?P<body>.*?)</｜DSML｜invoke>',
        r'<｜DSML｜parameter\s+name="(?P<key>[^"]+)"(?:\s+string="[^"]*")?\s*>(?P<value>.*?)</｜DSML｜parameter>',
        '<｜DSML｜function_calls>\n<｜DSML｜invoke name="{name}">\n', "</｜DSML｜invoke>\n</｜DSML｜function_calls>")
    D["step3"] = lambda: XmlParamDetector(
        "<｜tool_calls_begin｜>", "<｜tool_calls_end｜>", r'<steptml:invoke name="(?P<name>[^"]+)">(?P<body>.*?)</steptml:invoke>',
        r'<steptml:parameter name="(?P<key>[^"]+)">(?P<value>.*?)</steptml:parameter>',
        '<｜tool_calls_begin｜><｜tool_call_begin｜>function<｜tool_sep｜><steptml:invoke name="{name}">\n',
        "</steptml:invoke><｜tool_call_end｜><｜tool_calls_end｜>")
    D["dots"] = lambda: XmlParamDetector(
        "<dots_function_call>", "</dots_function_call>", r'<invoke name="(?P<name>[^"]+)">(?P<body>.*?)</invoke>',
        r'<parameter name="(?P<key>[^"]+)">(?P<value>.*?)</parameter>', '<dots_function_call>\n<invoke name="{name}">\n', "</invoke>\n</dots_function_call>")
    D["minicpm5"] = lambda: XmlParamDetector(
        "<function", "</function>", r"<function\s+name=['\"](?P<name>[^'\"]+)['\"][^>]*>(?P<body>.*?)</function>",
        r"<param\s+name=['\"](?P<key>[^'\"]+)['\"][^>]*>(?P<value>.*?)</param>", '<function name="{name}">\n', "</function>")


_register()
