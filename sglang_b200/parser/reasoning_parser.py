"""Reasoning-content separation, non-streaming and streaming.

Parity: reference ``python/sglang/srt/parser/reasoning_parser.py`` (per-family detectors behind ``ReasoningParser.parse_non_stream`` /
``parse_stream_chunk``).  Here ONE table-driven detector covers the tag-pair families (start / end marker, whether the model starts
inside the reasoning block, an optional answer wrapper to strip, an optional marker that ends the reasoning implicitly — a tool-call
opener) and a second one the channel-structured Harmony format of GPT-OSS (``analysis`` -> reasoning, ``final`` -> content,
``commentary`` passed through for the tool-call parser).  Streaming never emits a partial marker: the longest suffix that could still
become a marker is held back until the next chunk."""
from __future__ import annotations

from typing import Optional, Tuple


class BaseReasoningDetector:
    def __init__(self, start: str, end: str, force_reasoning: bool = False, stream_reasoning: bool = True, answer_wrap: Optional[Tuple[str, str]] = None,
                 implicit_end: Optional[str] = None):
        self.start, self.end = start, end
        self.in_reasoning = force_reasoning
        self.stream_reasoning = stream_reasoning
        self.answer_wrap = answer_wrap            # e.g. ("<answer>", "</answer>"): stripped from the content side
        self.implicit_end = implicit_end          # e.g. a tool-call opener: reasoning ends there even without the end marker
        self._buf = ""
        self._stripped_start = False

    def _unwrap(self, content: str) -> str:
        if self.answer_wrap:
            a, b = self.answer_wrap
            content = content.replace(a, "", 1).replace(b, "", 1)
        return content

    def detect_and_parse(self, text: str) -> Tuple[Optional[str], str]:
        in_r = self.in_reasoning or self.start in text
        if not in_r:
            return None, self._unwrap(text)
        t = text.replace(self.start, "", 1) if self.start in text else text
        cut = t.find(self.end)
        imp = t.find(self.implicit_end) if self.implicit_end else -1
        if imp >= 0 and (cut < 0 or imp < cut):   # the model jumped straight into a tool call: keep the opener on the content side
            return t[:imp].strip(), t[imp:]
        if cut < 0:
            return t.strip(), ""          # truncated while still thinking
        return t[:cut].strip(), self._unwrap(t[cut + len(self.end):].lstrip("\n"))

    def parse_stream_chunk(self, chunk: str) -> Tuple[str, str]:
        """returns (reasoning_delta, content_delta)"""
        self._buf += chunk
        buf = self._buf
        # hold back the longest tail that is a proper prefix of a marker (it may complete in the next chunk)
        hold = 0
        marks = [self.start, self.end] + ([self.implicit_end] if self.implicit_end else []) + (list(self.answer_wrap) if self.answer_wrap else [])
        for tok in marks:
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
            cut = text.find(self.end)
            imp = text.find(self.implicit_end) if self.implicit_end else -1
            if imp >= 0 and (cut < 0 or imp < cut):
                self.in_reasoning = False
                r_out, c_out = text[:imp], text[imp:]
            elif cut >= 0:
                self.in_reasoning = False
                r_out, c_out = text[:cut], self._unwrap(text[cut + len(self.end):].lstrip("\n"))
            else:
                r_out = text
        else:
            c_out = self._unwrap(text)
        if not self.stream_reasoning and self.in_reasoning:
            self._buf = r_out + self._buf
            r_out = ""
        return r_out, c_out


class HarmonyReasoningDetector:
    """GPT-OSS Harmony: ``<|channel|>analysis<|message|>…<|end|><|start|>assistant<|channel|>final<|message|>…<|return|>``.  Messages on the
    ``analysis`` channel are reasoning, ``final`` is the answer; ``commentary`` messages (tool calls / preambles) are left verbatim in the
    content for the function-call parser.  Streaming works message-wise on a small state machine."""

    _HDR = "<|channel|>"
    _MSG = "<|message|>"
    _ENDS = ("<|end|>", "<|return|>", "<|call|>")

    def __init__(self, stream_reasoning: bool = True):
        self.stream_reasoning = stream_reasoning
        self._buf = ""
        self._chan: Optional[str] = None       # channel of the message being streamed (None = between messages)

    def _split(self, text: str):
        """[(channel | None, body, verbatim)] for every complete or trailing message"""
        out, pos = [], 0
        while True:
            h = text.find(self._HDR, pos)
            if h < 0:
                if text[pos:]:
                    out.append((None, text[pos:], text[pos:]))
                return out
            if text[pos:h].replace("<|start|>assistant", "").strip():
                out.append((None, text[pos:h], text[pos:h]))
            m = text.find(self._MSG, h)
            if m < 0:
                out.append(("?", "", text[h:]))
                return out
            chan = text[h + len(self._HDR):m].split()[0] if text[h + len(self._HDR):m].split() else ""
            ends = [(text.find(e, m), e) for e in self._ENDS if text.find(e, m) >= 0]
            if not ends:
                out.append((chan, text[m + len(self._MSG):], text[h:]))
                return out
            e, tok = min(ends)
            out.append((chan, text[m + len(self._MSG):e], text[h:e + len(tok)]))
            pos = e + len(tok)

    def detect_and_parse(self, text: str) -> Tuple[Optional[str], str]:
        reasoning, content = [], []
        for chan, body, verbatim in self._split(text):
            if chan == "analysis":
                reasoning.append(body)
            elif chan == "final":
                content.append(body)
            elif chan is None:
                content.append(body)
            else:
                content.append(verbatim)           # commentary / unknown channel: the tool-call parser reads the header
        return ("\n".join(r.strip() for r in reasoning) if reasoning else None), "".join(content)

    def parse_stream_chunk(self, chunk: str) -> Tuple[str, str]:
        self._buf += chunk
        r_out, c_out = "", ""
        while self._buf:
            if self._chan is None:                                   # between messages: need a full header to decide
                h = self._buf.find(self._HDR)
                if h < 0:
                    t, S = self._buf.lstrip(), "<|start|>assistant"
                    if S.startswith(t):                              # (part of) the next message's role header: wait
                        break
                    if t.startswith(S):
                        rest = t[len(S):]
                        if self._HDR.startswith(rest):               # role header complete, channel header still arriving
                            break
                        self._buf = t = rest                         # a bare role header followed by plain text
                    hold = next((k for k in range(min(len(S), len(t)), 0, -1) if t.endswith(S[:k]) or t.endswith(self._HDR[:k])), 0)
                    c_out += t[:len(t) - hold]
                    self._buf = t[len(t) - hold:] if hold else ""
                    break
                pre = self._buf[:h].replace("<|start|>assistant", "")
                if pre.strip():
                    c_out += pre
                m = self._buf.find(self._MSG, h)
                if m < 0:
                    break                                            # header not complete yet
                names = self._buf[h + len(self._HDR):m].split()
                self._chan = names[0] if names else ""
                if self._chan not in ("analysis", "final"):
                    c_out += self._buf[h:m + len(self._MSG)]         # pass the header through
                self._buf = self._buf[m + len(self._MSG):]
                continue
            ends = [(self._buf.find(e), e) for e in self._ENDS if self._buf.find(e) >= 0]
            if ends:
                e, tok = min(ends)
                body, self._buf = self._buf[:e], self._buf[e + len(tok):]
                done = True
            else:
                hold = max((k for t in self._ENDS for k in range(min(len(t) - 1, len(self._buf)), 0, -1) if self._buf.endswith(t[:k])), default=0)
                body, self._buf = self._buf[:len(self._buf) - hold], (self._buf[len(self._buf) - hold:] if hold else "")
                done = False
            if self._chan == "analysis":
                r_out += body
            elif self._chan == "final":
                c_out += body
            else:
                c_out += body + (tok if done else "")
            if done:
                self._chan = None
            else:
                break
        if not self.stream_reasoning and r_out:
            self._held = getattr(self, "_held", "") + r_out
            r_out = ""
        elif getattr(self, "_held", "") and self.stream_reasoning:
            r_out, self._held = self._held + r_out, ""
        return r_out, c_out


_FAMILIES = {
    "deepseek-r1": dict(start="<think>", end="</think>", force_reasoning=True),
    "deepseek-v3": dict(start="<think>", end="</think>", force_reasoning=False),          # V3.1 hybrid: thinking is switched per request
    "qwen3": dict(start="<think>", end="</think>", force_reasoning=False),
    "qwen3-thinking": dict(start="<think>", end="</think>", force_reasoning=True),
    "kimi": dict(start="◁think▷", end="◁/think▷", force_reasoning=False),
    "kimi-k2": dict(start="<think>", end="</think>", force_reasoning=False, implicit_end="<|tool_calls_section_begin|>"),
    "glm45": dict(start="<think>", end="</think>", force_reasoning=False, implicit_end="<tool_call>"),
    "step3": dict(start="<think>", end="</think>", force_reasoning=True),
    "nemotron3": dict(start="<think>", end="</think>", force_reasoning=True),
    "minimax": dict(start="<think>", end="</think>", force_reasoning=False),
    "minimax-append-think": dict(start="<think>", end="</think>", force_reasoning=True),   # the template already emitted the opener
    "mistral": dict(start="[THINK]", end="[/THINK]", force_reasoning=False),                # Magistral
    "hunyuan": dict(start="<think>", end="</think>", force_reasoning=False, answer_wrap=("<answer>", "</answer>")),
    "mimo": dict(start="<think>", end="</think>", force_reasoning=False),
    "seed-oss": dict(start="<seed:think>", end="</seed:think>", force_reasoning=False),
    "gpt-oss": "harmony",
}


class ReasoningParser:
    def __init__(self, model_type: str = "qwen3", stream_reasoning: bool = True, force_reasoning: Optional[bool] = None):
        if model_type not in _FAMILIES:
            raise ValueError(f"unknown reasoning parser {model_type!r}; available {sorted(_FAMILIES)}")
        if _FAMILIES[model_type] == "harmony":
            self.detector = HarmonyReasoningDetector(stream_reasoning=stream_reasoning)
            return
        kw = dict(_FAMILIES[model_type])
        if force_reasoning is not None:
            kw["force_reasoning"] = force_reasoning
        self.detector = BaseReasoningDetector(stream_reasoning=stream_reasoning, **kw)

    def parse_non_stream(self, text: str) -> Tuple[Optional[str], str]:
        return self.detector.detect_and_parse(text)

    def parse_stream_chunk(self, chunk: str) -> Tuple[str, str]:
        return self.detector.parse_stream_chunk(chunk)
