"""Own regex/JSON-schema grammar backend: regex -> byte-level DFA -> per-state token masks, with *compressed-FSM
jump-forward* (a run of DFA states with a single outgoing byte is collapsed into one forced string).

This is the original SGLang v0.2 technique ("3x faster JSON decoding", reference README.md:32;
``constrained/outlines_jump_forward.py:142``); the reference still exposes ``try_jump_forward`` on grammar objects
(``base_grammar_backend.py:120-147``) but no longer calls it from the scheduler — ours does (``scheduler._try_jump_forward``).
Regex parsing uses ``interegular`` (pure python, in the image); characters are bytes (latin-1 subset of UTF-8).
"""
from __future__ import annotations

import json
from typing import Any, Dict, List, Optional, Tuple

import numpy as np

from .base_grammar import BaseGrammar, BaseGrammarBackend, register_grammar_backend, set_bits

# ---------------------------------------------------------------- JSON schema -> regex
_WS = r"[ \n\t]*"
_STRING_INNER = r'([^"\\\x00-\x1f\x7f-\x9f]|\\["\\/bfnrt])'
_STRING = f'"{_STRING_INNER}*"'
_INTEGER = r"(-)?(0|[1-9][0-9]*)"
_NUMBER = rf"({_INTEGER})(\.[0-9]+)?([eE][+-][0-9]+)?"
_BOOLEAN = r"(true|false)"
_NULL = r"null"


def _esc(s: str) -> str:
    out = []
    for ch in s:
        out.append("\\" + ch if ch in r".^$*+?{}[]\|()/" else ch)
    return "".join(out)


def schema_to_regex(schema: Any, ws: str = _WS, depth: int = 0) -> str:
    if isinstance(schema, str):
        schema = json.loads(schema)
    if depth > 12:
        raise ValueError("schema nesting too deep")
    if schema is True or schema == {}:
        return rf"({_STRING}|{_NUMBER}|{_BOOLEAN}|{_NULL})"
    if "enum" in schema:
        return "(" + "|".join(_esc(json.dumps(v)) for v in schema["enum"]) + ")"
    if "const" in schema:
        return _esc(json.dumps(schema["const"]))
    for key in ("anyOf", "oneOf"):
        if key in schema:
            return "(" + "|".join(schema_to_regex(s, ws, depth + 1) for s in schema[key]) + ")"
    t = schema.get("type")
    if isinstance(t, list):
        return "(" + "|".join(schema_to_regex({**schema, "type": x}, ws, depth + 1) for x in t) + ")"
    if t == "string":
        if "pattern" in schema:
            p = schema["pattern"]
            return '"' + p.lstrip("^").rstrip("$") + '"'
        lo, hi = schema.get("minLength"), schema.get("maxLength")
        if lo is not None or hi is not None:
            return f'"{_STRING_INNER}{{{lo or 0},{hi if hi is not None else ""}}}"'
        return _STRING
    if t == "integer":
        return _INTEGER
    if t == "number":
        return _NUMBER
    if t == "boolean":
        return _BOOLEAN
    if t == "null":
        return _NULL
    if t == "array":
        item = schema_to_regex(schema.get("items", {}), ws, depth + 1)
        lo = schema.get("minItems", 0)
        hi = schema.get("maxItems")
        rest = f"({ws},{ws}{item})"
        if hi is not None:
            body = f"{item}{rest}{{{max(0, lo - 1)},{hi - 1}}}"
        else:
            body = f"{item}{rest}{{{max(0, lo - 1)},}}" if lo > 1 else f"{item}{rest}*"
        return rf"\[{ws}({body}){'' if lo > 0 else '?'}{ws}\]"
    if t == "object" or "properties" in schema:
        props = schema.get("properties", {})
        req = set(schema.get("required", list(props)))
        parts = []
        for i, (k, sub) in enumerate(props.items()):
            kv = f'"{_esc(k)}"{ws}:{ws}{schema_to_regex(sub, ws, depth + 1)}'
            parts.append((kv, k in req))
        if not parts:
            return rf"\{{{ws}\}}"
        # required-first canonical order; optional properties may be omitted
        body = ""
        first = True
        for kv, required in parts:
            sep = "" if first else f"{ws},{ws}"
            if required:
                body += sep + kv
                first = False
            else:
                body += f"({sep}{kv})?" if not first else f"({kv})?"
        return rf"\{{{ws}{body}{ws}\}}"
    raise ValueError(f"unsupported JSON schema fragment: {schema}")


# ---------------------------------------------------------------- byte DFA
class ByteDFA:
    def __init__(self, regex: str):
        import interegular

        fsm = interegular.parse_pattern(regex).to_fsm().reduce()
        self.initial = fsm.initial
        self.finals = set(fsm.finals)
        n = max(list(fsm.states) + [0]) + 1
        self.n_states = n
        self.trans = np.full((n, 256), -1, dtype=np.int32)
        alpha = fsm.alphabet
        any_sym = alpha.get(interegular.fsm.anything_else) if hasattr(interegular.fsm, "anything_else") else None
        for b in range(256):
            ch = chr(b)
            sym = alpha[ch] if ch in alpha else any_sym
            if sym is None:
                continue
            for s, m in fsm.map.items():
                d = m.get(sym)
                if d is not None:
                    self.trans[s, b] = d
        # states from which a final state is reachable (live states)
        live = np.zeros(n, dtype=bool)
        live[list(self.finals)] = True
        changed = True
        while changed:
            nxt = (self.trans >= 0) & live[np.clip(self.trans, 0, n - 1)]
            new = live | nxt.any(1)
            changed = bool((new != live).any())
            live = new
        dead = ~live
        self.trans[(self.trans >= 0) & dead[np.clip(self.trans, 0, n - 1)]] = -1

    def forced_run(self, state: int, limit: int = 512) -> Tuple[bytes, int]:
        """longest run of single-choice transitions from ``state`` (stops at final states: the string may end there)"""
        out = bytearray()
        while len(out) < limit and state not in self.finals:
            nz = np.flatnonzero(self.trans[state] >= 0)
            if nz.size != 1:
                break
            out.append(int(nz[0]))
            state = int(self.trans[state, nz[0]])
        return bytes(out), state


class RegexFSMGrammar(BaseGrammar):
    def __init__(self, backend: "RegexFSMBackend", dfa: ByteDFA, mask_cache: Dict[int, np.ndarray], state: Optional[int] = None):
        self.b, self.dfa, self.cache = backend, dfa, mask_cache
        self.state = dfa.initial if state is None else state
        self.done = False

    def _walk(self, state: int, data: bytes) -> int:
        for x in data:
            if state < 0:
                return -1
            state = int(self.dfa.trans[state, x])
        return state

    def accept_token(self, token: int) -> None:
        if self.done:
            return
        if token in self.b.eos_ids:
            if self.state not in self.dfa.finals:
                raise ValueError("EOS before the grammar completed")
            self.done = True
            return
        ns = self._walk(self.state, self.b.token_bytes[token])
        if ns < 0:
            raise ValueError(f"token {token} rejected by the grammar")
        self.state = ns

    def _allowed(self, state: int) -> np.ndarray:
        m = self.cache.get(state)
        if m is None:
            b = self.b
            ok = np.zeros(b.vocab_size, dtype=bool)
            # single-byte tokens vectorised, the rest walked
            ok[b.single_ids] = self.dfa.trans[state, b.single_bytes] >= 0
            for t in b.multi_ids:
                ok[t] = self._walk(state, b.token_bytes[t]) >= 0
            if state in self.dfa.finals:
                ok[list(b.eos_ids)] = True
            m = self.cache[state] = ok
        return m

    def fill_vocab_mask(self, mask_row: np.ndarray, vocab_size: int) -> None:
        if self.done:
            mask_row[:] = -1
            return
        set_bits(mask_row, self._allowed(self.state)[:vocab_size])

    def is_terminated(self) -> bool:
        if self.done:
            return True
        # a final state with no outgoing transition admits only EOS
        return self.state in self.dfa.finals and not (self.dfa.trans[self.state] >= 0).any()

    def try_jump_forward(self, tokenizer: Any) -> Optional[str]:
        if self.done:
            return None
        data, _ = self.dfa.forced_run(self.state)
        if len(data) < 2:      # a single forced byte is not worth a re-prefill
            return None
        return data.decode("latin-1")

    def jump_forward_tokens(self, jf: str, tokenizer: Any, output_ids: List[int]) -> List[int]:
        data = jf.encode("latin-1")
        if self.b.byte_level:
            return [self.b.byte_to_id[x] for x in data if x in self.b.byte_to_id]
        return super().jump_forward_tokens(jf, tokenizer, output_ids)

    def copy(self) -> "RegexFSMGrammar":
        g = RegexFSMGrammar(self.b, self.dfa, self.cache, self.state)
        g.done = self.done
        return g


class RegexFSMBackend(BaseGrammarBackend):
    def __init__(self, tokenizer: Any, vocab_size: int):
        self.vocab_size = vocab_size
        self.tokenizer = tokenizer
        eos = getattr(tokenizer, "eos_token_id", None)
        self.eos_ids = {int(eos)} if eos is not None else set()
        self.token_bytes: List[bytes] = []
        special = set(getattr(tokenizer, "all_special_ids", []) or [])
        for i in range(vocab_size):
            if i in special:
                self.token_bytes.append(b"\xff\xfe\xfd")   # never matches
                continue
            try:
                s = tokenizer.decode([i], skip_special_tokens=False)
                self.token_bytes.append(s.encode("latin-1", errors="replace") if s else b"\xff\xfe\xfd")
            except Exception:
                self.token_bytes.append(b"\xff\xfe\xfd")
        self.single_ids = np.asarray([i for i, t in enumerate(self.token_bytes) if len(t) == 1], dtype=np.int64)
        self.single_bytes = np.asarray([self.token_bytes[i][0] for i in self.single_ids], dtype=np.int64)
        self.multi_ids = [i for i, t in enumerate(self.token_bytes) if len(t) != 1 and t != b"\xff\xfe\xfd"]
        self.byte_to_id = {int(b): int(i) for i, b in zip(self.single_ids[::-1], self.single_bytes[::-1])}
        self.byte_level = type(tokenizer).__name__ == "ByteTokenizer"
        self._dfa_cache: Dict[str, Tuple[ByteDFA, Dict[int, np.ndarray]]] = {}

    def compile(self, kind: str, spec: str) -> BaseGrammar:
        try:
            if kind == "json":
                regex = schema_to_regex(json.loads(spec) if spec not in ("", "$$ANY$$") else {"type": "object"})
            elif kind == "regex":
                regex = spec
            else:
                raise ValueError(f"the regex_fsm backend supports json / regex (got {kind}); use --grammar-backend xgrammar")
            if regex not in self._dfa_cache:
                self._dfa_cache[regex] = (ByteDFA(regex), {})
        except ValueError:
            raise
        except Exception as e:
            raise ValueError(f"invalid {kind} grammar: {type(e).__name__}: {e}") from e
        dfa, cache = self._dfa_cache[regex]
        return RegexFSMGrammar(self, dfa, cache)


register_grammar_backend("regex_fsm", RegexFSMBackend)
