"""Per-request sampling parameters.

Field list and defaults follow the reference contract (SURVEY Appendix A.2;
``python/sglang/srt/sampling/sampling_params.py``): ``max_new_tokens=128, temperature=1.0,
top_p=1.0, top_k=-1, min_p=0.0, penalties, stop / stop_token_ids / stop_regex,
json_schema | regex | ebnf | structural_tag (mutually exclusive), n, ignore_eos, ...``.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Dict, List, Optional, Union

_SAMPLING_EPS = 1e-6
TOP_K_ALL = 1 << 30


@dataclass
class SamplingParams:
    max_new_tokens: int = 128
    min_new_tokens: int = 0
    stop: Optional[Union[str, List[str]]] = None
    stop_token_ids: Optional[List[int]] = None
    stop_regex: Optional[Union[str, List[str]]] = None
    temperature: float = 1.0
    top_p: float = 1.0
    top_k: int = -1
    min_p: float = 0.0
    frequency_penalty: float = 0.0
    presence_penalty: float = 0.0
    repetition_penalty: float = 1.0
    n: int = 1
    json_schema: Optional[str] = None
    regex: Optional[str] = None
    ebnf: Optional[str] = None
    structural_tag: Optional[str] = None
    ignore_eos: bool = False
    skip_special_tokens: bool = True
    spaces_between_special_tokens: bool = True
    no_stop_trim: bool = False
    stream_interval: Optional[int] = None
    logit_bias: Optional[Dict[str, float]] = None
    sampling_seed: Optional[int] = None
    custom_params: Optional[Dict[str, Any]] = None

    def __post_init__(self) -> None:
        if isinstance(self.stop, str):
            self.stop = [self.stop]
        self.stop = list(self.stop or [])
        if isinstance(self.stop_regex, str):
            self.stop_regex = [self.stop_regex]
        self.stop_regex = list(self.stop_regex or [])
        self.stop_token_ids = list(self.stop_token_ids or [])
        if isinstance(self.json_schema, dict):
            import json

            self.json_schema = json.dumps(self.json_schema)
        # greedy normalisation (reference: temperature < eps -> top_k = 1)
        if 0.0 <= self.temperature < _SAMPLING_EPS:
            self.temperature = 1.0
            self.top_k = 1
        if self.top_k == -1 or self.top_k is None:
            self.top_k = TOP_K_ALL
        self.verify()

    @property
    def is_greedy(self) -> bool:
        return self.top_k == 1

    def verify(self, vocab_size: Optional[int] = None) -> None:
        if self.temperature < 0.0:
            raise ValueError(f"temperature must be non-negative, got {self.temperature}")
        if not 0.0 < self.top_p <= 1.0:
            raise ValueError(f"top_p must be in (0, 1], got {self.top_p}")
        if not 0.0 <= self.min_p <= 1.0:
            raise ValueError(f"min_p must be in [0, 1], got {self.min_p}")
        if self.top_k < 1:
            raise ValueError(f"top_k must be -1 (all) or >= 1, got {self.top_k}")
        for name in ("frequency_penalty", "presence_penalty"):
            v = getattr(self, name)
            if not -2.0 <= v <= 2.0:
                raise ValueError(f"{name} must be in [-2, 2], got {v}")
        if not 0.0 < self.repetition_penalty <= 2.0:
            raise ValueError("repetition_penalty must be in (0, 2]")
        if self.max_new_tokens is not None and self.max_new_tokens < 0:
            raise ValueError("max_new_tokens must be >= 0")
        if not 0 <= self.min_new_tokens <= (self.max_new_tokens if self.max_new_tokens is not None else 1 << 30):
            raise ValueError("min_new_tokens must be in [0, max_new_tokens]")
        grammars = [self.json_schema, self.regex, self.ebnf, self.structural_tag]
        if sum(g is not None for g in grammars) > 1:
            raise ValueError("only one of json_schema / regex / ebnf / structural_tag can be set")
        if self.n < 1:
            raise ValueError("n must be >= 1")
        if vocab_size is not None and self.logit_bias:
            for k in self.logit_bias:
                if not 0 <= int(k) < vocab_size:
                    raise ValueError(f"logit_bias token id {k} out of range")

    @property
    def has_grammar(self) -> bool:
        return any(g is not None for g in (self.json_schema, self.regex, self.ebnf, self.structural_tag))

    def grammar_key(self):
        if self.json_schema is not None:
            return ("json", self.json_schema)
        if self.regex is not None:
            return ("regex", self.regex)
        if self.ebnf is not None:
            return ("ebnf", self.ebnf)
        if self.structural_tag is not None:
            return ("structural_tag", self.structural_tag)
        return None

    @classmethod
    def from_dict(cls, d: Optional[Dict[str, Any]]) -> "SamplingParams":
        d = dict(d or {})
        known = {f for f in cls.__dataclass_fields__}
        extra = {k: d.pop(k) for k in list(d) if k not in known}
        if extra:
            d.setdefault("custom_params", {}).update(extra)
        return cls(**d)

    def to_dict(self) -> Dict[str, Any]:
        import dataclasses

        d = dataclasses.asdict(self)
        if d["top_k"] == TOP_K_ALL:
            d["top_k"] = -1
        return d
