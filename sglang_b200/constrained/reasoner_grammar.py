"""Grammar that only starts constraining AFTER the model has closed its reasoning block.

Parity: reference ``python/sglang/srt/constrained/reasoner_grammar_backend.py``: with ``--reasoning-parser`` set, a request's json / regex grammar
applies to the *answer*; the chain of thought between ``<think>`` and ``</think>`` stays free text.  The wrapper watches the accepted token ids
for the think-end marker (a token id when the marker is one token, otherwise the marker's tokenisation as a suffix of the output).
"""
from __future__ import annotations

from typing import Any, List, Optional

import numpy as np

from .base_grammar import BaseGrammar


class ReasonerGrammar(BaseGrammar):
    def __init__(self, inner: BaseGrammar, end_ids: List[int], in_reasoning: bool = True):
        self.inner = inner
        self.end_ids = list(end_ids)
        self.in_reasoning = in_reasoning and bool(end_ids)
        self._tail: List[int] = []

    def accept_token(self, token: int) -> None:
        if not self.in_reasoning:
            self.inner.accept_token(token)
            return
        self._tail = (self._tail + [int(token)])[-len(self.end_ids):]
        if self._tail == self.end_ids:
            self.in_reasoning = False

    def fill_vocab_mask(self, mask_row: np.ndarray, vocab_size: int) -> None:
        if self.in_reasoning:
            mask_row[:] = -1            # every token allowed while thinking
        else:
            self.inner.fill_vocab_mask(mask_row, vocab_size)

    def is_terminated(self) -> bool:
        return (not self.in_reasoning) and self.inner.is_terminated()

    @property
    def finished(self) -> bool:          # type: ignore[override]
        return (not self.in_reasoning) and bool(getattr(self.inner, "finished", False))

    def try_jump_forward(self, tokenizer: Any) -> Optional[str]:
        return None if self.in_reasoning else self.inner.try_jump_forward(tokenizer)

    def jump_forward_tokens(self, jf: str, tokenizer: Any, output_ids: List[int]) -> List[int]:
        return [] if self.in_reasoning else self.inner.jump_forward_tokens(jf, tokenizer, output_ids)

    def copy(self) -> "ReasonerGrammar":
        g = ReasonerGrammar(self.inner.copy(), self.end_ids, self.in_reasoning)
        g._tail = list(self._tail)
        return g


def think_end_ids(tokenizer: Any, reasoning_parser: Optional[str]) -> List[int]:
    """token ids of the think-end marker of this parser family ([] when reasoning is not configured)"""
    if not reasoning_parser or tokenizer is None:
        return []
    from ..parser.reasoning_parser import _FAMILIES

    end = _FAMILIES.get(reasoning_parser, {}).get("end")
    if not end:
        return []
    try:
        tid = tokenizer.convert_tokens_to_ids(end)
        if isinstance(tid, int) and tid >= 0 and tid != getattr(tokenizer, "unk_token_id", -1):
            return [tid]
    except Exception:
        pass
    return [int(t) for t in tokenizer.encode(end, add_special_tokens=False)] if hasattr(tokenizer, "encode") else []
