"""Grammar object / backend contracts.

Parity: reference ``python/sglang/srt/constrained/base_grammar_backend.py`` (``BaseGrammarObject``:
``accept_token``, ``fill_vocab_mask``, ``try_jump_forward:120-147``, ``jump_forward_str_state``, ``copy``; backend
registry ``:347``).  The vocab mask is the xgrammar-convention packed bitmask: int32 words, bit = 1 -> token allowed;
it is applied on the device by ``ops.apply_token_bitmask`` (``csrc/kernels/sampling.cu``).
"""
from __future__ import annotations

from typing import Any, Callable, Dict, List, Optional

import numpy as np


class BaseGrammar:
    finished = False

    def accept_token(self, token: int) -> None:
        raise NotImplementedError

    def fill_vocab_mask(self, mask_row: np.ndarray, vocab_size: int) -> None:
        """write the allowed-token bitmask for the NEXT token into ``mask_row`` (int32 [ceil(V/32)])"""
        raise NotImplementedError

    def is_terminated(self) -> bool:
        return False

    def try_jump_forward(self, tokenizer: Any) -> Optional[str]:
        """the unique string every valid continuation must start with ('' / None if there is none)"""
        return None

    def jump_forward_tokens(self, jf: str, tokenizer: Any, output_ids: List[int]) -> List[int]:
        """tokens to append for the forced string.  Re-tokenising at the boundary can merge with the previous
        token for BPE vocabularies (reference ``jump_forward_str_state``); we keep the conservative choice of
        tokenising the forced string alone and only returning a prefix the grammar accepts token by token."""
        ids = tokenizer.encode(jf) if jf else []
        probe = self.copy()
        out: List[int] = []
        for t in ids:
            try:
                probe.accept_token(int(t))
            except Exception:
                break
            out.append(int(t))
            if probe.is_terminated():
                break
        return out

    def copy(self) -> "BaseGrammar":
        raise NotImplementedError


class BaseGrammarBackend:
    def compile(self, kind: str, spec: str) -> BaseGrammar:
        """kind in {json, regex, ebnf, structural_tag}; raises ValueError on an invalid spec"""
        raise NotImplementedError


_BACKENDS: Dict[str, Callable[..., BaseGrammarBackend]] = {}


def register_grammar_backend(name: str, factory: Callable[..., BaseGrammarBackend]) -> None:
    _BACKENDS[name] = factory


def create_grammar_backend(name: str, tokenizer: Any, vocab_size: int) -> BaseGrammarBackend:
    from . import regex_fsm_backend, xgrammar_backend  # noqa: F401  (register)

    if name == "auto":
        for cand in ("xgrammar", "regex_fsm"):
            try:
                return _BACKENDS[cand](tokenizer, vocab_size)
            except Exception:
                continue
        raise RuntimeError("no grammar backend available")
    if name not in _BACKENDS:
        raise ValueError(f"unknown grammar backend {name!r}; registered: {sorted(_BACKENDS)}")
    return _BACKENDS[name](tokenizer, vocab_size)


def set_bits(mask_row: np.ndarray, allowed: np.ndarray) -> None:
    """allowed: bool [V] -> packed little-endian-bit int32 words"""
    V = allowed.size
    words = mask_row.size
    pad = np.zeros(words * 32, dtype=bool)
    pad[:V] = allowed
    mask_row[:] = np.packbits(pad.reshape(words, 32), axis=1, bitorder="little").view(np.int32).reshape(words)
