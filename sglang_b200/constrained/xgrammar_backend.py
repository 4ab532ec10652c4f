"""xgrammar adapter (library present in the image).  Reference: ``constrained/xgrammar_backend.py`` (default
backend; ``fill_vocab_mask:126``, ``apply_vocab_mask:258``, ``find_jump_forward_string``)."""
from __future__ import annotations

from typing import Any, List, Optional

import numpy as np

from .base_grammar import BaseGrammar, BaseGrammarBackend, register_grammar_backend


class XGrammar(BaseGrammar):
    def __init__(self, matcher: Any, compiled: Any, xgr: Any):
        self.m, self.compiled, self.xgr = matcher, compiled, xgr

    def accept_token(self, token: int) -> None:
        if self.m.is_terminated():
            return
        if not self.m.accept_token(int(token)):
            raise ValueError(f"token {token} rejected by the grammar")

    def fill_vocab_mask(self, mask_row: np.ndarray, vocab_size: int) -> None:
        if self.m.is_terminated():
            mask_row[:] = -1
            return
        import torch

        t = torch.from_numpy(mask_row.reshape(1, -1))
        self.m.fill_next_token_bitmask(t, 0)

    def is_terminated(self) -> bool:
        return bool(self.m.is_terminated())

    def try_jump_forward(self, tokenizer: Any) -> Optional[str]:
        if self.m.is_terminated():
            return None
        s = self.m.find_jump_forward_string()
        return s or None

    def copy(self) -> "XGrammar":
        return XGrammar(self.m.fork(), self.compiled, self.xgr)


class XGrammarBackend(BaseGrammarBackend):
    def __init__(self, tokenizer: Any, vocab_size: int):
        import xgrammar as xgr

        self.xgr = xgr
        if hasattr(tokenizer, "backend_tokenizer") or hasattr(tokenizer, "vocab_files_names"):
            info = xgr.TokenizerInfo.from_huggingface(tokenizer, vocab_size=vocab_size)
        else:
            vocab: List[bytes] = []
            nbytes = getattr(tokenizer, "_nbytes", 256)
            for i in range(vocab_size):
                vocab.append(bytes([i]) if i < nbytes else f"<extra_{i}>".encode())
            stop = [getattr(tokenizer, "eos_token_id", vocab_size - 1)]
            info = xgr.TokenizerInfo(vocab, xgr.VocabType.RAW, vocab_size=vocab_size, stop_token_ids=stop)
        self.compiler = xgr.GrammarCompiler(info, max_threads=4)

    def compile(self, kind: str, spec: str) -> BaseGrammar:
        c = self.compiler
        try:
            if kind == "json":
                comp = c.compile_builtin_json_grammar() if spec in ("$$ANY$$", "") else c.compile_json_schema(spec)
            elif kind == "regex":
                comp = c.compile_regex(spec)
            elif kind == "ebnf":
                comp = c.compile_grammar(spec)
            elif kind == "structural_tag":
                comp = c.compile_structural_tag(spec) if isinstance(spec, str) else c.compile_structural_tag(**spec)
            else:
                raise ValueError(f"unknown grammar kind {kind}")
        except ValueError:
            raise
        except Exception as e:
            raise ValueError(f"invalid {kind} grammar: {e}") from e
        return XGrammar(self.xgr.GrammarMatcher(comp), comp, self.xgr)


register_grammar_backend("xgrammar", XGrammarBackend)
