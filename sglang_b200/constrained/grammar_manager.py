"""Asynchronous grammar compilation + cache.

Parity: reference ``python/sglang/srt/constrained/grammar_manager.py:26,131,184`` (compile queue polled by the
scheduler; requests wait in ``grammar_queue`` until their grammar is ready; invalid grammars abort the request).
"""
from __future__ import annotations

import logging
from concurrent.futures import Future, ThreadPoolExecutor
from typing import Any, Dict, Optional, Tuple

from ..managers.schedule_batch import FINISH_ABORT
from .base_grammar import BaseGrammar, create_grammar_backend

logger = logging.getLogger(__name__)


class GrammarManager:
    def __init__(self, tokenizer: Any, vocab_size: int, backend: str = "auto", jump_forward_backend: Optional[str] = None,
                 reasoning_parser: Optional[str] = None):
        from .reasoner_grammar import think_end_ids

        self.backend = create_grammar_backend(backend, tokenizer, vocab_size)
        # with a reasoning parser the grammar constrains the answer only (constrained/reasoner_grammar.py)
        GrammarManager.think_end = think_end_ids(tokenizer, reasoning_parser)
        self.backend_name = type(self.backend).__name__
        self.pool = ThreadPoolExecutor(max_workers=2, thread_name_prefix="sglb-grammar")
        self.cache: Dict[Tuple[str, str], Any] = {}     # key -> compiled template grammar | Exception
        self.inflight: Dict[Tuple[str, str], Future] = {}

    def _compile(self, key: Tuple[str, str]) -> Any:
        try:
            return self.backend.compile(key[0], key[1])
        except Exception as e:
            return e

    def attach(self, req: Any) -> bool:
        """True if ``req.grammar`` (or ``req.to_abort``) is set now; False if compilation is in flight."""
        key = req.sampling_params.grammar_key()
        if key is None:
            return True
        key = (key[0], key[1] if isinstance(key[1], str) else str(key[1]))
        req._grammar_key = key
        if key in self.cache:
            self._bind(req, self.cache[key])
            return True
        if key not in self.inflight:
            self.inflight[key] = self.pool.submit(self._compile, key)
        return False

    def poll(self, req: Any) -> bool:
        key = req._grammar_key
        if key in self.cache:
            self._bind(req, self.cache[key])
            return True
        fut = self.inflight.get(key)
        if fut is None or not fut.done():
            return False
        self.cache[key] = fut.result()
        self.inflight.pop(key, None)
        self._bind(req, self.cache[key])
        return True

    @staticmethod
    def _bind(req: Any, compiled: Any) -> None:
        if isinstance(compiled, Exception):
            req.to_abort = FINISH_ABORT(f"invalid grammar: {compiled}", 400)
        else:
            req.grammar = compiled.copy()
            end = getattr(GrammarManager, "think_end", None)
            if end and (getattr(req.sampling_params, "custom_params", None) or {}).get("require_reasoning"):
                from .reasoner_grammar import ReasonerGrammar

                req.grammar = ReasonerGrammar(req.grammar, end)

    def shutdown(self) -> None:
        self.pool.shutdown(wait=False, cancel_futures=True)
