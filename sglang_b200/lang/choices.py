"""``select`` decision rules over per-choice token logprobs (reference ``python/sglang/lang/choices.py:32-110``)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Dict, List, Optional

import numpy as np


@dataclass
class ChoicesDecision:
    decision: str
    meta_info: Optional[Dict[str, Any]] = None


class ChoicesSamplingMethod:
    requires_unconditional_logprobs = False

    def __call__(self, *, choices: List[str], normalized_prompt_logprobs: List[float], input_token_logprobs: List[List[Any]],
                 output_token_logprobs: List[List[Any]], unconditional_token_logprobs: Optional[List[List[Any]]] = None) -> ChoicesDecision:
        raise NotImplementedError


class TokenLengthNormalized(ChoicesSamplingMethod):
    def __call__(self, *, choices, normalized_prompt_logprobs, input_token_logprobs, output_token_logprobs, unconditional_token_logprobs=None):
        best = int(np.argmax(normalized_prompt_logprobs))
        return ChoicesDecision(choices[best], {"normalized_prompt_logprobs": normalized_prompt_logprobs,
                                               "input_token_logprobs": input_token_logprobs})


class GreedyTokenSelection(ChoicesSamplingMethod):
    def __call__(self, *, choices, normalized_prompt_logprobs, input_token_logprobs, output_token_logprobs, unconditional_token_logprobs=None):
        n = len(choices)
        L = max(len(x) for x in input_token_logprobs)
        mat = np.full((n, L), np.nan)
        for i, x in enumerate(input_token_logprobs):
            vals = [t[0] for t in x]
            mat[i, : len(vals)] = vals
            if len(vals) < L:
                mat[i, len(vals):] = np.mean(vals) if vals else 0.0
        alive = np.arange(n)
        for j in range(L):
            col = mat[alive, j]
            alive = alive[col == col.max()]
            if len(alive) == 1:
                break
        return ChoicesDecision(choices[int(alive[0])], {"normalized_prompt_logprobs": normalized_prompt_logprobs})


class UnconditionalLikelihoodNormalized(ChoicesSamplingMethod):
    requires_unconditional_logprobs = True

    def __call__(self, *, choices, normalized_prompt_logprobs, input_token_logprobs, output_token_logprobs, unconditional_token_logprobs=None):
        assert unconditional_token_logprobs is not None
        scores = []
        for c, u in zip(input_token_logprobs, unconditional_token_logprobs):
            cv = np.array([t[0] for t in c], dtype=float)
            uv = np.array([t[0] if t[0] is not None else 0.0 for t in u][-len(cv):], dtype=float)
            scores.append(float(np.mean(cv - uv)) if len(cv) else -1e9)
        return ChoicesDecision(choices[int(np.argmax(scores))], {"normalized_unconditional_prompt_logprobs": scores})


token_length_normalized = TokenLengthNormalized()
greedy_token_selection = GreedyTokenSelection()
unconditional_likelihood_normalized = UnconditionalLikelihoodNormalized()
