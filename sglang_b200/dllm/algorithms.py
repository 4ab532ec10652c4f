"""Block denoising rules.  Each rule sees the block's current ids [B, bs] and logits [B, bs, V] and returns the new ids plus a
per-row ``finished`` flag; rows are independent so a batch can mix blocks at different stages."""
from __future__ import annotations

from typing import Any, Dict, Optional, Type

import torch


class DllmAlgorithm:
    def __init__(self, mask_id: int, config: Optional[Dict[str, Any]] = None):
        self.mask_id, self.config = mask_id, dict(config or {})

    def max_steps(self, block_size: int) -> int:
        return block_size + 1

    def init_state(self, ids: torch.Tensor) -> Dict[str, torch.Tensor]:
        return {}

    def step(self, ids: torch.Tensor, logits: torch.Tensor, state: Dict[str, torch.Tensor]):
        raise NotImplementedError

    @staticmethod
    def _pick(logits: torch.Tensor):
        x = torch.argmax(logits, dim=-1)
        p = torch.gather(torch.softmax(logits.float(), dim=-1), -1, x.unsqueeze(-1)).squeeze(-1)
        return x, p


class LowConfidence(DllmAlgorithm):
    """Commit every masked position whose top-1 probability exceeds ``threshold``; when none does, commit the single most
    confident one (so a block needs at most ``block_size`` forwards, plus the final all-committed forward)."""

    def step(self, ids, logits, state):
        th = float(self.config.get("threshold", 0.95))
        masked = ids == self.mask_id
        done = ~masked.any(dim=1)
        x, p = self._pick(logits)
        conf = torch.where(masked, p, torch.full_like(p, float("-inf")))
        move = conf > th
        force = torch.zeros_like(move)
        force[torch.arange(ids.shape[0]), conf.argmax(dim=1)] = True
        move = torch.where(move.any(dim=1, keepdim=True), move, force) & masked
        return torch.where(move, x, ids), done


class JointThreshold(DllmAlgorithm):
    """Mask-to-token commits above ``threshold`` (forcing the best one) plus token-to-token EDITS of already committed
    positions whose new argmax differs with probability above ``edit_threshold``; a block finishes when a forward changes
    nothing or its post-unmask edit budget ``max_post_edit_steps`` is spent.  ``penalty_lambda`` discourages repeating the
    previous position's token."""

    def max_steps(self, block_size: int) -> int:
        return block_size + int(self.config.get("max_post_edit_steps", 16)) + 1

    def init_state(self, ids):
        B = ids.shape[0]
        return {"finished": torch.zeros(B, dtype=torch.bool, device=ids.device), "post_edit": torch.zeros(B, dtype=torch.long, device=ids.device)}

    def step(self, ids, logits, state):
        th = float(self.config.get("threshold", 0.5))
        eth = float(self.config.get("edit_threshold", 0.0))
        budget = int(self.config.get("max_post_edit_steps", 16))
        lam = float(self.config.get("penalty_lambda", 0.0))
        fin, post = state["finished"], state["post_edit"]
        was_fin = fin.clone()
        if lam > 0:
            logits = logits.clone()
            logits[:, 1:].scatter_add_(2, ids[:, :-1].unsqueeze(-1), torch.full_like(ids[:, :-1], -lam, dtype=logits.dtype).unsqueeze(-1))
        x, p = self._pick(logits)
        masked = ids == self.mask_id
        has_mask = masked.any(dim=1)
        active = ~fin
        no_mask = active & ~has_mask
        post += no_mask.long()
        exceeded = post > budget
        fin |= no_mask & exceeded
        eligible = active & ~(no_mask & exceeded)
        conf = torch.where(masked, p, torch.full_like(p, float("-inf")))
        m2t = (conf > th) & (eligible & has_mask).unsqueeze(1)
        need = (eligible & has_mask) & ~m2t.any(dim=1)
        rows = torch.arange(ids.shape[0], device=ids.device)
        m2t[rows, conf.argmax(dim=1)] |= need
        t2t = (p > eth) & (ids != x) & ~masked & eligible.unsqueeze(1) if eth > 0 else torch.zeros_like(m2t)
        move = m2t | t2t
        fin |= eligible & ~move.any(dim=1)
        # a row is DONE for the worker only when the forward that just ran already had its final ids
        return torch.where(move, x, ids), was_fin | (fin & ~move.any(dim=1) & ~has_mask)


ALGORITHMS: Dict[str, Type[DllmAlgorithm]] = {"LowConfidence": LowConfidence, "JointThreshold": JointThreshold}


def register_algorithm(name: str, cls: Type[DllmAlgorithm]) -> None:
    ALGORITHMS[name] = cls
