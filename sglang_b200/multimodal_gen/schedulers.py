"""Sampling schedules of the diffusion runtime (reference: ``multimodal_gen/runtime/models/schedulers/*``)."""
from __future__ import annotations

import math
from typing import List

import torch


class FlowMatchEuler:
    """Rectified-flow Euler sampler: x_t = (1 - s) x0 + s * noise, the model predicts v = noise - x0; ``shift`` is the
    SD3 / Flux resolution-dependent time shift  s' = shift * s / (1 + (shift - 1) s)."""

    def __init__(self, num_train_timesteps: int = 1000, shift: float = 1.0):
        self.T, self.shift = num_train_timesteps, shift

    def set_timesteps(self, n: int) -> None:
        s = torch.linspace(1.0, 1.0 / self.T, n, dtype=torch.float64)
        s = self.shift * s / (1 + (self.shift - 1) * s)
        self.sigmas = torch.cat([s, torch.zeros(1, dtype=torch.float64)])
        self.timesteps = (s * self.T).float()

    def init_noise_sigma(self) -> float:
        return 1.0

    def step(self, model_out: torch.Tensor, i: int, x: torch.Tensor) -> torch.Tensor:
        return (x.float() + float(self.sigmas[i + 1] - self.sigmas[i]) * model_out.float()).to(x.dtype)


class DDIM:
    """Deterministic DDIM (eta = 0) for epsilon-prediction models with a scaled-linear beta schedule."""

    def __init__(self, num_train_timesteps: int = 1000, beta_start: float = 0.00085, beta_end: float = 0.012):
        betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float64) ** 2
        self.acp = torch.cumprod(1 - betas, dim=0)
        self.T = num_train_timesteps

    def set_timesteps(self, n: int) -> None:
        step = self.T // n
        self.timesteps = (torch.arange(n - 1, -1, -1) * step).float()
        self._prev = self.timesteps - step

    def init_noise_sigma(self) -> float:
        return 1.0

    def step(self, eps: torch.Tensor, i: int, x: torch.Tensor) -> torch.Tensor:
        t, tp = int(self.timesteps[i]), int(self._prev[i])
        a, ap = float(self.acp[t]), float(self.acp[tp]) if tp >= 0 else 1.0
        x0 = (x.float() - math.sqrt(1 - a) * eps.float()) / math.sqrt(a)
        return (math.sqrt(ap) * x0 + math.sqrt(1 - ap) * eps.float()).to(x.dtype)


def make_scheduler(prediction: str, shift: float = 1.0):
    return FlowMatchEuler(shift=shift) if prediction == "flow" else DDIM()


class TeaCache:
    """Step-skipping cache (reference ``multimodal_gen/runtime/cache``): when the timestep-modulated input of the first block
    has moved less than ``thresh`` (accumulated relative L1) since the last real forward, reuse that forward's residual."""

    def __init__(self, thresh: float):
        self.thresh, self.acc, self.prev_sig, self.residual = thresh, 0.0, None, None
        self.skipped: List[int] = []

    def should_skip(self, sig: torch.Tensor, i: int, last: bool) -> bool:
        if self.prev_sig is None or self.residual is None or last or self.thresh <= 0:
            self.prev_sig, self.acc = sig, 0.0
            return False
        self.acc += float((sig - self.prev_sig).abs().mean() / self.prev_sig.abs().mean().clamp(min=1e-8))
        self.prev_sig = sig
        if self.acc < self.thresh:
            self.skipped.append(i)
            return True
        self.acc = 0.0
        return False
