"""Weight loading.

Parity: reference ``python/sglang/srt/model_loader/loader.py`` (``DefaultModelLoader:357`` safetensors / pt,
``DummyModelLoader:1602`` random init of the real architecture, ``ShardedStateLoader``) and ``weight_utils.py``.

``dummy``: every parameter is filled *directly on the target device* from a per-parameter seeded generator — no
host staging, so a 70B random init costs seconds, and all TP ranks agree on replicated tensors.
"""
from __future__ import annotations

import glob
import json
import os
import zlib
from typing import Dict, Iterable, Iterator, Optional, Tuple

import torch
from torch import nn

from ..configs.model_config import ModelConfig
from ..models.registry import get_model_class


def _default_device_ctx(device: torch.device):
    return torch.device(device)


def build_model(cfg: ModelConfig, device: torch.device, quant: Optional[str] = None) -> nn.Module:
    cls = get_model_class(cfg.architecture)
    with torch.device(device):
        model = cls(cfg, quant, device=device)
    return model.eval()


@torch.no_grad()
def init_dummy_weights(model: nn.Module, seed: int = 0, scale: float = 0.02) -> None:
    """N(0, scale) for matrices, ones for norm weights (zeros for gemma-style), zeros for biases.
    The seed of each tensor is a hash of its *global* name so results do not depend on sharding order."""
    for name, p in model.named_parameters():
        if p.dim() <= 1:
            if name.endswith("bias"):
                p.zero_()
            elif "norm" in name:
                gemma = getattr(dict(model.named_modules()).get(name.rsplit(".", 1)[0]), "gemma", False)
                p.fill_(0.0 if gemma else 1.0)
            else:
                p.normal_(0.0, scale)
            continue
        g = torch.Generator(device=p.device)
        g.manual_seed((seed * 1000003 + zlib.crc32(name.encode())) & 0x7FFFFFFF)
        if p.dtype in (torch.float32, torch.bfloat16, torch.float16):
            p.normal_(0.0, scale, generator=g)
        else:
            tmp = torch.empty(p.shape, dtype=torch.float32, device=p.device).normal_(0.0, scale, generator=g)
            p.copy_(tmp.to(p.dtype))


def iter_safetensors(path: str) -> Iterator[Tuple[str, torch.Tensor]]:
    from safetensors import safe_open

    files = sorted(glob.glob(os.path.join(path, "*.safetensors")))
    if not files:
        raise FileNotFoundError(f"no *.safetensors under {path}")
    for f in files:
        with safe_open(f, framework="pt", device="cpu") as sf:
            for k in sf.keys():
                yield k, sf.get_tensor(k)


def iter_pt(path: str) -> Iterator[Tuple[str, torch.Tensor]]:
    files = sorted(glob.glob(os.path.join(path, "*.bin")) + glob.glob(os.path.join(path, "*.pt")))
    for f in files:
        sd = torch.load(f, map_location="cpu", weights_only=True)
        yield from sd.items()


def load_model(cfg: ModelConfig, model_path: str, load_format: str, device: torch.device, quant: Optional[str] = None,
               seed: int = 0) -> nn.Module:
    model = build_model(cfg, device, quant)
    if load_format == "dummy":
        init_dummy_weights(model, seed)
    elif load_format in ("safetensors", "auto"):
        model.load_weights(iter_safetensors(model_path))
    elif load_format == "pt":
        model.load_weights(iter_pt(model_path))
    elif load_format == "sharded_state":
        load_sharded_state(model, model_path)
    else:
        raise ValueError(f"unsupported load format {load_format}")
    process_weights_after_loading(model)
    return model


def process_weights_after_loading(model: nn.Module) -> None:
    for m in model.modules():
        q = getattr(m, "quant", None)
        if q is not None:
            q.process_weights_after_loading(m)


def save_sharded_state(model: nn.Module, path: str, rank: int) -> None:
    """pre-sharded checkpoint, one file per rank (reference ``Engine.save_sharded_model``)"""
    from safetensors.torch import save_file

    os.makedirs(path, exist_ok=True)
    sd = {k: v.detach().cpu().contiguous() for k, v in model.state_dict().items()}
    save_file(sd, os.path.join(path, f"rank-{rank:03d}.safetensors"))


def load_sharded_state(model: nn.Module, path: str, rank: Optional[int] = None) -> None:
    from safetensors.torch import load_file

    from ..parallel.parallel_state import get_state

    rank = get_state().rank if rank is None else rank
    sd = load_file(os.path.join(path, f"rank-{rank:03d}.safetensors"))
    missing = model.load_state_dict(sd, strict=False)
    if missing.unexpected_keys:
        raise RuntimeError(f"unexpected keys in sharded state: {missing.unexpected_keys[:5]}")
