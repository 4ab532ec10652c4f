"""Multi-adapter LoRA serving.

Parity: reference ``python/sglang/srt/lora/`` (``LoRAManager`` ``lora_manager.py:65``: dynamic load/unload, a bounded
number of adapters resident per batch, per-token adapter routing; batched shrink/expand (SGMV) kernels
``kernels/ops/gemm/sgemm_lora_{a,b}.py``).

Design: tokens of a batch are grouped by adapter (stable sort -> contiguous segments); for every LoRA-wrapped linear the
delta ``(x_seg @ A^T) @ B^T * scaling`` is computed per segment with two skinny GEMMs and added in place.  The radix
cache is namespaced by adapter id (``Req.extra_key``) so KV computed under one adapter is never reused under another.
Adapter weights live in a paged pool of ``max_loras_per_batch`` slots on the device; the registry on the host can hold
more (LRU eviction of device slots).

Tensor parallelism: adapters are sharded exactly like the layer they modify.  Column-parallel layers (qkv, gate_up) keep ``A`` whole and take the
rows of ``B`` that belong to this rank's heads / channels (KV heads follow the base layer's replication rule); row-parallel layers (o_proj,
down_proj) take the matching input columns of ``A`` and keep ``B`` whole — their delta is a partial sum that rides the layer's own all-reduce.
"""
from __future__ import annotations

import json
import logging
import os
import re
import zlib
from collections import OrderedDict
from typing import Any, Dict, List, Optional, Tuple

import numpy as np
import torch

logger = logging.getLogger(__name__)

# name of the wrapped module  ->  [(checkpoint sub-module name, output slice index)]
TARGETS = {"qkv_proj": ["q_proj", "k_proj", "v_proj"], "o_proj": ["o_proj"], "gate_up_proj": ["gate_proj", "up_proj"],
           "down_proj": ["down_proj"]}
_CTX: Dict[str, Any] = {"segments": None}


class LoRAAdapter:
    def __init__(self, name: str, rank: int, alpha: float):
        self.name, self.rank, self.scaling = name, rank, alpha / max(1, rank)
        # weights[(layer_idx, sub_module)] = (A [r, in], B [out, r])
        self.weights: Dict[Tuple[int, str], Tuple[torch.Tensor, torch.Tensor]] = {}


def _dummy_adapter(name: str, spec: str, model: Any, device, dtype) -> LoRAAdapter:
    m = re.search(r"rank=(\d+)", spec)
    rank = int(m.group(1)) if m else 8
    ad = LoRAAdapter(name, rank, alpha=float(2 * rank))
    g = torch.Generator(device=device)
    for li, layer in enumerate(model.model.layers):
        mods = {"qkv_proj": layer.self_attn.qkv_proj, "o_proj": layer.self_attn.o_proj}
        if hasattr(layer.mlp, "gate_up_proj"):
            mods.update({"gate_up_proj": layer.mlp.gate_up_proj, "down_proj": layer.mlp.down_proj})
        for mod_name, mod in mods.items():
            subs = TARGETS[mod_name]
            W = mod.tp.world_size if getattr(mod, "tp", None) is not None else 1
            outs = _sub_out_sizes(mod_name, mod)
            for sub, n_out in zip(subs, outs):
                g.manual_seed(zlib.crc32(f"{name}/{li}/{sub}".encode()) & 0x7FFFFFFF)
                # full-size factors (identical on every rank), then this rank's shard: results do not depend on the TP layout
                full_in = mod.in_features * (W if mod_name in ("o_proj", "down_proj") else 1)
                if mod_name in ("o_proj", "down_proj"):
                    full_out = n_out
                elif mod_name == "qkv_proj" and sub != "q_proj":
                    full_out = n_out * max(1, W // getattr(mod, "kv_replicas", 1))
                else:
                    full_out = n_out * W
                A = torch.empty(rank, full_in, dtype=torch.float32, device=device).normal_(0, 0.05, generator=g).to(dtype)
                B = torch.empty(full_out, rank, dtype=torch.float32, device=device).normal_(0, 0.05, generator=g).to(dtype)
                ad.weights[(li, sub)] = shard_adapter_weight(mod_name, sub, mod, A, B)
    return ad


def shard_adapter_weight(mod_name: str, sub: str, mod: Any, A: torch.Tensor, B: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """full (unsharded) LoRA factors of one checkpoint sub-module -> this rank's factors (see module docstring)"""
    tp = getattr(mod, "tp", None)
    if tp is None or tp.world_size == 1:
        return A, B
    r = tp.rank
    if mod_name in ("o_proj", "down_proj"):                      # row parallel: input features sharded
        k = mod.in_features
        return A[:, r * k:(r + 1) * k].contiguous(), B
    if mod_name == "qkv_proj":
        if sub == "q_proj":
            return A, B[r * mod.q_size:(r + 1) * mod.q_size].contiguous()
        src = (r // getattr(mod, "kv_replicas", 1)) * mod.kv_size
        return A, B[src:src + mod.kv_size].contiguous()
    n = list(mod.local_sizes)[0 if sub == "gate_proj" else 1]     # gate_up: both halves sharded alike
    return A, B[r * n:(r + 1) * n].contiguous()


def _sub_out_sizes(mod_name: str, mod: Any) -> List[int]:
    if mod_name == "qkv_proj":
        return [mod.q_size, mod.kv_size, mod.kv_size]
    if mod_name == "gate_up_proj":
        return list(mod.local_sizes)
    return [mod.out_features]


def _load_peft_adapter(name: str, path: str, model: Any, device, dtype) -> LoRAAdapter:
    from safetensors.torch import load_file

    cfg = json.load(open(os.path.join(path, "adapter_config.json")))
    ad = LoRAAdapter(name, int(cfg["r"]), float(cfg.get("lora_alpha", cfg["r"])))
    f = os.path.join(path, "adapter_model.safetensors")
    sd = load_file(f) if os.path.exists(f) else torch.load(os.path.join(path, "adapter_model.bin"), map_location="cpu", weights_only=True)
    pat = re.compile(r"layers\.(\d+)\..*?(q_proj|k_proj|v_proj|o_proj|gate_proj|up_proj|down_proj|qkv_proj|gate_up_proj)\.lora_(A|B)")
    tmp: Dict[Tuple[int, str], Dict[str, torch.Tensor]] = {}
    for k, v in sd.items():
        m = pat.search(k)
        if m:
            tmp.setdefault((int(m.group(1)), m.group(2)), {})[m.group(3)] = v.to(device=device, dtype=dtype)
    layers = model.model.layers
    where = {"q_proj": "qkv_proj", "k_proj": "qkv_proj", "v_proj": "qkv_proj", "o_proj": "o_proj", "gate_proj": "gate_up_proj", "up_proj": "gate_up_proj",
             "down_proj": "down_proj"}
    for (li, sub), ab in tmp.items():
        if not ("A" in ab and "B" in ab) or li >= len(layers):
            continue
        layer = layers[li]
        mods = {"qkv_proj": layer.self_attn.qkv_proj, "o_proj": layer.self_attn.o_proj, "gate_up_proj": getattr(layer.mlp, "gate_up_proj", None),
                "down_proj": getattr(layer.mlp, "down_proj", None)}
        if sub in ("qkv_proj", "gate_up_proj"):          # adapters trained on fused projections (Phi-3 style): split B by the fused layout
            mod = mods[sub]
            if mod is None:
                continue
            W = mod.tp.world_size
            if sub == "qkv_proj":
                sizes = [mod.q_size * W, mod.kv_size * max(1, W // getattr(mod, "kv_replicas", 1))] + [mod.kv_size * max(1, W // getattr(mod, "kv_replicas", 1))]
            else:
                sizes = [n * W for n in mod.local_sizes]
            for part, Bp in zip(TARGETS[sub], ab["B"].split(sizes, 0)):
                ad.weights[(li, part)] = shard_adapter_weight(sub, part, mod, ab["A"], Bp)
            continue
        mod = mods[where[sub]]
        if mod is not None:
            ad.weights[(li, sub)] = shard_adapter_weight(where[sub], sub, mod, ab["A"], ab["B"])
    if not ad.weights:
        raise ValueError(f"no LoRA weights found under {path}")
    return ad


class LoRAManager:
    def __init__(self, runner: Any, max_loras_per_batch: int = 8, max_lora_rank: int = 64):
        self.runner = runner
        self.max_per_batch, self.max_rank = max_loras_per_batch, max_lora_rank
        self.adapters: "OrderedDict[str, LoRAAdapter]" = OrderedDict()
        self._wrap_model()

    # ------------------------------------------------------------------ registry
    def load_adapter(self, name: str, path: str) -> None:
        dev, dt = self.runner.device, self.runner.dtype
        ad = _dummy_adapter(name, path, self.runner.model, dev, dt) if path.startswith("dummy") else \
            _load_peft_adapter(name, path, self.runner.model, dev, dt)
        if ad.rank > self.max_rank:
            raise ValueError(f"adapter rank {ad.rank} exceeds --max-lora-rank {self.max_rank}")
        self.adapters[name] = ad
        logger.info("loaded LoRA adapter %s (rank %d, %d tensors)", name, ad.rank, len(ad.weights))

    def unload_adapter(self, name: str) -> None:
        if self.adapters.pop(name, None) is None:
            raise KeyError(f"adapter {name!r} is not loaded")

    # ------------------------------------------------------------------ per-batch routing
    def prepare_batch(self, sb: Any, fb: Any) -> None:
        names = [r.lora_id for r in sb.reqs]
        if not any(names):
            _CTX["segments"] = None
            return
        used = sorted({n for n in names if n})
        unknown = [n for n in used if n not in self.adapters]
        if unknown:
            raise KeyError(f"LoRA adapter(s) not loaded: {unknown}")
        if len(used) > self.max_per_batch:
            raise RuntimeError(f"{len(used)} adapters in one batch exceeds --max-loras-per-batch {self.max_per_batch}")
        q = np.ones(len(names), dtype=np.int64) if sb.q_lens is None else sb.q_lens.astype(np.int64)
        tok_adapter = np.repeat(np.asarray([used.index(n) if n else -1 for n in names]), q)
        segs = []
        for ai, n in enumerate(used):
            idx = np.flatnonzero(tok_adapter == ai)
            if idx.size:
                segs.append((self.adapters[n], torch.from_numpy(idx).to(self.runner.device)))
        _CTX["segments"] = segs

    # ------------------------------------------------------------------ model wrapping
    def _wrap_model(self) -> None:
        for li, layer in enumerate(self.runner.model.model.layers):
            mods = {"qkv_proj": layer.self_attn.qkv_proj, "o_proj": layer.self_attn.o_proj}
            if hasattr(layer.mlp, "gate_up_proj"):
                mods.update({"gate_up_proj": layer.mlp.gate_up_proj, "down_proj": layer.mlp.down_proj})
            for mod_name, mod in mods.items():
                _install(mod, li, mod_name)


def _install(mod: Any, layer_idx: int, mod_name: str) -> None:
    base_apply = mod.apply
    subs = TARGETS[mod_name]

    def apply_with_lora(x: torch.Tensor, bias: Optional[torch.Tensor] = None) -> torch.Tensor:
        y = base_apply(x, bias)
        segs = _CTX["segments"]
        if not segs:
            return y
        outs = _sub_out_sizes(mod_name, mod)
        x2 = x.reshape(-1, x.shape[-1])
        y2 = y.view(-1, y.shape[-1])
        for ad, idx in segs:
            xs = x2.index_select(0, idx)
            off = 0
            for sub, n_out in zip(subs, outs):
                w = ad.weights.get((layer_idx, sub))
                if w is not None:
                    A, B = w
                    delta = torch.nn.functional.linear(torch.nn.functional.linear(xs, A), B) * ad.scaling
                    y2[:, off:off + n_out].index_add_(0, idx, delta.to(y2.dtype))
                off += n_out
        return y

    mod.apply = apply_with_lora
    mod.fast_path_ok = False     # fused GEMM paths bypass .apply(); keep them off for adapter-wrapped layers
