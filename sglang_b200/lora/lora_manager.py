"""Multi-adapter LoRA serving.

Parity: reference ``python/sglang/srt/lora/`` (``LoRAManager`` ``lora_manager.py:65``: dynamic load/unload, a bounded
number of adapters resident per batch, per-token adapter routing; batched shrink/expand (SGMV) kernels
``kernels/ops/gemm/sgemm_lora_{a,b}.py``).

Design: tokens of a batch are grouped by adapter slot (stable sort -> ``perm``) and cut into tiles of <= 16 tokens; for every LoRA-wrapped
linear two CUDA kernels (``csrc/kernels/lora.cu``: shrink ``h = A x``, expand ``y += B h``) add the delta in place.  The tile table lives
in ONE static device tensor refreshed before each step, so the kernels are part of the captured decode graph (a LoRA batch no longer falls
back to eager).  Adapter factors live in a slot-paged device pool (``LoRAPool``: ``max_loras_per_batch`` slots + the empty slot 0, LRU
residency); the host registry can hold more adapters than slots.  The radix cache is namespaced by adapter id (``Req.extra_key``) so KV
computed under one adapter is never reused under another.

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


class LoRAPool:
    """Slot-paged adapter weights on the device (reference ``lora/mem_pool.py``): for every wrapped linear one pair of tensors
    ``A [slots, n_sub, R, in]`` / ``B [slots, out, R]`` (``R`` = pool rank = max rank rounded up to 8/16/32/64; ranks below it are
    zero-padded, ``alpha / r`` is folded into ``B``).  Slot 0 is the empty adapter.  Residency is LRU over ``slots - 1`` adapters; an
    adapter that a batch needs is copied into a slot before the step (a few MB of device-to-device copies, off the captured graph)."""

    def __init__(self, model: Any, n_slots: int, max_rank: int, device, dtype):
        self.R = 8 if max_rank <= 8 else 16 if max_rank <= 16 else 32 if max_rank <= 32 else 64
        if max_rank > 64:
            raise ValueError("--max-lora-rank above 64 is not supported by the SGMV kernels")
        self.n_slots = n_slots + 1
        self.A: Dict[Tuple[int, str], torch.Tensor] = {}
        self.B: Dict[Tuple[int, str], torch.Tensor] = {}
        self.bounds: Dict[Tuple[int, str], Tuple[int, int]] = {}
        self.mods: Dict[Tuple[int, str], Any] = {}
        for li, layer in enumerate(model.model.layers):
            mods = {"qkv_proj": layer.self_attn.qkv_proj, "o_proj": layer.self_attn.o_proj}
            if hasattr(layer.mlp, "gate_up_proj"):
                mods.update({"gate_up_proj": layer.mlp.gate_up_proj, "down_proj": layer.mlp.down_proj})
            for name, mod in mods.items():
                outs = _sub_out_sizes(name, mod)
                n_sub, N = len(outs), int(sum(outs))
                self.A[(li, name)] = torch.zeros(self.n_slots, n_sub, self.R, mod.in_features, dtype=dtype, device=device)
                self.B[(li, name)] = torch.zeros(self.n_slots, N, self.R, dtype=dtype, device=device)
                b1 = outs[0] if n_sub > 1 else N
                b2 = outs[0] + outs[1] if n_sub > 2 else N
                self.bounds[(li, name)] = (b1, b2)
                self.mods[(li, name)] = mod
        self.slot_of: "OrderedDict[str, int]" = OrderedDict()      # resident adapters, least recently used first
        self.free = list(range(self.n_slots - 1, 0, -1))

    def bytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in list(self.A.values()) + list(self.B.values()))

    def evict(self, name: str) -> None:
        s = self.slot_of.pop(name, None)
        if s is not None:
            self.free.append(s)

    def ensure(self, needed: List[str], adapters: Dict[str, "LoRAAdapter"]) -> Dict[str, int]:
        """make every adapter in ``needed`` resident; returns name -> slot"""
        for n in needed:
            if n in self.slot_of:
                self.slot_of.move_to_end(n)
        for n in needed:
            if n in self.slot_of:
                continue
            if not self.free:
                victim = next((v for v in self.slot_of if v not in needed), None)
                if victim is None:
                    raise RuntimeError(f"{len(needed)} adapters in one batch exceed the {self.n_slots - 1} pool slots (--max-loras-per-batch)")
                self.evict(victim)
            slot = self.free.pop()
            self._write(slot, adapters[n])
            self.slot_of[n] = slot
        return {n: self.slot_of[n] for n in needed}

    def _write(self, slot: int, ad: "LoRAAdapter") -> None:
        for (li, name), A in self.A.items():
            B = self.B[(li, name)]
            A[slot].zero_()
            B[slot].zero_()
            off = 0
            for si, (sub, n_out) in enumerate(zip(TARGETS[name], _sub_out_sizes(name, self.mods[(li, name)]))):
                w = ad.weights.get((li, sub))
                if w is not None:
                    a, b = w
                    r = a.shape[0]
                    A[slot, si, :r].copy_(a)
                    B[slot, off:off + n_out, :r].copy_((b.float() * ad.scaling).to(B.dtype))
                off += n_out


class LoRAManager:
    def __init__(self, runner: Any, max_loras_per_batch: int = 8, max_lora_rank: int = 64):
        self.runner = runner
        self.max_per_batch, self.max_rank = max_loras_per_batch, max_lora_rank
        self.adapters: "OrderedDict[str, LoRAAdapter]" = OrderedDict()
        self.pool = LoRAPool(runner.model, max_loras_per_batch, max_lora_rank, runner.device, runner.dtype)
        # static step metadata (graph input): perm | tile_slot | tile_start | tile_cnt, sized for the largest batch the runner forms
        self.max_tokens = 0
        self.meta = self.meta_host = self.scratch = None
        self.active = False
        a = runner.args
        self._reserve(max(256, int(getattr(a, "cuda_graph_max_bs", 0) or 0), int(getattr(runner, "max_running", 0) or 0),
                          int(getattr(a, "chunked_prefill_size", 0) or 0), int(getattr(a, "max_prefill_tokens", 0) or 0)))
        self._wrap_model()
        logger.info("LoRA pool: %d slots, rank %d, %.1f MB", self.pool.n_slots - 1, self.pool.R, self.pool.bytes() / 2 ** 20)

    def _reserve(self, tokens: int) -> None:
        if tokens <= self.max_tokens:
            return
        if self.max_tokens and getattr(self.runner, "graphs", None):
            raise RuntimeError(f"LoRA step of {tokens} tokens exceeds the metadata reserved before CUDA-graph capture ({self.max_tokens})")
        dev = self.runner.device
        self.max_tokens = tokens
        self.tile_cap = (tokens + 15) // 16 + self.pool.n_slots
        n = tokens + 3 * self.tile_cap
        self.meta = torch.zeros(n, dtype=torch.int32, device=dev)
        self.meta_host = torch.zeros(n, dtype=torch.int32)
        if dev.type == "cuda":
            self.meta_host = self.meta_host.pin_memory()
        self.scratch = torch.zeros(tokens * 3 * self.pool.R, dtype=torch.float32, device=dev)

    def views(self, tokens: Optional[int] = None):
        """(perm, tile_slot, tile_start, tile_cnt) views; ``tokens`` limits the tile grid to what a batch of that size can need"""
        T, C = self.max_tokens, self.tile_cap
        m = self.meta
        c = C if tokens is None else min(C, (tokens + 15) // 16 + self.pool.n_slots)
        return m[:T], m[T:T + c], m[T + C:T + C + c], m[T + 2 * C:T + 2 * C + c]

    # ------------------------------------------------------------------ registry
    def load_adapter(self, name: str, path: str) -> None:
        dev, dt = self.runner.device, self.runner.dtype
        ad = _dummy_adapter(name, path, self.runner.model, dev, dt) if path.startswith("dummy") else \
            _load_peft_adapter(name, path, self.runner.model, dev, dt)
        if ad.rank > self.max_rank:
            raise ValueError(f"adapter rank {ad.rank} exceeds --max-lora-rank {self.max_rank}")
        self.pool.evict(name)                    # a re-load under the same name must not serve stale slot contents
        self.adapters[name] = ad
        logger.info("loaded LoRA adapter %s (rank %d, %d tensors)", name, ad.rank, len(ad.weights))

    def unload_adapter(self, name: str) -> None:
        if self.adapters.pop(name, None) is None:
            raise KeyError(f"adapter {name!r} is not loaded")
        self.pool.evict(name)

    # ------------------------------------------------------------------ per-batch routing
    def prepare_batch(self, sb: Any, fb: Any = None) -> None:
        """group the step's tokens by adapter slot and publish the tile table (host -> the static device tensor, stream-ordered)"""
        names = [r.lora_id for r in sb.reqs]
        used = sorted({n for n in names if n})
        unknown = [n for n in used if n not in self.adapters]
        if unknown:
            raise KeyError(f"LoRA adapter(s) not loaded: {unknown}")
        if len(used) > self.max_per_batch:
            raise RuntimeError(f"{len(used)} adapters in one batch exceeds --max-loras-per-batch {self.max_per_batch}")
        q = np.ones(len(names), dtype=np.int64) if sb.q_lens is None else sb.q_lens.astype(np.int64)
        T = int(q.sum())
        self._reserve(T)
        h = self.meta_host.numpy()
        Tc, C = self.max_tokens, self.tile_cap
        h[Tc + 2 * C:Tc + 3 * C] = 0                              # tile_cnt: everything idle unless filled below
        self.active = bool(used)
        if used:
            slots = self.pool.ensure(used, self.adapters)
            tok_slot = np.repeat(np.asarray([slots[n] if n else 0 for n in names], dtype=np.int32), q)
            order = np.argsort(tok_slot, kind="stable").astype(np.int32)
            sorted_slots = tok_slot[order]
            first = int(np.searchsorted(sorted_slots, 1))          # tokens without an adapter sort first and get no tile
            h[:T] = order
            t = 0
            i = first
            while i < T:
                s = int(sorted_slots[i])
                j = int(np.searchsorted(sorted_slots, s, side="right"))
                while i < j:
                    c = min(16, j - i)
                    h[Tc + t], h[Tc + C + t], h[Tc + 2 * C + t] = s, i, c
                    t += 1
                    i += c
        self.meta.copy_(self.meta_host, non_blocking=True)

    # ------------------------------------------------------------------ model wrapping
    def _wrap_model(self) -> None:
        for (li, name), mod in self.pool.mods.items():
            _install(self, mod, li, name)


def lora_delta_reference(x2: torch.Tensor, y2: torch.Tensor, A: torch.Tensor, B: torch.Tensor, perm, tile_slot, tile_start, tile_cnt,
                         b1: int, b2: int) -> None:
    """plain-PyTorch fp32 reference of kernels/lora.cu (CPU path and the numerics oracle of the GPU test)"""
    n_sub, N = A.shape[1], B.shape[1]
    bounds = [0, b1, b2, N][: n_sub] + [N]
    for s, st, c in zip(tile_slot.tolist(), tile_start.tolist(), tile_cnt.tolist()):
        if c <= 0:
            continue
        idx = perm[st:st + c].long()
        xs = x2.index_select(0, idx).float()
        for si in range(n_sub):
            lo, hi = bounds[si], bounds[si + 1]
            hmid = xs @ A[s, si].float().t()
            y2[idx, lo:hi] = (y2[idx, lo:hi].float() + hmid @ B[s, lo:hi].float().t()).to(y2.dtype)


def _install(mgr: "LoRAManager", mod: Any, layer_idx: int, mod_name: str) -> None:
    base_apply = mod.apply
    key = (layer_idx, mod_name)

    def apply_with_lora(x: torch.Tensor, bias: Optional[torch.Tensor] = None) -> torch.Tensor:
        y = base_apply(x, bias)
        on_gpu = y.is_cuda
        if not mgr.active and not (on_gpu and torch.cuda.is_current_stream_capturing()):
            return y          # no adapter in this step (inside a graph capture the kernels are always recorded: idle tiles cost ~2 us)
        x2 = x.reshape(-1, x.shape[-1])
        y2 = y.view(-1, y.shape[-1])
        perm, ts, tst, tc = mgr.views(x2.shape[0])
        A, B = mgr.pool.A[key], mgr.pool.B[key]
        b1, b2 = mgr.pool.bounds[key]
        from .. import ops

        if on_gpu and ops.load_native() and x2.dtype == torch.bfloat16 and x2.stride(0) % 8 == 0 and x2.stride(1) == 1:
            torch.ops.sglb.lora_sgmv(x2, A, B, ts, tst, tc, perm, mgr.scratch, y2, b1, b2)
        else:
            lora_delta_reference(x2, y2, A, B, perm, ts, tst, tc, b1, b2)
        return y

    mod.apply = apply_with_lora
    mod.fast_path_ok = False     # fused GEMM paths bypass .apply(); keep them off for adapter-wrapped layers
