"""Weight loading.

Parity: reference ``python/sglang/srt/model_loader/loader.py`` (``DefaultModelLoader:357`` safetensors / pt,
``DummyModelLoader:1602`` random init of the real architecture, ``ShardedStateLoader``) and ``weight_utils.py``.

``dummy``: every parameter is filled *directly on the target device* from a per-parameter seeded generator — no
host staging, so a 70B random init costs seconds, and all TP ranks agree on replicated tensors.
"""
from __future__ import annotations

import glob
import os
import zlib
from typing import Iterator, Optional, Tuple

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


def dummy_checkpoint_iterator(cfg: ModelConfig, device: torch.device, dtype: torch.dtype, seed: int = 0,
                              scale: float = 0.02) -> Iterator[Tuple[str, torch.Tensor]]:
    """A synthetic *HF-named, unsharded* checkpoint generated tensor by tensor on the device from name-hashed seeds.
    Feeding it through ``model.load_weights`` makes dummy weights independent of the TP/PP/EP layout (every rank slices
    the same global tensor), which is what lets TP=N runs be compared with TP=1 and exercises the real loader path."""
    def rnd(name: str, *shape: int) -> torch.Tensor:
        g = torch.Generator(device=device)
        g.manual_seed((seed * 1000003 + zlib.crc32(name.encode())) & 0x7FFFFFFF)
        return torch.empty(shape, dtype=torch.float32, device=device).normal_(0.0, scale, generator=g).to(dtype)

    H, V, D = cfg.hidden_size, cfg.vocab_size, cfg.head_dim
    Hq, Hkv, I = cfg.num_attention_heads, cfg.num_key_value_heads, cfg.intermediate_size
    gemma = "Gemma" in cfg.architecture
    norm = (lambda n: torch.zeros(n, dtype=dtype, device=device)) if gemma else (lambda n: torch.ones(n, dtype=dtype, device=device))
    yield "model.embed_tokens.weight", rnd("model.embed_tokens.weight", V, H)
    for i in range(cfg.num_hidden_layers):
        p = f"model.layers.{i}."
        for nm, n_out in (("q_proj", Hq * D), ("k_proj", Hkv * D), ("v_proj", Hkv * D)):
            yield p + f"self_attn.{nm}.weight", rnd(p + nm, n_out, H)
            if cfg.attention_bias:
                yield p + f"self_attn.{nm}.bias", rnd(p + nm + ".bias", n_out)
        yield p + "self_attn.o_proj.weight", rnd(p + "o_proj", H, Hq * D)
        if cfg.qk_norm:
            yield p + "self_attn.q_norm.weight", norm(D)
            yield p + "self_attn.k_norm.weight", norm(D)
        if cfg.is_moe:
            yield p + "block_sparse_moe.gate.weight", rnd(p + "gate", cfg.num_local_experts, H)
            Im = cfg.moe_intermediate_size
            for e in range(cfg.num_local_experts):
                q = p + f"block_sparse_moe.experts.{e}."
                yield q + "w1.weight", rnd(q + "w1", Im, H)
                yield q + "w3.weight", rnd(q + "w3", Im, H)
                yield q + "w2.weight", rnd(q + "w2", H, Im)
        else:
            yield p + "mlp.gate_proj.weight", rnd(p + "gate_proj", I, H)
            yield p + "mlp.up_proj.weight", rnd(p + "up_proj", I, H)
            yield p + "mlp.down_proj.weight", rnd(p + "down_proj", H, I)
        yield p + "input_layernorm.weight", norm(H)
        yield p + "post_attention_layernorm.weight", norm(H)
    yield "model.norm.weight", norm(H)
    if not cfg.tie_word_embeddings:
        yield "lm_head.weight", rnd("lm_head.weight", V, H)


@torch.no_grad()
def init_dummy_weights(model: nn.Module, seed: int = 0, scale: float = 0.02, cfg: Optional[ModelConfig] = None) -> None:
    """Random-init weights of the real architecture (reference ``DummyModelLoader``).  The language model is filled
    through the layout-independent synthetic checkpoint above; anything it does not cover (vision tower, heads) gets a
    per-parameter name-seeded init."""
    params = dict(model.named_parameters())
    if cfg is not None and hasattr(model, "load_weights"):
        for p in params.values():
            if p.numel():
                p.data.view(-1)[:1].fill_(float("nan"))
        first = next(iter(params.values()))
        # families whose tensors are not Llama-shaped describe their own full-size synthetic checkpoint (same loader path, so the
        # result is independent of the TP/EP/PP layout)
        it = model.dummy_checkpoint(cfg, first.device, first.dtype, seed, scale) if hasattr(model, "dummy_checkpoint") else \
            dummy_checkpoint_iterator(cfg, first.device, first.dtype, seed, scale)
        model.load_weights(it)
        todo = {n: p for n, p in params.items() if p.numel() and bool(torch.isnan(p.data.view(-1)[0].float()))}
    else:
        todo = params
    for name, p in todo.items():
        if p.dim() <= 1:
            if name.endswith("bias"):
                p.zero_()
            elif "norm" in name:
                gemma = getattr(dict(model.named_modules()).get(name.rsplit(".", 1)[0]), "gemma", False)
                p.fill_(0.0 if gemma else 1.0)
            else:
                g1 = torch.Generator(device=p.device)
                g1.manual_seed((seed * 1000003 + zlib.crc32(name.encode())) & 0x7FFFFFFF)
                p.normal_(0.0, scale, generator=g1)
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
               seed: int = 0, remote: Optional[dict] = None) -> nn.Module:
    model = build_model(cfg, device, quant)
    if load_format == "remote_instance":
        # clone the weights of a RUNNING instance with the same parallel layout (reference RemoteInstanceModelLoader): no disk, no re-sharding
        from ..model_executor.control import receive_weights_from_seed
        from ..parallel.parallel_state import get_tp_group

        process_weights_after_loading(model)          # bring the parameters into their runtime format first; the seed sends that format
        receive_weights_from_seed(model, remote["seed_url"], get_tp_group().rank, device.type == "cuda", int(remote["port"]),
                                  remote.get("address", "127.0.0.1"))
        return model
    if load_format == "dummy":
        init_dummy_weights(model, seed, cfg=cfg)
    elif load_format == "gguf" or (load_format == "auto" and model_path.endswith(".gguf") and os.path.isfile(model_path)):
        from .gguf import iter_gguf

        model.load_weights(iter_gguf(model_path, next(model.parameters()).dtype))
    elif load_format in ("safetensors", "auto", "pt"):
        it = iter_pt(model_path) if load_format == "pt" else iter_safetensors(model_path)
        from .quant_adapters import adapt_quantized_checkpoint, read_quant_config

        qcfg = read_quant_config(model_path)
        if qcfg:      # AWQ / GPTQ / compressed-tensors / fp8 / ModelOpt / mxfp4 / bitsandbytes checkpoint: de-quantise into the serving dtype
            it = adapt_quantized_checkpoint(it, qcfg, next(model.parameters()).dtype, keep_block_fp8=quant in ("fp8_block", "mxfp8"))
        model.load_weights(it)
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
