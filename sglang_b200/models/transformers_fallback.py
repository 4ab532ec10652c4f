"""Generic fallback: serve any decoder-only Hugging Face architecture that this package has no native class for.

Parity: reference ``python/sglang/srt/models/transformers.py`` (``TransformersForCausalLM``: the HF module tree with the engine's attention
plugged in through ``ALL_ATTENTION_FUNCTIONS``) and ``--model-impl {auto,sglang,transformers}``.

The Hugging Face modules do the projections / norms / MLP / rotary embedding; attention itself is ours: the registered attention function
receives the already-rotated ``[1, H, T, D]`` query/key/value of one layer, writes K/V into the paged pool at this step's slots and calls
``ops.paged_attention`` (the same tcgen05 prefill / decode kernels as every native model), so radix-cache reuse, chunked prefill, continuous
batching and the sampler all work unchanged.  What it does not get: tensor parallelism (weights stay whole on one GPU), quantised linears,
CUDA graphs, and recurrent-state layers (Mamba / conv hybrids need the native classes).
"""
from __future__ import annotations

import logging
import threading
from typing import Iterable, List, Optional, Tuple

import torch
from torch import nn

from .. import ops
from ..configs.model_config import ModelConfig
from ..model_executor.forward_batch import ForwardBatch

logger = logging.getLogger(__name__)
ATTN_NAME = "sglang_b200"
_CTX = threading.local()
_DT = {"bfloat16": torch.bfloat16, "float16": torch.float16, "float32": torch.float32}


def _paged_attention_forward(module, query, key, value, attention_mask=None, dropout: float = 0.0, scaling: Optional[float] = None,
                             sliding_window: Optional[int] = None, softcap: Optional[float] = None, **kwargs):
    """the ``AttentionInterface`` hook: query [1, Hq, T, D], key/value [1, Hkv, T, D] (rotary already applied) -> ([1, T, Hq, D], None)"""
    fb: ForwardBatch = _CTX.fb
    if kwargs.get("s_aux") is not None:
        raise NotImplementedError("attention sinks are not supported by the Transformers fallback (use the native GPT-OSS class)")
    if not getattr(module, "is_causal", True):
        raise NotImplementedError("bidirectional attention layers are not supported by the Transformers fallback")
    lid = int(getattr(module, "layer_idx", 0) or 0)
    _, Hq, T, D = query.shape
    Hkv = key.shape[1]
    pool = fb.kv_pool
    kc, vc = pool.k(lid), pool.v(lid)
    if kc.shape[-2] != Hkv or kc.shape[-1] != D or value.shape[-1] != D:
        raise NotImplementedError(f"layer {lid}: K/V geometry [{Hkv}, {D}]/[{value.shape[1]}, {value.shape[-1]}] differs from the pool's "
                                  f"[{kc.shape[-2]}, {kc.shape[-1]}] (per-layer head geometry needs a native model class)")
    q3 = query[0].transpose(0, 1).contiguous()
    loc = fb.out_cache_loc.long()
    kc[loc] = key[0].transpose(0, 1).to(kc.dtype)
    vc[loc] = value[0].transpose(0, 1).to(vc.dtype)
    o = ops.paged_attention(q3, kc, vc, fb.page_table, fb.req_slots, fb.seq_lens, fb.cu_q, fb.q_lens, fb.max_q_len, fb.page_size,
                            float(scaling) if scaling is not None else D ** -0.5, causal=True,
                            window=int(sliding_window) if sliding_window else -1, logit_cap=float(softcap or 0.0),
                            k_scale=pool.k_scale, v_scale=pool.v_scale, num_splits=fb.num_splits, workspace=fb.attn_workspace)
    return o.view(1, T, Hq, D), None


def register_attention() -> None:
    from transformers import AttentionInterface

    AttentionInterface.register(ATTN_NAME, _paged_attention_forward)
    try:        # no mask tensor is ever needed: causality / windows come from the paged kernel's own length arithmetic
        from transformers.masking_utils import AttentionMaskInterface

        AttentionMaskInterface.register(ATTN_NAME, lambda *a, **k: None)
    except Exception:      # older transformers: custom attention names skip mask creation on their own
        pass


def _set_attn_impl(hf_cfg) -> None:
    hf_cfg._attn_implementation = ATTN_NAME
    for sub in getattr(hf_cfg, "sub_configs", {}) or {}:
        c = getattr(hf_cfg, sub, None)
        if c is not None:
            c._attn_implementation = ATTN_NAME


class _DummyInit:
    def __init__(self, seed: int, scale: float):
        self.seed, self.scale = seed, scale


class TransformersForCausalLM(nn.Module):
    supports_cuda_graph = False          # the HF module tree allocates and branches on the host

    def __init__(self, cfg: ModelConfig, quant_name: Optional[str] = None, device="cpu"):
        super().__init__()
        import transformers as tf

        from ..parallel.parallel_state import get_pp_group, get_tp_group

        if get_tp_group().world_size > 1 or get_pp_group().world_size > 1:
            raise NotImplementedError("the Transformers fallback runs on one GPU (tp_size = pp_size = 1); "
                                      f"{(cfg.hf_fallback or {}).get('arch')} has no native tensor-parallel class here")
        if quant_name:
            raise NotImplementedError("quantisation is not available through the Transformers fallback")
        register_attention()
        self.cfg = cfg
        self.path = cfg.hf_fallback["path"]
        self.hf_cfg = tf.AutoConfig.from_pretrained(self.path)
        _set_attn_impl(self.hf_cfg)
        self.hf = tf.AutoModelForCausalLM.from_config(self.hf_cfg, dtype=_DT[cfg.dtype])
        for p in self.hf.parameters():
            p.requires_grad_(False)
        # the runner sizes the KV pool and picks attention splits from ``model.model.layers[i].self_attn.attn``: describe the geometry there
        from types import SimpleNamespace as NS

        L = cfg.num_hidden_layers
        self.model = NS(start_layer=0, end_layer=L, layers=[NS(self_attn=NS(attn=NS(num_kv_heads=cfg.num_key_value_heads, window=-1,
                                                                                     kv_head_dim=None, layer_id=i))) for i in range(L)])
        logger.warning("%s has no native implementation: serving it through the Transformers fallback (single GPU, eager)",
                       cfg.hf_fallback.get("arch"))

    # ---- forward
    def forward(self, fb: ForwardBatch, inputs_embeds: Optional[torch.Tensor] = None) -> torch.Tensor:
        if inputs_embeds is None:
            ids = fb.next_token_buf[fb.req_slots.long()] if fb.input_ids is None else fb.input_ids
            kw = {"input_ids": ids.long().unsqueeze(0)}
        else:
            kw = {"inputs_embeds": inputs_embeds.unsqueeze(0)}
        _CTX.fb = fb
        try:
            out = self.hf.base_model(position_ids=fb.positions.long().unsqueeze(0), attention_mask=None, use_cache=False, **kw)
        finally:
            _CTX.fb = None
        return out[0][0]

    def compute_logits(self, hidden: torch.Tensor, key: str = "logits") -> torch.Tensor:
        y = self.hf.get_output_embeddings()(hidden)[:, : self.cfg.vocab_size]
        if self.cfg.logits_scaling:
            y = y / self.cfg.logits_scaling
        if self.cfg.final_logit_softcapping > 0:
            c = self.cfg.final_logit_softcapping
            y = torch.tanh(y.float() / c) * c
        return y

    # ---- weights: Hugging Face's own loader (it knows each family's checkpoint-name conversions)
    @staticmethod
    def dummy_checkpoint(cfg: ModelConfig, device, dtype, seed: int = 0, scale: float = 0.02):
        return _DummyInit(seed, scale)

    def load_weights(self, weights: Iterable[Tuple[str, torch.Tensor]]) -> List[str]:
        import transformers as tf

        if isinstance(weights, _DummyInit):          # --load-format dummy: name-seeded random matrices, unit norms, zero biases
            import zlib

            for name, p in self.hf.named_parameters():
                if p.dim() > 1:
                    g = torch.Generator(device=p.device)
                    g.manual_seed((weights.seed * 1000003 + zlib.crc32(name.encode())) & 0x7FFFFFFF)
                    p.copy_(torch.empty(p.shape, dtype=torch.float32, device=p.device).normal_(0.0, weights.scale, generator=g).to(p.dtype))
                elif name.endswith("bias"):
                    p.zero_()
                elif "norm" in name:
                    p.fill_(1.0)
            return [n for n, _ in self.hf.named_parameters()]
        dev = next(self.hf.parameters()).device
        src = tf.AutoModelForCausalLM.from_pretrained(self.path, config=self.hf_cfg, dtype=_DT[self.cfg.dtype])
        self.hf.load_state_dict(src.state_dict(), strict=True)
        self.hf.to(dev)
        return list(src.state_dict().keys())
