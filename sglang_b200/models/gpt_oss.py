"""GPT-OSS: GQA with learned per-head attention *sinks*, alternating sliding-window / full layers, YaRN RoPE, and a top-k MoE whose experts
use a clamped SwiGLU variant with biases everywhere.

Parity: reference ``python/sglang/srt/models/gpt_oss.py`` (``GptOssAttention`` sinks + ``sliding_window`` per ``layer_types``, ``GptOssSparseMoeBlock``
router with bias, expert ``gate_up_proj`` stored interleaved, ``swiglu_limit`` clamp, ``alpha = 1.702``) and the MXFP4 expert checkpoint layout
(``*_blocks`` / ``*_scales``; de-quantised at load by ``model_loader/quant_adapters.py``).

Sinks without a new kernel: a sink adds ``exp(s_h)`` to the soft-max denominator of head ``h`` and nothing to the numerator, so with the
attention output ``o`` and its log-sum-exp ``lse`` (which the paged kernels already return) the result is ``o * sigmoid(lse - s_h)`` — one fused
elementwise op after the unmodified tcgen05 / reference attention (also exact under split-KV, which merges by lse first).
"""
from __future__ import annotations

import zlib
from typing import Iterable, List, Tuple

import torch
from torch import nn

from ..configs.model_config import ModelConfig
from ..layers.common import RadixAttention
from ..layers.linear import QKVParallelLinear, ReplicatedLinear, RowParallelLinear
from ..model_executor.forward_batch import ForwardBatch
from ..parallel.parallel_state import get_attn_tp_group, get_tp_group
from .llama import LlamaDecoderLayer, LlamaForCausalLM, LlamaMLP, _layer_id


class GptOssAttention(nn.Module):
    def __init__(self, cfg: ModelConfig, layer_id: int, global_layer: int, dtype, quant, rope, prefix: str):
        super().__init__()
        atp = get_attn_tp_group()
        self.atp = atp
        self.qkv_proj = QKVParallelLinear(cfg.hidden_size, cfg.head_dim, cfg.num_attention_heads, cfg.num_key_value_heads, True, dtype, quant,
                                          prefix + ".qkv_proj", tp_group=atp)
        self.o_proj = RowParallelLinear(cfg.num_attention_heads * cfg.head_dim, cfg.hidden_size, True, True, dtype, quant, prefix + ".o_proj",
                                        tp_group=atp)
        self.q_size, self.kv_size = self.qkv_proj.q_size, self.qkv_proj.kv_size
        kinds = cfg.layer_types or ["sliding_attention" if i % 2 == 0 else "full_attention" for i in range(cfg.num_hidden_layers)]
        window = cfg.sliding_window if (kinds[global_layer] == "sliding_attention" and cfg.sliding_window) else -1
        self.attn = RadixAttention(self.qkv_proj.num_heads, cfg.head_dim, self.qkv_proj.num_kv_heads, layer_id, rope=rope, sliding_window=window)
        self.attn.sinks = nn.Parameter(torch.zeros(self.qkv_proj.num_heads, dtype=torch.float32), requires_grad=False)

    def load_sinks(self, w: torch.Tensor) -> None:
        n = self.attn.sinks.shape[0]
        self.attn.sinks.data.copy_(w.float().narrow(0, self.atp.rank * n, n))

    def forward(self, x: torch.Tensor, fb: ForwardBatch) -> torch.Tensor:
        q, k, v = self.qkv_proj(x).split([self.q_size, self.kv_size, self.kv_size], dim=-1)
        return self.o_proj(self.attn(q.contiguous(), k.contiguous(), v.contiguous(), fb))


class GptOssMoE(nn.Module):
    """router (with bias) -> top-k over the raw logits -> softmax over the selected k; experts sharded over TP along the intermediate dim"""

    def __init__(self, cfg: ModelConfig, layer_id: int, dtype, quant, prefix: str):
        super().__init__()
        tp = get_tp_group()
        self.tp = tp
        E, H, I = cfg.num_local_experts, cfg.hidden_size, cfg.moe_intermediate_size
        assert I % tp.world_size == 0
        self.E, self.k, self.Il = E, cfg.num_experts_per_tok, I // tp.world_size
        self.router = ReplicatedLinear(H, E, True, dtype, None, prefix + ".router")
        mk = lambda *s: nn.Parameter(torch.zeros(*s, dtype=dtype), requires_grad=False)
        self.w_gate, self.w_up, self.w_down = mk(E, self.Il, H), mk(E, self.Il, H), mk(E, H, self.Il)
        self.b_gate, self.b_up, self.b_down = mk(E, self.Il), mk(E, self.Il), mk(E, H)
        self.limit = float(getattr(cfg, "swiglu_limit", None) or 7.0)
        self.alpha = 1.702

    # ---- checkpoint tensors (HF: gate_up_proj [E, H, 2I], gate / up interleaved on the last dim, down_proj [E, I, H]; the MXFP4 adapter
    #      emits the same layout)
    def load(self, leaf: str, w: torch.Tensor) -> bool:
        r, Il = self.tp.rank, self.Il
        H = self.w_gate.shape[2]
        if leaf == "gate_up_proj":
            w = w.transpose(1, 2)                           # [E, H, 2I] -> [E, 2I, H]
            self.w_gate.data.copy_(w[:, 0::2][:, r * Il:(r + 1) * Il])
            self.w_up.data.copy_(w[:, 1::2][:, r * Il:(r + 1) * Il])
        elif leaf == "gate_up_proj_bias":
            self.b_gate.data.copy_(w[:, 0::2][:, r * Il:(r + 1) * Il])
            self.b_up.data.copy_(w[:, 1::2][:, r * Il:(r + 1) * Il])
        elif leaf == "down_proj":
            w = w.transpose(1, 2)                           # [E, I, H] -> [E, H, I]
            self.w_down.data.copy_(w[:, :, r * Il:(r + 1) * Il])
        elif leaf == "down_proj_bias":
            self.b_down.data.copy_(w)
        else:
            return False
        return True

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        T, H = x.shape
        logits = self.router(x).float()
        v, ids = torch.topk(logits, self.k, dim=-1)
        w = torch.softmax(v, dim=-1)
        out = torch.zeros(T, H, dtype=torch.float32, device=x.device)
        flat_ids = ids.reshape(-1)
        order = torch.argsort(flat_ids, stable=True)
        counts = torch.bincount(flat_ids, minlength=self.E).tolist()
        tok = (order // self.k)
        wk = w.reshape(-1)[order]
        s = 0
        for e, n in enumerate(counts):
            if n == 0:
                continue
            rows = tok[s:s + n]
            xe = x[rows]
            g = (xe @ self.w_gate[e].t() + self.b_gate[e]).float().clamp(max=self.limit)
            u = (xe @ self.w_up[e].t() + self.b_up[e]).float().clamp(-self.limit, self.limit)
            h = ((u + 1.0) * (g * torch.sigmoid(g * self.alpha))).to(x.dtype)
            y = (h @ self.w_down[e].t()).float()
            if self.tp.rank == 0:
                y = y + self.b_down[e].float()
            out.index_add_(0, rows, y * wk[s:s + n, None])
            s += n
        out = out.to(x.dtype)
        return self.tp.all_reduce(out) if self.tp.world_size > 1 else out


class GptOssDecoderLayer(LlamaDecoderLayer):
    def __init__(self, cfg: ModelConfig, layer_id: int, dtype, quant, rope, prefix: str, mlp_cls=LlamaMLP):
        super().__init__(cfg, layer_id, dtype, quant, rope, prefix, GptOssMoE)
        self.self_attn = GptOssAttention(cfg, layer_id, int(prefix.rsplit(".", 1)[1]), dtype, quant, rope, prefix + ".self_attn")


class GptOssForCausalLM(LlamaForCausalLM):
    decoder_layer_cls = GptOssDecoderLayer

    def _mlp_cls(self):
        return GptOssMoE

    def load_weights(self, weights: Iterable[Tuple[str, torch.Tensor]]) -> List[str]:
        mods = dict(self.named_modules())
        rest, loaded = [], []
        for name, w in weights:
            lid = _layer_id(name) if ".layers." in name else None
            local = None if lid is None else lid - self.model.start_layer
            if lid is not None and not (self.model.start_layer <= lid < self.model.end_layer):
                continue
            if name.endswith(".self_attn.sinks"):
                mods[f"model.layers.{local}.self_attn"].load_sinks(w)
                loaded.append(name)
            elif ".mlp.experts." in name:
                if mods[f"model.layers.{local}.mlp"].load(name.rsplit(".", 1)[1], w):
                    loaded.append(name)
            else:
                rest.append((name, w))
        return loaded + super().load_weights(rest)

    def dummy_checkpoint(self, cfg: ModelConfig, device, dtype, seed: int = 0, scale: float = 0.02):
        from ..model_loader.loader import dummy_checkpoint_iterator

        def rnd(name, *shape, s=scale):
            g = torch.Generator(device="cpu").manual_seed((seed * 1000003 + zlib.crc32(name.encode())) & 0x7FFFFFFF)
            return (torch.randn(*shape, generator=g) * s).to(dtype)

        dense = dict(cfg.__dict__)
        dense.update(num_local_experts=0, num_experts_per_tok=0, architecture="LlamaForCausalLM", attention_bias=True)
        for name, w in dummy_checkpoint_iterator(ModelConfig(**dense), device, dtype, seed, scale):
            if ".mlp." in name:
                continue
            yield name, w
        E, H, I = cfg.num_local_experts, cfg.hidden_size, cfg.moe_intermediate_size
        for i in range(cfg.num_hidden_layers):
            pre = f"model.layers.{i}."
            yield pre + "self_attn.sinks", rnd(pre + "sinks", cfg.num_attention_heads, s=1.0).float()
            yield pre + "self_attn.o_proj.bias", rnd(pre + "ob", H)
            yield pre + "mlp.router.weight", rnd(pre + "rw", E, H, s=0.5)
            yield pre + "mlp.router.bias", rnd(pre + "rb", E, s=0.1)
            yield pre + "mlp.experts.gate_up_proj", rnd(pre + "gu", E, H, 2 * I, s=0.2)
            yield pre + "mlp.experts.gate_up_proj_bias", rnd(pre + "gub", E, 2 * I, s=0.1)
            yield pre + "mlp.experts.down_proj", rnd(pre + "dn", E, I, H, s=0.1)
            yield pre + "mlp.experts.down_proj_bias", rnd(pre + "dnb", E, H, s=0.05)
