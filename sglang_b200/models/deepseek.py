"""DeepSeek-V2 / V3 family: Multi-head Latent Attention (MLA) + DeepSeekMoE (shared + routed experts, grouped top-k).

Parity: reference ``python/sglang/srt/models/deepseek_v2.py`` (``DeepseekV2AttentionMLA``: q LoRA path, ``kv_a_proj_with_mqa`` ->
compressed latent + decoupled RoPE key, weight absorption ``w_kc`` / ``w_vc`` for decode; ``DeepseekV2MoE``: shared experts +
``TopK`` with ``n_group`` / ``topk_group`` / ``e_score_correction_bias`` / ``routed_scaling_factor``; ``first_k_dense_replace``).

Cache layout (B200 sizing: 576 bf16 = 1,152 B per token per layer instead of 2 x heads x 128): ONE latent "KV head" per token,
K plane = [c_kv (kv_lora_rank) ; k_pe (rope dim)], V plane = [c_kv ; 0].  Attention always runs in the ABSORBED form (MQA over the
latent): q_abs = [q_nope · W_UK ; q_pe], o = (softmax(q_abs · Kᵀ) · c_kv) · W_UVᵀ.

Status: numerics-complete; attention goes through the fp32 torch reference (`ops.torch_ref.paged_attention`, head dim 576) — the
tcgen05 kernel is specialised for D=128 and a latent-MQA variant (576-wide K, 512-wide V, 128 query heads per KV row) is not written
yet, so CUDA graphs are disabled for these models (``supports_cuda_graph = False``).
"""
from __future__ import annotations

import math
from typing import Iterable, List, Tuple

import torch
from torch import nn

from .. import ops
from ..configs.model_config import ModelConfig
from ..layers.common import RMSNorm, RotaryEmbedding
from ..layers.linear import ColumnParallelLinear, MergedColumnParallelLinear, ReplicatedLinear, RowParallelLinear
from ..layers.moe import FusedMoE
from ..model_executor.forward_batch import ForwardBatch
from ..ops import torch_ref
from ..parallel.parallel_state import get_tp_group
from .llama import LlamaDecoderLayer, LlamaForCausalLM, LlamaMLP, _layer_id
from .mixtral import _EP_SIZE


class _LatentAttn:
    """duck-typed stand-in for RadixAttention so the runner sizes the KV pool for the latent cache"""

    def __init__(self, num_heads: int, kv_dim: int, layer_id: int, scale: float):
        self.num_heads, self.num_kv_heads, self.head_dim, self.kv_head_dim = num_heads, 1, kv_dim, kv_dim
        self.latent_v_dim = 0          # set by the attention module: kv_lora_rank (the runner then builds a one-plane latent pool)
        self.layer_id, self.scale, self.window, self.logit_cap, self.rope = layer_id, scale, -1, 0.0, None


class DeepseekV2AttentionMLA(nn.Module):
    def __init__(self, cfg: ModelConfig, layer_id: int, dtype, quant, rope: RotaryEmbedding, prefix: str):
        super().__init__()
        tp = get_tp_group().world_size
        H = cfg.hidden_size
        self.nh = cfg.num_attention_heads // tp
        self.nope, self.rope_dim, self.vdim, self.kv_rank = cfg.qk_nope_head_dim, cfg.qk_rope_head_dim, cfg.v_head_dim, cfg.kv_lora_rank
        qk = self.nope + self.rope_dim
        if cfg.q_lora_rank:
            self.q_a_proj = ReplicatedLinear(H, cfg.q_lora_rank, False, dtype, quant, prefix + ".q_a_proj")
            self.q_a_layernorm = RMSNorm(cfg.q_lora_rank, cfg.rms_norm_eps, dtype)
            self.q_b_proj = ColumnParallelLinear(cfg.q_lora_rank, cfg.num_attention_heads * qk, False, False, dtype, quant, prefix + ".q_b_proj")
        else:
            self.q_proj = ColumnParallelLinear(H, cfg.num_attention_heads * qk, False, False, dtype, quant, prefix + ".q_proj")
        self.q_lora = bool(cfg.q_lora_rank)
        self.kv_a_proj_with_mqa = ReplicatedLinear(H, self.kv_rank + self.rope_dim, False, dtype, quant, prefix + ".kv_a_proj_with_mqa")
        self.kv_a_layernorm = RMSNorm(self.kv_rank, cfg.rms_norm_eps, dtype)
        # kept unquantised: used in absorbed form (einsum over the head-split weight), not as a GEMM on activations
        self.kv_b_proj = ColumnParallelLinear(self.kv_rank, cfg.num_attention_heads * (self.nope + self.vdim), False, False, dtype, None,
                                              prefix + ".kv_b_proj")
        self.o_proj = RowParallelLinear(cfg.num_attention_heads * self.vdim, H, False, True, dtype, quant, prefix + ".o_proj")
        scale = qk ** -0.5
        rs = cfg.rope_scaling or {}
        if rs.get("mscale_all_dim"):
            m = 0.1 * float(rs["mscale_all_dim"]) * math.log(float(rs.get("factor", 1.0))) + 1.0 if float(rs.get("factor", 1.0)) > 1 else 1.0
            scale *= m * m
        self.attn = _LatentAttn(self.nh, self.kv_rank + self.rope_dim, layer_id, scale)
        self.attn.latent_v_dim = self.kv_rank
        self.layer_id = layer_id
        self.cos_sin = torch_ref.build_cos_sin_cache(self.rope_dim, cfg.max_position_embeddings, cfg.rope_theta, None, self.rope_dim)

    def forward(self, x: torch.Tensor, fb: ForwardBatch):
        T = x.shape[0]
        if self.q_lora:
            q = self.q_b_proj(self.q_a_layernorm(self.q_a_proj(x)))
        else:
            q = self.q_proj(x)
        q = q.view(T, self.nh, self.nope + self.rope_dim)
        q_nope, q_pe = q[..., : self.nope], q[..., self.nope:]
        lat = self.kv_a_proj_with_mqa(x)
        c_kv = self.kv_a_layernorm(lat[:, : self.kv_rank].contiguous())
        k_pe = lat[:, self.kv_rank:].unsqueeze(1)
        if self.cos_sin.device != x.device:
            self.cos_sin = self.cos_sin.to(x.device)
        q_pe = torch_ref.apply_rope(q_pe, fb.positions, self.cos_sin, False)
        k_pe = torch_ref.apply_rope(k_pe, fb.positions, self.cos_sin, False)
        w = self.kv_b_proj.weight.view(self.nh, self.nope + self.vdim, self.kv_rank)
        w_kc, w_vc = w[:, : self.nope], w[:, self.nope:]
        q_lat = torch.einsum("thd,hdk->thk", q_nope.float(), w_kc.float()).to(x.dtype)      # absorb W_UK into the query
        q_abs = torch.cat([q_lat, q_pe.to(x.dtype)], dim=-1)                                  # [T, nh, kv_rank + rope]
        pool = fb.kv_pool
        kc, vc = pool.k(self.layer_id), pool.v(self.layer_id)
        loc = fb.out_cache_loc.long()
        kc[loc] = torch.cat([c_kv, k_pe.squeeze(1).to(c_kv.dtype)], dim=-1).unsqueeze(1).to(kc.dtype)
        if getattr(pool, "planes", 2) == 1:       # latent pool: the value IS the first kv_rank columns of the row just written (vc = column view)
            o_lat = torch_ref.paged_attention(q_abs, kc, vc, fb.page_table, fb.req_slots, fb.seq_lens, fb.q_lens, fb.page_size, self.attn.scale,
                                              True, -1, 0.0)
        else:
            vc[loc] = torch.cat([c_kv, torch.zeros(T, self.rope_dim, dtype=c_kv.dtype, device=x.device)], dim=-1).unsqueeze(1).to(vc.dtype)
            o_lat = torch_ref.paged_attention(q_abs, kc, vc, fb.page_table, fb.req_slots, fb.seq_lens, fb.q_lens, fb.page_size, self.attn.scale,
                                              True, -1, 0.0)[..., : self.kv_rank]
        o = torch.einsum("thk,hdk->thd", o_lat.float(), w_vc.float()).to(x.dtype).reshape(T, self.nh * self.vdim)
        return self.o_proj(o, defer_reduce=True)


class DeepseekV2MoE(nn.Module):
    def __init__(self, cfg: ModelConfig, layer_id: int, dtype, quant, prefix: str):
        super().__init__()
        self.cfg = cfg
        E = cfg.num_local_experts
        self.gate = ReplicatedLinear(cfg.hidden_size, E, False, dtype, None, prefix + ".gate")
        self.gate.e_score_correction_bias = nn.Parameter(torch.zeros(E, dtype=torch.float32), requires_grad=False) \
            if cfg.topk_method == "noaux_tc" else None
        self.experts = FusedMoE(E, cfg.num_experts_per_tok, cfg.hidden_size, cfg.moe_intermediate_size, dtype, cfg.norm_topk_prob,
                                _EP_SIZE["v"], quant, layer_id)
        self.shared_experts = None
        if cfg.n_shared_experts:
            I = cfg.moe_intermediate_size * cfg.n_shared_experts
            self.shared_experts = _DenseMLP(cfg.hidden_size, I, dtype, quant, prefix + ".shared_experts")

    def route(self, x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        cfg = self.cfg
        logits = self.gate(x).float()
        scores = torch.sigmoid(logits) if cfg.scoring_func == "sigmoid" else torch.softmax(logits, dim=-1)
        choice = scores if self.gate.e_score_correction_bias is None else scores + self.gate.e_score_correction_bias
        T, E = scores.shape
        if cfg.n_group and cfg.n_group > 1:
            g = choice.view(T, cfg.n_group, E // cfg.n_group)
            gs = g.topk(2, dim=-1).values.sum(-1) if self.gate.e_score_correction_bias is not None else g.max(-1).values
            keep = torch.zeros_like(gs).scatter_(1, gs.topk(cfg.topk_group, dim=-1).indices, 1.0)
            choice = choice.masked_fill(keep.unsqueeze(-1).expand(-1, -1, E // cfg.n_group).reshape(T, E) == 0, float("-inf"))
        ids = choice.topk(cfg.num_experts_per_tok, dim=-1).indices
        w = scores.gather(1, ids)
        if cfg.norm_topk_prob:
            w = w / (w.sum(-1, keepdim=True) + 1e-20)
        return (w * float(cfg.routed_scaling_factor or 1.0)).float().contiguous(), ids.to(torch.int32).contiguous()

    def forward(self, x: torch.Tensor):
        w, ids = self.route(x)
        y = self.experts.forward_routed(x, w, ids)
        if self.shared_experts is None:
            return y
        if hasattr(y, "resolve"):
            y = y.resolve()
        return y + self.shared_experts(x)


class _DenseMLP(nn.Module):
    def __init__(self, H: int, I: int, dtype, quant, prefix: str):
        super().__init__()
        self.gate_up_proj = MergedColumnParallelLinear(H, [I, I], False, dtype, quant, prefix + ".gate_up_proj")
        self.down_proj = RowParallelLinear(I, H, False, True, dtype, quant, prefix + ".down_proj")

    def forward(self, x):
        return self.down_proj(ops.silu_and_mul(self.gate_up_proj(x)))


class DeepseekV2DecoderLayer(LlamaDecoderLayer):
    def __init__(self, cfg: ModelConfig, layer_id: int, dtype, quant, rope, prefix: str, mlp_cls=LlamaMLP):
        gid = int(prefix.rsplit(".", 1)[1])
        sparse = cfg.num_local_experts > 0 and gid >= (cfg.first_k_dense_replace or 0) and gid % max(1, cfg.moe_layer_freq or 1) == 0
        nn.Module.__init__(self)
        self.self_attn = DeepseekV2AttentionMLA(cfg, layer_id, dtype, quant, rope, prefix + ".self_attn")
        self.mlp = DeepseekV2MoE(cfg, layer_id, dtype, quant, prefix + ".mlp") if sparse else LlamaMLP(cfg, dtype, quant, prefix + ".mlp")
        self.input_layernorm = RMSNorm(cfg.hidden_size, cfg.rms_norm_eps, dtype)
        self.post_attention_layernorm = RMSNorm(cfg.hidden_size, cfg.rms_norm_eps, dtype)


class DeepseekV2ForCausalLM(LlamaForCausalLM):
    decoder_layer_cls = DeepseekV2DecoderLayer
    supports_cuda_graph = False      # reference-path attention reads lengths on the host

    @staticmethod
    def dummy_checkpoint(cfg: ModelConfig, device, dtype, seed: int = 0, scale: float = 0.02):
        """full-size random tensors under the HF DeepSeek names (name-seeded, so every TP/EP/PP layout slices the same values)"""
        import zlib

        def rnd(name: str, *shape: int) -> torch.Tensor:
            g = torch.Generator(device=device)
            g.manual_seed((seed * 1000003 + zlib.crc32(name.encode())) & 0x7FFFFFFF)
            return torch.empty(shape, dtype=torch.float32, device=device).normal_(0.0, scale, generator=g).to(dtype)

        ones = lambda n: torch.ones(n, dtype=dtype, device=device)
        H, V, nh = cfg.hidden_size, cfg.vocab_size, cfg.num_attention_heads
        qk = cfg.qk_nope_head_dim + cfg.qk_rope_head_dim
        yield "model.embed_tokens.weight", rnd("embed", V, H)
        for i in range(cfg.num_hidden_layers):
            p = f"model.layers.{i}."
            a = p + "self_attn."
            if cfg.q_lora_rank:
                yield a + "q_a_proj.weight", rnd(a + "q_a", cfg.q_lora_rank, H)
                yield a + "q_a_layernorm.weight", ones(cfg.q_lora_rank)
                yield a + "q_b_proj.weight", rnd(a + "q_b", nh * qk, cfg.q_lora_rank)
            else:
                yield a + "q_proj.weight", rnd(a + "q", nh * qk, H)
            yield a + "kv_a_proj_with_mqa.weight", rnd(a + "kv_a", cfg.kv_lora_rank + cfg.qk_rope_head_dim, H)
            yield a + "kv_a_layernorm.weight", ones(cfg.kv_lora_rank)
            yield a + "kv_b_proj.weight", rnd(a + "kv_b", nh * (cfg.qk_nope_head_dim + cfg.v_head_dim), cfg.kv_lora_rank)
            yield a + "o_proj.weight", rnd(a + "o", H, nh * cfg.v_head_dim)
            yield p + "input_layernorm.weight", ones(H)
            yield p + "post_attention_layernorm.weight", ones(H)
            sparse = cfg.num_local_experts > 0 and i >= (cfg.first_k_dense_replace or 0) and i % max(1, cfg.moe_layer_freq or 1) == 0
            m = p + "mlp."
            if not sparse:
                for nm, shp in (("gate_proj", (cfg.intermediate_size, H)), ("up_proj", (cfg.intermediate_size, H)), ("down_proj", (H, cfg.intermediate_size))):
                    yield m + nm + ".weight", rnd(m + nm, *shp)
                continue
            yield m + "gate.weight", rnd(m + "gate", cfg.num_local_experts, H)
            if cfg.topk_method == "noaux_tc":
                yield m + "gate.e_score_correction_bias", rnd(m + "bias", cfg.num_local_experts).float()
            Im = cfg.moe_intermediate_size
            for e in range(cfg.num_local_experts):
                q = m + f"experts.{e}."
                yield q + "gate_proj.weight", rnd(q + "g", Im, H)
                yield q + "up_proj.weight", rnd(q + "u", Im, H)
                yield q + "down_proj.weight", rnd(q + "d", H, Im)
            if cfg.n_shared_experts:
                Is = Im * cfg.n_shared_experts
                for nm, shp in (("gate_proj", (Is, H)), ("up_proj", (Is, H)), ("down_proj", (H, Is))):
                    yield m + "shared_experts." + nm + ".weight", rnd(m + "s" + nm, *shp)
        yield "model.norm.weight", ones(H)
        if not cfg.tie_word_embeddings:
            yield "lm_head.weight", rnd("lm_head", V, H)
    stacked_params = [(".gate_up_proj", ".gate_proj", 0), (".gate_up_proj", ".up_proj", 1)]

    def load_weights(self, weights: Iterable[Tuple[str, torch.Tensor]]) -> List[str]:
        loaded, rest = load_moe_tensors(self, weights)
        return loaded + super().load_weights(rest)


def load_moe_tensors(self, weights):
    """route the DeepSeek-style expert / router tensors into ``FusedMoE`` and the gate; returns (loaded names, remaining (name, tensor) pairs)"""
    if True:
        rest, loaded = [], []
        modules = dict(self.named_modules())
        shard_of = {"gate_proj": "w1", "up_proj": "w3", "down_proj": "w2"}
        for name, w in weights:
            if ".mlp.experts." in name:
                lid = _layer_id(name)
                if not (self.model.start_layer <= lid < self.model.end_layer):
                    continue
                mod = modules.get(f"model.layers.{lid - self.model.start_layer}.mlp.experts")
                if mod is None:
                    continue
                eid = int(name.split(".experts.")[1].split(".")[0])
                shard = shard_of[name.split(".")[-2]]
                mod.weight_loader(mod.w2 if shard == "w2" else mod.w13, w, eid, shard)
                loaded.append(name)
            elif name.endswith("mlp.gate.e_score_correction_bias"):
                lid = _layer_id(name)
                m = modules.get(f"model.layers.{lid - self.model.start_layer}.mlp.gate")
                if m is not None and m.e_score_correction_bias is not None:
                    m.e_score_correction_bias.data.copy_(w.float())
                    loaded.append(name)
            else:
                rest.append((name, w))
        return loaded, rest


class DeepseekV3ForCausalLM(DeepseekV2ForCausalLM):
    pass
