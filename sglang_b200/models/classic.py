"""Pre-LayerNorm decoders with plain (non-gated) MLPs: GPT-2, OPT, GPT-NeoX / Pythia, StarCoder2.

Parity: reference ``python/sglang/srt/models/{gpt2,opt,gpt_neox,starcoder2}.py``.  One implementation, four checkpoint dialects:

* LayerNorm with bias (``F.layer_norm``), biased linears, ``fc1 -> act -> fc2``;
* positions: learned absolute table (GPT-2 ``wpe``, OPT ``embed_positions`` with its +2 offset) or RoPE (NeoX partial rotary, StarCoder2);
* residual wiring: sequential, or NeoX's parallel residual ``x + attn(ln1(x)) + mlp(ln2(x))``;
* attention / paged KV / TP sharding / LM head are the shared building blocks (``RadixAttention`` etc.), so prefix caching, chunked
  prefill, CUDA graphs and tensor parallelism work unchanged.

Internal (canonical) parameter names are Llama-like; each family only maps its checkpoint names (and un-fuses / transposes what HF stores
fused or as ``Conv1D``).
"""
from __future__ import annotations

import zlib
from typing import Iterable, List, Optional, Tuple

import torch
from torch import nn

from .. import ops
from ..configs.model_config import ModelConfig
from ..layers.common import RadixAttention, RotaryEmbedding
from ..layers.linear import ColumnParallelLinear, ParallelLMHead, QKVParallelLinear, RowParallelLinear, VocabParallelEmbedding
from ..model_executor.forward_batch import ForwardBatch
from ..parallel.parallel_state import get_attn_tp_group, get_pp_group, get_tp_group
from .llama import _DT, _load_stacked_bias

_ACT = {"gelu": lambda x: torch.nn.functional.gelu(x), "gelu_new": lambda x: torch.nn.functional.gelu(x, approximate="tanh"),
        "gelu_pytorch_tanh": lambda x: torch.nn.functional.gelu(x, approximate="tanh"), "gelu_fast": lambda x: torch.nn.functional.gelu(x, approximate="tanh"),
        "relu": torch.relu, "silu": torch.nn.functional.silu, "relu2": lambda x: torch.relu(x).square()}


class LayerNorm(nn.Module):
    def __init__(self, n: int, eps: float, dtype):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(n, dtype=dtype), requires_grad=False)
        self.bias = nn.Parameter(torch.zeros(n, dtype=dtype), requires_grad=False)
        self.eps = eps

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return torch.nn.functional.layer_norm(x.float(), (x.shape[-1],), self.weight.float(), self.bias.float(), self.eps).to(x.dtype)


class ClassicMLP(nn.Module):
    def __init__(self, cfg: ModelConfig, dtype, quant, prefix: str):
        super().__init__()
        self.fc1 = ColumnParallelLinear(cfg.hidden_size, cfg.intermediate_size, True, False, dtype, quant, prefix + ".fc1")
        self.fc2 = RowParallelLinear(cfg.intermediate_size, cfg.hidden_size, True, True, dtype, quant, prefix + ".fc2")
        self.act = _ACT[cfg.hidden_act]

    def forward(self, x):
        return self.fc2(self.act(self.fc1(x)))


class ClassicAttention(nn.Module):
    def __init__(self, cfg: ModelConfig, layer_id: int, dtype, quant, rope: Optional[RotaryEmbedding], prefix: str):
        super().__init__()
        atp = get_attn_tp_group()
        self.qkv_proj = QKVParallelLinear(cfg.hidden_size, cfg.head_dim, cfg.num_attention_heads, cfg.num_key_value_heads, True, dtype, quant,
                                          prefix + ".qkv_proj", tp_group=atp)
        self.o_proj = RowParallelLinear(cfg.num_attention_heads * cfg.head_dim, cfg.hidden_size, True, True, dtype, quant, prefix + ".o_proj",
                                        tp_group=atp)
        self.q_size, self.kv_size = self.qkv_proj.q_size, self.qkv_proj.kv_size
        window = cfg.sliding_window if (cfg.sliding_window and cfg.sliding_window > 0) else -1
        self.attn = RadixAttention(self.qkv_proj.num_heads, cfg.head_dim, self.qkv_proj.num_kv_heads, layer_id, rope=rope, sliding_window=window)
        self.q_layernorm = self.k_layernorm = None
        if cfg.qk_norm:          # Persimmon: LayerNorm (weight + bias) over each head's q / k before the rotary embedding
            self.q_layernorm = LayerNorm(cfg.head_dim, cfg.rms_norm_eps, dtype)
            self.k_layernorm = LayerNorm(cfg.head_dim, cfg.rms_norm_eps, dtype)

    def forward(self, x, fb: ForwardBatch):
        q, k, v = self.qkv_proj(x).split([self.q_size, self.kv_size, self.kv_size], dim=-1)
        if self.q_layernorm is not None:
            T, D = q.shape[0], self.attn.head_dim
            q = self.q_layernorm(q.reshape(T * self.attn.num_heads, D)).view(T, self.q_size)
            k = self.k_layernorm(k.reshape(T * self.attn.num_kv_heads, D)).view(T, self.kv_size)
        return self.o_proj(self.attn(q.contiguous(), k.contiguous(), v.contiguous(), fb))


class ClassicLayer(nn.Module):
    def __init__(self, cfg: ModelConfig, layer_id: int, dtype, quant, rope, prefix: str):
        super().__init__()
        self.input_layernorm = LayerNorm(cfg.hidden_size, cfg.rms_norm_eps, dtype)
        self.post_attention_layernorm = LayerNorm(cfg.hidden_size, cfg.rms_norm_eps, dtype)
        self.self_attn = ClassicAttention(cfg, layer_id, dtype, quant, rope, prefix + ".self_attn")
        self.mlp = ClassicMLP(cfg, dtype, quant, prefix + ".mlp")
        self.parallel = bool(cfg.parallel_residual)
        self.shared_ln = cfg.norm_type == "shared_layernorm"       # Phi / GPT-J: ONE LayerNorm feeds both branches of the parallel block

    def forward(self, x, fb: ForwardBatch):
        n = self.input_layernorm(x)
        a = self.self_attn(n, fb)
        if self.parallel:
            return x + a + self.mlp(n if self.shared_ln else self.post_attention_layernorm(x))
        x = x + a
        return x + self.mlp(self.post_attention_layernorm(x))


class ClassicModel(nn.Module):
    def __init__(self, cfg: ModelConfig, quant_name: Optional[str], device):
        super().__init__()
        from ..layers.quantization import get_quant_method

        self.cfg = cfg
        dtype = _DT[cfg.dtype]
        pp = get_pp_group()
        self.pp = pp
        assert pp.world_size == 1, "classic decoders: pipeline parallelism is not wired (residual stream is a single tensor)"
        self.start_layer, self.end_layer = 0, cfg.num_hidden_layers
        self.embed_tokens = VocabParallelEmbedding(cfg.vocab_size, cfg.hidden_size, dtype)
        self.embed_positions = None
        rope = None
        if cfg.pos_emb == "learned":
            self.embed_positions = nn.Embedding(cfg.max_position_embeddings + cfg.pos_offset, cfg.hidden_size, dtype=dtype)
            self.embed_positions.weight.requires_grad_(False)
        else:
            rope = RotaryEmbedding(cfg.head_dim, cfg.max_position_embeddings, cfg.rope_theta, cfg.rope_scaling, cfg.rope_is_neox,
                                   cfg.partial_rotary_factor, device)
        self.layers = nn.ModuleList([ClassicLayer(cfg, i, dtype, get_quant_method(quant_name), rope, f"model.layers.{i}")
                                     for i in range(cfg.num_hidden_layers)])
        self.norm = LayerNorm(cfg.hidden_size, cfg.rms_norm_eps, dtype)

    def forward(self, fb: ForwardBatch, inputs_embeds: Optional[torch.Tensor] = None) -> torch.Tensor:
        if inputs_embeds is not None:
            h = inputs_embeds
        elif fb.input_ids is None:
            h = self.embed_tokens(fb.next_token_buf, fb.req_slots)
        else:
            h = self.embed_tokens(fb.input_ids)
        if self.embed_positions is not None:
            h = h + self.embed_positions((fb.positions.long() + self.cfg.pos_offset))
        for layer in self.layers:
            h = layer(h, fb)
        return self.norm(h)


class ClassicForCausalLM(nn.Module):
    """canonical names: model.embed_tokens / model.embed_positions / model.layers.N.{input_layernorm,post_attention_layernorm,
    self_attn.{q,k,v,o}_proj, mlp.{fc1,fc2}} / model.norm / lm_head"""
    stacked_params = [(".qkv_proj", ".q_proj", "q"), (".qkv_proj", ".k_proj", "k"), (".qkv_proj", ".v_proj", "v")]

    def __init__(self, cfg: ModelConfig, quant_name: Optional[str] = None, device="cpu"):
        super().__init__()
        self.cfg = cfg
        self.model = ClassicModel(cfg, quant_name, device)
        self.lm_head = None if cfg.tie_word_embeddings else ParallelLMHead(cfg.vocab_size, cfg.hidden_size, _DT[cfg.dtype])
        self.register_buffer("lm_head_bias", torch.zeros(cfg.vocab_size, dtype=_DT[cfg.dtype]), persistent=False)      # Phi / GPT-J (zero otherwise)

    def forward(self, fb: ForwardBatch, inputs_embeds=None):
        return self.model(fb, inputs_embeds)

    def compute_logits(self, hidden: torch.Tensor, key: str = "logits") -> torch.Tensor:
        head = self.lm_head if self.lm_head is not None else self.model.embed_tokens
        y = ops.linear(hidden, head.weight)
        tp = get_tp_group()
        if tp.world_size > 1:
            y = tp.all_gather(y, dim=-1)
        return y[:, : self.cfg.vocab_size] + self.lm_head_bias

    # ---- checkpoint dialects
    def remap(self, name: str, w: torch.Tensor) -> Iterable[Tuple[str, torch.Tensor]]:
        yield name, w

    def load_weights(self, weights: Iterable[Tuple[str, torch.Tensor]]) -> List[str]:
        params = dict(self.named_parameters())
        modules = dict(self.named_modules())
        loaded: List[str] = []
        for raw_name, raw_w in weights:
            for name, w in self.remap(raw_name, raw_w):
                if name == "lm_head.bias":
                    self.lm_head_bias.copy_(w[: self.lm_head_bias.shape[0]])
                    loaded.append(name)
                    continue
                if name.startswith("lm_head") and self.lm_head is None:
                    continue
                target, shard = name, None
                for merged, orig, sid in self.stacked_params:
                    if orig in name:
                        target, shard = name.replace(orig, merged), sid
                        break
                if target not in params:
                    continue
                p = params[target]
                mod = modules[target.rsplit(".", 1)[0]]
                if hasattr(mod, "weight_loader"):
                    if target.endswith(".bias") and shard is not None:
                        _load_stacked_bias(mod, p, w, shard)
                    elif target.endswith(".bias") and p.dim() == 1 and not isinstance(mod, RowParallelLinear):
                        n = p.shape[0]
                        p.data.copy_(w.narrow(0, get_tp_group().rank * n, n) if w.shape[0] != n else w)
                    else:
                        mod.weight_loader(p, w, shard) if shard is not None else mod.weight_loader(p, w)
                else:
                    p.data.copy_(w[: p.shape[0]] if w.shape != p.shape and w.dim() == p.dim() and w.shape[1:] == p.shape[1:] else w)
                loaded.append(name)
        return loaded

    @staticmethod
    def dummy_checkpoint(cfg: ModelConfig, device, dtype, seed: int = 0, scale: float = 0.02):
        def rnd(name: str, *shape: int) -> torch.Tensor:
            g = torch.Generator(device=device)
            g.manual_seed((seed * 1000003 + zlib.crc32(name.encode())) & 0x7FFFFFFF)
            return torch.empty(shape, dtype=torch.float32, device=device).normal_(0.0, scale, generator=g).to(dtype)

        H, I, V, D = cfg.hidden_size, cfg.intermediate_size, cfg.vocab_size, cfg.head_dim
        Hq, Hkv = cfg.num_attention_heads, cfg.num_key_value_heads
        yield "model.embed_tokens.weight", rnd("emb", V, H)
        if cfg.pos_emb == "learned":
            yield "model.embed_positions.weight", rnd("pos", cfg.max_position_embeddings + cfg.pos_offset, H)
        for i in range(cfg.num_hidden_layers):
            p = f"model.layers.{i}."
            for nm, n_out, n_in in (("self_attn.q_proj", Hq * D, H), ("self_attn.k_proj", Hkv * D, H), ("self_attn.v_proj", Hkv * D, H),
                                    ("self_attn.o_proj", H, Hq * D), ("mlp.fc1", I, H), ("mlp.fc2", H, I)):
                yield p + nm + ".weight", rnd(p + nm, n_out, n_in)
                yield p + nm + ".bias", rnd(p + nm + "b", n_out)
            for nm in ("input_layernorm", "post_attention_layernorm"):
                yield p + nm + ".weight", torch.ones(H, dtype=dtype, device=device)
                yield p + nm + ".bias", rnd(p + nm + "b", H)
        yield "model.norm.weight", torch.ones(H, dtype=dtype, device=device)
        yield "model.norm.bias", rnd("normb", H)
        if not cfg.tie_word_embeddings:
            yield "lm_head.weight", rnd("lm_head", V, H)


class Starcoder2ForCausalLM(ClassicForCausalLM):
    def remap(self, name, w):
        yield name.replace(".mlp.c_fc.", ".mlp.fc1.").replace(".mlp.c_proj.", ".mlp.fc2."), w


class GPT2LMHeadModel(ClassicForCausalLM):
    """HF stores Conv1D weights ([in, out]) and a fused c_attn"""

    def remap(self, name, w):
        H = self.cfg.hidden_size
        if name.startswith("model."):         # already canonical (synthetic checkpoint)
            yield name, w
            return
        name = name[len("transformer."):] if name.startswith("transformer.") else name
        if name in ("wte.weight",):
            yield "model.embed_tokens.weight", w
        elif name == "wpe.weight":
            yield "model.embed_positions.weight", w
        elif name.startswith("ln_f."):
            yield "model.norm." + name[5:], w
        elif name.startswith("h."):
            _, lid, rest = name.split(".", 2)
            p = f"model.layers.{lid}."
            t = (lambda x: x.t().contiguous()) if rest.endswith("weight") else (lambda x: x)
            if rest.startswith("ln_1."):
                yield p + "input_layernorm." + rest[5:], w
            elif rest.startswith("ln_2."):
                yield p + "post_attention_layernorm." + rest[5:], w
            elif rest.startswith("attn.c_attn."):
                full = t(w)                                   # [3H, H] or [3H]
                for j, nm in enumerate(("q_proj", "k_proj", "v_proj")):
                    yield p + f"self_attn.{nm}." + rest.rsplit(".", 1)[1], full[j * H:(j + 1) * H]
            elif rest.startswith("attn.c_proj."):
                yield p + "self_attn.o_proj." + rest.rsplit(".", 1)[1], t(w)
            elif rest.startswith("mlp.c_fc."):
                yield p + "mlp.fc1." + rest.rsplit(".", 1)[1], t(w)
            elif rest.startswith("mlp.c_proj."):
                yield p + "mlp.fc2." + rest.rsplit(".", 1)[1], t(w)
        elif name == "lm_head.weight":
            yield name, w


class OPTForCausalLM(ClassicForCausalLM):
    def remap(self, name, w):
        name = name.replace("model.decoder.", "model.")
        name = name.replace(".self_attn.out_proj.", ".self_attn.o_proj.").replace(".self_attn_layer_norm.", ".input_layernorm.")
        name = name.replace(".final_layer_norm.", ".post_attention_layernorm.") if ".layers." in name else name.replace("model.final_layer_norm.", "model.norm.")
        name = name.replace(".fc1.", ".mlp.fc1.").replace(".fc2.", ".mlp.fc2.") if ".mlp." not in name else name
        yield name, w


class GPTNeoXForCausalLM(ClassicForCausalLM):
    """``query_key_value`` is interleaved per head: [heads, 3, head_dim]"""

    def remap(self, name, w):
        cfg = self.cfg
        name = name.replace("gpt_neox.", "model.").replace("model.embed_in.", "model.embed_tokens.").replace("model.final_layer_norm.", "model.norm.")
        name = name.replace("embed_out.", "lm_head.")
        if ".attention.query_key_value." in name:
            nh, D = cfg.num_attention_heads, cfg.head_dim
            w3 = w.view(nh, 3, D, *w.shape[1:])
            for j, nm in enumerate(("q_proj", "k_proj", "v_proj")):
                yield name.replace(".attention.query_key_value.", f".self_attn.{nm}."), w3[:, j].reshape(nh * D, *w.shape[1:])
            return
        name = name.replace(".attention.dense.", ".self_attn.o_proj.").replace(".mlp.dense_h_to_4h.", ".mlp.fc1.").replace(".mlp.dense_4h_to_h.", ".mlp.fc2.")
        if "rotary_emb" in name or name.endswith(".attention.bias") or name.endswith(".masked_bias"):
            return
        yield name, w


class PersimmonForCausalLM(GPTNeoXForCausalLM):
    """Adept Persimmon / Fuyu text tower: NeoX-style fused ``query_key_value`` ([heads, 3, head_dim]), per-head q/k LayerNorm, partial rotary,
    squared-ReLU MLP, sequential residual (reference models/persimmon.py)"""

    def remap(self, name, w):
        name = name.replace(".self_attn.query_key_value.", ".attention.query_key_value.").replace(".self_attn.dense.", ".attention.dense.")
        name = name.replace("model.final_layernorm.", "model.norm.")
        yield from super().remap(name, w)


class PhiForCausalLM(ClassicForCausalLM):
    """phi-1 / phi-2: parallel block with one shared LayerNorm, partial rotary, biased lm_head"""

    def remap(self, name, w):
        yield name.replace(".self_attn.dense.", ".self_attn.o_proj.").replace("model.final_layernorm.", "model.norm."), w


class GPTJForCausalLM(ClassicForCausalLM):
    def remap(self, name, w):
        if name.startswith("model."):
            yield name, w
            return
        name = name.replace("transformer.wte.", "model.embed_tokens.").replace("transformer.ln_f.", "model.norm.").replace("transformer.h.", "model.layers.")
        name = name.replace(".ln_1.", ".input_layernorm.").replace(".attn.out_proj.", ".self_attn.o_proj.").replace(".attn.", ".self_attn.")
        name = name.replace(".mlp.fc_in.", ".mlp.fc1.").replace(".mlp.fc_out.", ".mlp.fc2.")
        if name.endswith((".self_attn.bias", ".masked_bias")):
            return
        yield name, w


class CodeGenForCausalLM(GPTJForCausalLM):
    """Salesforce CodeGen: the GPT-J block with a fused ``qkv_proj`` whose rows are grouped in ``mp_num = 4`` model-parallel blocks of
    [q | v | k] (that order) — reference models/codegen-style checkpoints"""

    def remap(self, name, w):
        if ".attn.qkv_proj." in name:
            mp, H = 4, self.cfg.hidden_size
            w4 = w.view(mp, 3, H // mp, *w.shape[1:])                # [block][q, v, k][rows]
            base = name.replace("transformer.h.", "model.layers.").replace(".attn.qkv_proj.", ".self_attn.{}.")
            for j, nm in ((0, "q_proj"), (2, "k_proj"), (1, "v_proj")):
                yield base.format(nm), w4[:, j].reshape(H, *w.shape[1:])
            return
        if name.endswith((".attn.causal_mask", ".attn.bias", ".masked_bias")):
            return
        yield from super().remap(name, w)


class GPTNeoForCausalLM(ClassicForCausalLM):
    """EleutherAI GPT-Neo: GPT-2 skeleton with nn.Linear weights, alternating global / local (window 256) attention and NO 1/sqrt(d) scaling
    of the attention scores"""

    def __init__(self, cfg: ModelConfig, quant_name: Optional[str] = None, device="cpu"):
        super().__init__(cfg, quant_name, device)
        lt = cfg.layer_types or []
        for i, layer in enumerate(self.model.layers):
            a = layer.self_attn.attn
            a.scale = 1.0
            a.window = cfg.sliding_window if (i < len(lt) and lt[i] == "sliding_attention" and cfg.sliding_window) else -1

    def remap(self, name, w):
        if name.startswith("model."):
            yield name, w
            return
        name = name[len("transformer."):] if name.startswith("transformer.") else name
        if name == "wte.weight":
            yield "model.embed_tokens.weight", w
        elif name == "wpe.weight":
            yield "model.embed_positions.weight", w
        elif name.startswith("ln_f."):
            yield "model.norm." + name[5:], w
        elif name.startswith("h."):
            _, lid, rest = name.split(".", 2)
            p, leaf = f"model.layers.{lid}.", rest.rsplit(".", 1)[1]
            if rest.startswith("ln_1."):
                yield p + "input_layernorm." + leaf, w
            elif rest.startswith("ln_2."):
                yield p + "post_attention_layernorm." + leaf, w
            elif rest.startswith("attn.attention."):
                sub = rest.split(".")[2]
                if sub in ("q_proj", "k_proj", "v_proj"):
                    yield p + f"self_attn.{sub}." + leaf, w
                elif sub == "out_proj":
                    yield p + "self_attn.o_proj." + leaf, w
            elif rest.startswith("mlp.c_fc."):
                yield p + "mlp.fc1." + leaf, w
            elif rest.startswith("mlp.c_proj."):
                yield p + "mlp.fc2." + leaf, w
        elif name == "lm_head.weight":
            yield name, w


class GPTBigCodeForCausalLM(ClassicForCausalLM):
    """StarCoder-1 / SantaCoder: GPT-2 layout with nn.Linear weights and multi-query attention (fused c_attn = [q | k | v])"""

    def remap(self, name, w):
        cfg = self.cfg
        if name.startswith("model."):
            yield name, w
            return
        name = name[len("transformer."):] if name.startswith("transformer.") else name
        if name == "wte.weight":
            yield "model.embed_tokens.weight", w
        elif name == "wpe.weight":
            yield "model.embed_positions.weight", w
        elif name.startswith("ln_f."):
            yield "model.norm." + name[5:], w
        elif name.startswith("h."):
            _, lid, rest = name.split(".", 2)
            p, leaf = f"model.layers.{lid}.", rest.rsplit(".", 1)[1]
            if rest.startswith("ln_1."):
                yield p + "input_layernorm." + leaf, w
            elif rest.startswith("ln_2."):
                yield p + "post_attention_layernorm." + leaf, w
            elif rest.startswith("attn.c_attn."):
                q, kv = cfg.num_attention_heads * cfg.head_dim, cfg.num_key_value_heads * cfg.head_dim
                for nm, part in zip(("q_proj", "k_proj", "v_proj"), w.split([q, kv, kv], dim=0)):
                    yield p + f"self_attn.{nm}." + leaf, part
            elif rest.startswith("attn.c_proj."):
                yield p + "self_attn.o_proj." + leaf, w
            elif rest.startswith("mlp.c_fc."):
                yield p + "mlp.fc1." + leaf, w
            elif rest.startswith("mlp.c_proj."):
                yield p + "mlp.fc2." + leaf, w
        elif name == "lm_head.weight":
            yield name, w
