"""Whisper (speech -> text): convolutional + transformer audio encoder, transformer decoder with cross-attention.

Parity: reference ``python/sglang/srt/models/whisper.py`` + ``entrypoints/openai/serving_transcription.py`` (``/v1/audio/transcriptions``).

How an encoder-decoder fits the decoder-only engine:

* the *decoder* is an ordinary pre-LayerNorm decoder with learned positions on the paged KV cache (continuous batching, chunked prefill, CUDA
  graphs of the self-attention path all apply); every layer has one extra cross-attention block;
* the *encoder* runs once per request, in the first prefill chunk (``prepare_batch`` hook of the model runner): its output is projected to the
  cross-attention K/V of every decoder layer and parked in a ``cross`` pool indexed by the request slot — [layers, 2, slots, frames, heads*dim];
  decode steps gather their rows by slot (device indices only);
* two requests with the same decoder prompt but different audio must not share KV, so prefix caching is disabled for this family.
"""
from __future__ import annotations

import zlib
from typing import Any, Iterable, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

from .. import ops
from ..configs.model_config import ModelConfig
from ..layers.linear import ColumnParallelLinear, RowParallelLinear, VocabParallelEmbedding
from ..model_executor.forward_batch import ForwardBatch
from ..parallel.parallel_state import get_pp_group, get_tp_group
from .classic import ClassicAttention, ClassicMLP, LayerNorm
from .llama import _DT


class _EncLayer(nn.Module):
    def __init__(self, H: int, heads: int, ffn: int, dtype):
        super().__init__()
        self.heads = heads
        self.self_attn_layer_norm = nn.LayerNorm(H, dtype=dtype)
        self.q_proj, self.v_proj, self.out_proj = nn.Linear(H, H, dtype=dtype), nn.Linear(H, H, dtype=dtype), nn.Linear(H, H, dtype=dtype)
        self.k_proj = nn.Linear(H, H, bias=False, dtype=dtype)
        self.final_layer_norm = nn.LayerNorm(H, dtype=dtype)
        self.fc1, self.fc2 = nn.Linear(H, ffn, dtype=dtype), nn.Linear(ffn, H, dtype=dtype)

    def forward(self, x, mask=None):           # [B, S, H]; mask: boolean key-padding mask broadcastable to [B, heads, S, S] (Qwen2-Audio)
        B, S, H = x.shape
        h = self.self_attn_layer_norm(x)
        sp = lambda t: t.view(B, S, self.heads, H // self.heads).transpose(1, 2)
        a = F.scaled_dot_product_attention(sp(self.q_proj(h)), sp(self.k_proj(h)), sp(self.v_proj(h)), attn_mask=mask)
        x = x + self.out_proj(a.transpose(1, 2).reshape(B, S, H))
        return x + self.fc2(F.gelu(self.fc1(self.final_layer_norm(x))))


class WhisperEncoder(nn.Module):
    def __init__(self, a: dict, dtype):
        super().__init__()
        H = a["d_model"]
        self.conv1 = nn.Conv1d(a["num_mel_bins"], H, 3, padding=1, dtype=dtype)
        self.conv2 = nn.Conv1d(H, H, 3, stride=2, padding=1, dtype=dtype)
        self.embed_positions = nn.Embedding(a["max_source_positions"], H, dtype=dtype)
        self.layers = nn.ModuleList([_EncLayer(H, a["encoder_attention_heads"], a["encoder_ffn_dim"], dtype) for _ in range(a["encoder_layers"])])
        self.layer_norm = nn.LayerNorm(H, dtype=dtype)

    def forward(self, feats: torch.Tensor) -> torch.Tensor:          # [B, mel, frames] -> [B, frames / 2, H]
        x = F.gelu(self.conv2(F.gelu(self.conv1(feats)))).transpose(1, 2)
        x = x + self.embed_positions.weight[: x.shape[1]]
        for l in self.layers:
            x = l(x)
        return self.layer_norm(x)


class CrossAttention(nn.Module):
    def __init__(self, cfg: ModelConfig, layer_id: int, dtype, prefix: str):
        super().__init__()
        H, nh, hd = cfg.hidden_size, cfg.num_attention_heads, cfg.head_dim
        tp = get_tp_group()
        self.q_proj = ColumnParallelLinear(H, nh * hd, True, False, dtype, None, prefix + ".q_proj")
        self.k_proj = ColumnParallelLinear(H, nh * hd, False, False, dtype, None, prefix + ".k_proj")
        self.v_proj = ColumnParallelLinear(H, nh * hd, True, False, dtype, None, prefix + ".v_proj")
        self.o_proj = RowParallelLinear(nh * hd, H, True, True, dtype, None, prefix + ".o_proj")
        self.heads, self.hd, self.layer_id = nh // tp.world_size, hd, layer_id

    def project_kv(self, enc: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:       # [S, H] -> [S, heads*hd] x 2
        return self.k_proj(enc), self.v_proj(enc)

    def forward(self, x: torch.Tensor, fb: ForwardBatch, pool: torch.Tensor, lens: torch.Tensor) -> torch.Tensor:
        """x [T, H]; pool [2, slots, S, heads*hd]; every token attends to the encoder frames of its request"""
        T = x.shape[0]
        q = self.q_proj(x).view(T, self.heads, self.hd)
        slots = fb.req_slots.long()
        if fb.cu_q is not None:
            slots = torch.repeat_interleave(slots, fb.q_lens.long())
        S = pool.shape[2]
        k = pool[0][slots].view(T, S, self.heads, self.hd).transpose(1, 2)              # [T, heads, S, hd]
        v = pool[1][slots].view(T, S, self.heads, self.hd).transpose(1, 2)
        mask = (torch.arange(S, device=x.device)[None, :] < lens[slots][:, None])[:, None, None, :]
        o = F.scaled_dot_product_attention(q.unsqueeze(2), k, v, attn_mask=mask)        # [T, heads, 1, hd]
        return self.o_proj(o.reshape(T, self.heads * self.hd))


class WhisperDecoderLayer(nn.Module):
    def __init__(self, cfg: ModelConfig, layer_id: int, dtype, prefix: str):
        super().__init__()
        H = cfg.hidden_size
        self.input_layernorm = LayerNorm(H, cfg.rms_norm_eps, dtype)
        self.self_attn = ClassicAttention(cfg, layer_id, dtype, None, None, prefix + ".self_attn")
        self.encoder_attn_layer_norm = LayerNorm(H, cfg.rms_norm_eps, dtype)
        self.encoder_attn = CrossAttention(cfg, layer_id, dtype, prefix + ".encoder_attn")
        self.post_attention_layernorm = LayerNorm(H, cfg.rms_norm_eps, dtype)
        self.mlp = ClassicMLP(cfg, dtype, None, prefix + ".mlp")

    def forward(self, x, fb, cross, lens):
        x = x + self.self_attn(self.input_layernorm(x), fb)
        x = x + self.encoder_attn(self.encoder_attn_layer_norm(x), fb, cross, lens)
        return x + self.mlp(self.post_attention_layernorm(x))


class WhisperDecoder(nn.Module):
    def __init__(self, cfg: ModelConfig):
        super().__init__()
        dtype = _DT[cfg.dtype]
        self.embed_tokens = VocabParallelEmbedding(cfg.vocab_size, cfg.hidden_size, dtype)
        self.embed_positions = nn.Embedding(cfg.max_position_embeddings, cfg.hidden_size, dtype=dtype)
        self.layers = nn.ModuleList([WhisperDecoderLayer(cfg, i, dtype, f"model.layers.{i}") for i in range(cfg.num_hidden_layers)])
        self.norm = LayerNorm(cfg.hidden_size, cfg.rms_norm_eps, dtype)


class WhisperForConditionalGeneration(nn.Module):
    supports_cuda_graph = False
    disable_radix_cache = True
    stacked_params = [(".qkv_proj", ".q_proj", "q"), (".qkv_proj", ".k_proj", "k"), (".qkv_proj", ".v_proj", "v")]

    def __init__(self, cfg: ModelConfig, quant_name: Optional[str] = None, device="cpu"):
        super().__init__()
        assert get_pp_group().world_size == 1
        self.cfg = cfg
        dtype = _DT[cfg.dtype]
        self.encoder = WhisperEncoder(cfg.audio, dtype)
        self.model = WhisperDecoder(cfg)
        self.cross: Optional[torch.Tensor] = None          # [L, 2, slots, S, heads*hd]
        self.cross_len: Optional[torch.Tensor] = None      # [slots] valid encoder frames per request

    # ---- runner hook: encode the audio of requests that start their prefill in this batch
    @torch.no_grad()
    def prepare_batch(self, sb: Any, fb: ForwardBatch, runner: Any) -> None:
        dev, dtype = runner.device, _DT[self.cfg.dtype]
        S = self.cfg.audio["max_source_positions"]
        if self.cross is None:
            L, w = len(self.model.layers), self.model.layers[0].encoder_attn.heads * self.cfg.head_dim
            self.cross = torch.zeros(L, 2, runner.max_running + 1, S, w, dtype=dtype, device=dev)
            self.cross_len = torch.zeros(runner.max_running + 1, dtype=torch.long, device=dev)
        for i, r in enumerate(sb.reqs):
            if int(sb.prefix_lens[i]) != 0:
                continue
            slot = int(sb.req_slots[i])
            mm = r.mm_inputs or {}
            if mm.get("input_features") is None:
                self.cross_len[slot] = 0
                continue
            feats = torch.as_tensor(np.asarray(mm["input_features"], dtype=np.float32), device=dev).to(dtype)
            enc = self.encoder(feats[None])[0]                                      # [S', H]
            n = min(enc.shape[0], S)
            for li, layer in enumerate(self.model.layers):
                k, v = layer.encoder_attn.project_kv(enc[:n])
                self.cross[li, 0, slot, :n] = k
                self.cross[li, 1, slot, :n] = v
            self.cross_len[slot] = n

    def forward(self, fb: ForwardBatch, inputs_embeds=None):
        d = self.model
        h = d.embed_tokens(fb.next_token_buf, fb.req_slots) if fb.input_ids is None else d.embed_tokens(fb.input_ids)
        h = h + d.embed_positions(fb.positions.long())
        for li, layer in enumerate(d.layers):
            h = layer(h, fb, self.cross[li], self.cross_len)
        return d.norm(h)

    def compute_logits(self, hidden: torch.Tensor, key: str = "logits") -> torch.Tensor:
        y = ops.linear(hidden, self.model.embed_tokens.weight)          # proj_out is tied to the token embedding
        tp = get_tp_group()
        if tp.world_size > 1:
            y = tp.all_gather(y, dim=-1)
        return y[:, : self.cfg.vocab_size]

    # ---- weights (HF: model.encoder.* / model.decoder.* / proj_out)
    def load_weights(self, weights: Iterable[Tuple[str, torch.Tensor]]) -> List[str]:
        from .llama import _load_stacked_bias

        params = dict(self.named_parameters())
        mods = dict(self.named_modules())
        loaded = []
        for name, w in weights:
            if name.startswith("proj_out"):
                continue
            if name.startswith("model.encoder."):
                n = name[len("model."):].replace(".self_attn.", ".")
                if n in params:
                    params[n].data.copy_(w)
                    loaded.append(name)
                continue
            n = name.replace("model.decoder.", "model.").replace("model.layer_norm.", "model.norm.")
            n = n.replace(".self_attn_layer_norm.", ".input_layernorm.").replace(".final_layer_norm.", ".post_attention_layernorm.")
            n = n.replace(".self_attn.out_proj.", ".self_attn.o_proj.").replace(".encoder_attn.out_proj.", ".encoder_attn.o_proj.")
            n = n.replace(".fc1.", ".mlp.fc1.").replace(".fc2.", ".mlp.fc2.") if ".mlp." not in n else n
            target, shard = n, None
            if ".self_attn." in n:
                for merged, orig, sid in self.stacked_params:
                    if orig + "." in n:
                        target, shard = n.replace(orig, merged), sid
                        break
            if target not in params:
                continue
            p, mod = params[target], mods[target.rsplit(".", 1)[0]]
            if hasattr(mod, "weight_loader"):
                if target.endswith(".bias") and shard is not None:
                    _load_stacked_bias(mod, p, w, shard)
                elif target.endswith(".bias") and p.dim() == 1 and not isinstance(mod, RowParallelLinear):
                    k = p.shape[0]
                    p.data.copy_(w.narrow(0, get_tp_group().rank * k, k) if w.shape[0] != k else w)
                else:
                    mod.weight_loader(p, w, shard) if shard is not None else mod.weight_loader(p, w)
            else:
                p.data.copy_(w[: p.shape[0]] if w.shape != p.shape else w)
            loaded.append(name)
        return loaded

    def dummy_checkpoint(self, cfg: ModelConfig, device, dtype, seed: int = 0, scale: float = 0.05):
        for n, p in self.named_parameters():
            g = torch.Generator(device="cpu").manual_seed((seed * 1000003 + zlib.crc32(n.encode())) & 0x7FFFFFFF)
            if p.dim() > 1:
                w = torch.randn(p.shape, generator=g) * scale
            elif "norm" in n and n.endswith("weight"):
                w = torch.ones(p.shape)
            else:
                w = torch.randn(p.shape, generator=g) * 0.02
            p.data.copy_(w.to(p.dtype))
        return iter(())
