"""Encoder-only models for embeddings and cross-encoder scoring: BERT and (XLM-)RoBERTa.

Parity: reference ``python/sglang/srt/models/{bert,roberta}.py`` + ``layers/pooler.py`` (CLS pooling, ``CrossEncodingPooler``).

They run through the same engine as decoders (requests are batched ragged by the scheduler) but need nothing from the KV cache: every request
is encoded in ONE pass with bidirectional attention, so chunked prefill and prefix caching are switched off for these architectures
(``disable_chunked_prefill`` / ``disable_radix_cache``).  The pooled vector of a request (its [CLS] row, optionally through the BERT pooler and a
classification head) is written to the request's LAST row, which is the row the runner hands to ``pool``.
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Tuple

import torch
import torch.nn.functional as F
from torch import nn

from ..configs.model_config import ModelConfig
from ..model_executor.forward_batch import ForwardBatch
from .llama import _DT


class _Layer(nn.Module):
    def __init__(self, H: int, heads: int, inter: int, eps: float, act: str, dtype):
        super().__init__()
        self.heads = heads
        self.query, self.key, self.value, self.attn_out = (nn.Linear(H, H, dtype=dtype) for _ in range(4))
        self.attn_norm = nn.LayerNorm(H, eps=eps, dtype=dtype)
        self.inter, self.out = nn.Linear(H, inter, dtype=dtype), nn.Linear(inter, H, dtype=dtype)
        self.out_norm = nn.LayerNorm(H, eps=eps, dtype=dtype)
        self.act = F.gelu if act == "gelu" else (lambda t: F.gelu(t, approximate="tanh")) if "tanh" in act or act == "gelu_new" else torch.relu

    def forward(self, x: torch.Tensor, lens: List[int]) -> torch.Tensor:
        T, H = x.shape
        hd = H // self.heads
        q, k, v = (t.view(T, self.heads, hd) for t in (self.query(x), self.key(x), self.value(x)))
        outs, s = [], 0
        for n in lens:          # bidirectional attention inside each request
            o = F.scaled_dot_product_attention(q[s:s + n].transpose(0, 1), k[s:s + n].transpose(0, 1), v[s:s + n].transpose(0, 1))
            outs.append(o.transpose(0, 1).reshape(n, H))
            s += n
        x = self.attn_norm(x + self.attn_out(torch.cat(outs, 0) if outs else x[:0]))
        return self.out_norm(x + self.out(self.act(self.inter(x))))


class _Backbone(nn.Module):
    layers: list = []          # no paged-KV attention layers: the runner sizes an empty KV pool


class BertModel(nn.Module):
    """``BertModel`` / ``XLMRobertaModel``: normalised [CLS] embedding"""
    is_embedding = True
    supports_cuda_graph = False
    disable_radix_cache = True
    disable_chunked_prefill = True
    pos_offset = 0
    prefix = "bert"
    use_segments = True         # BERT tokenisers mark the second sentence of a pair with token type 1; the RoBERTa family never does

    def __init__(self, cfg: ModelConfig, quant_name: Optional[str] = None, device="cpu"):
        super().__init__()
        self.cfg = cfg
        dtype = _DT[cfg.dtype]
        H, e = cfg.hidden_size, cfg.rms_norm_eps
        x = cfg.encoder or {}
        self.word_embeddings = nn.Embedding(cfg.vocab_size, H, dtype=dtype)
        self.position_embeddings = nn.Embedding(cfg.max_position_embeddings, H, dtype=dtype)
        self.token_type_embeddings = nn.Embedding(x.get("type_vocab_size", 2), H, dtype=dtype)
        self.emb_norm = nn.LayerNorm(H, eps=e, dtype=dtype)
        self.enc = nn.ModuleList([_Layer(H, cfg.num_attention_heads, cfg.intermediate_size, e, cfg.hidden_act, dtype) for _ in range(cfg.num_hidden_layers)])
        self.model = _Backbone()
        self.sep_token_id = x.get("sep_token_id")
        for p in self.parameters():
            p.requires_grad_(False)

    # ---- encoding
    def encode(self, fb: ForwardBatch) -> Tuple[torch.Tensor, List[int]]:
        lens = [int(n) for n in fb.q_lens_cpu]
        ids = fb.input_ids.long()
        types = torch.zeros_like(ids)
        if self.use_segments and self.sep_token_id is not None and self.token_type_embeddings.num_embeddings > 1:      # sentence pairs: segment B starts after the first [SEP]
            s = 0
            for n in lens:
                seps = (ids[s:s + n] == self.sep_token_id).nonzero().flatten()
                if seps.numel() >= 2:
                    types[s + int(seps[0]) + 1: s + n] = 1
                s += n
        h = self.word_embeddings(ids) + self.position_embeddings(fb.positions.long() + self.pos_offset) + self.token_type_embeddings(types)
        h = self.emb_norm(h)
        for l in self.enc:
            h = l(h, lens)
        return h, lens

    def pooled(self, h: torch.Tensor, lens: List[int]) -> torch.Tensor:
        """one vector per request, from its first ([CLS] / <s>) row"""
        first = torch.tensor([sum(lens[:i]) for i in range(len(lens))], dtype=torch.long, device=h.device)
        return h[first]

    def forward(self, fb: ForwardBatch, inputs_embeds=None) -> torch.Tensor:
        h, lens = self.encode(fb)
        if not lens:
            return h
        last = torch.tensor([sum(lens[:i + 1]) - 1 for i in range(len(lens))], dtype=torch.long, device=h.device)
        vec = self.head(self.pooled(h, lens))
        out = torch.zeros(h.shape[0], vec.shape[-1], dtype=vec.dtype, device=h.device)
        out[last] = vec
        return out

    def head(self, cls: torch.Tensor) -> torch.Tensor:
        return cls

    def pool(self, hidden_last: torch.Tensor, normalize: bool = True) -> torch.Tensor:
        return F.normalize(hidden_last.float(), dim=-1)

    def compute_logits(self, hidden, key: str = "logits"):
        raise RuntimeError("encoder-only model: use /v1/embeddings, /v1/rerank or /v1/classify")

    # ---- weights
    def _map(self, n: str) -> Optional[str]:
        for pre in (self.prefix + ".", "roberta.", "bert.", "model."):
            if n.startswith(pre):
                n = n[len(pre):]
                break
        n = n.replace("embeddings.LayerNorm.", "emb_norm.")
        n = n[len("embeddings."):] if n.startswith("embeddings.") else n
        n = n.replace("encoder.layer.", "enc.")
        n = n.replace(".attention.self.", ".").replace(".attention.output.dense.", ".attn_out.").replace(".attention.output.LayerNorm.", ".attn_norm.")
        n = n.replace(".intermediate.dense.", ".inter.").replace(".output.dense.", ".out.").replace(".output.LayerNorm.", ".out_norm.")
        return n

    def load_weights(self, weights: Iterable[Tuple[str, torch.Tensor]]) -> List[str]:
        params = dict(self.named_parameters())
        loaded = []
        for name, w in weights:
            n = self._map(name)
            if n in params:
                params[n].data.copy_(w)
                loaded.append(name)
        return loaded

    def dummy_checkpoint(self, cfg, device, dtype, seed: int = 0, scale: float = 0.05):
        import zlib

        for n, p in self.named_parameters():
            g = torch.Generator(device="cpu").manual_seed((seed * 1000003 + zlib.crc32(n.encode())) & 0x7FFFFFFF)
            p.data.copy_((torch.randn(p.shape, generator=g) * scale).to(p.dtype) if p.dim() > 1 else
                         (torch.ones(p.shape) if "norm" in n and n.endswith("weight") else torch.zeros(p.shape)).to(p.dtype))
        return iter(())


class XLMRobertaModel(BertModel):
    pos_offset = 2          # RoBERTa positions start after the padding index (1)
    prefix = "roberta"
    use_segments = False


class BertForSequenceClassification(BertModel):
    """cross-encoder / classifier: tanh pooler over [CLS] then a linear head (``/v1/rerank``, ``/v1/classify``)"""

    def __init__(self, cfg: ModelConfig, quant_name: Optional[str] = None, device="cpu"):
        super().__init__(cfg, quant_name, device)
        dtype = _DT[cfg.dtype]
        self.pooler_dense = nn.Linear(cfg.hidden_size, cfg.hidden_size, dtype=dtype)
        self.classifier = nn.Linear(cfg.hidden_size, cfg.num_labels, dtype=dtype)

    def head(self, cls):
        return self.classifier(torch.tanh(self.pooler_dense(cls)))

    def pool(self, hidden_last, normalize: bool = False):
        return hidden_last.float()

    def _map(self, n):
        if n.startswith("classifier."):
            return n
        n = super()._map(n)
        return n.replace("pooler.dense.", "pooler_dense.")


class XLMRobertaForSequenceClassification(XLMRobertaModel):
    """bge-reranker style head: dense -> tanh -> out_proj on the <s> row"""

    def __init__(self, cfg: ModelConfig, quant_name: Optional[str] = None, device="cpu"):
        super().__init__(cfg, quant_name, device)
        dtype = _DT[cfg.dtype]
        self.cls_dense = nn.Linear(cfg.hidden_size, cfg.hidden_size, dtype=dtype)
        self.cls_out = nn.Linear(cfg.hidden_size, cfg.num_labels, dtype=dtype)

    def head(self, cls):
        return self.cls_out(torch.tanh(self.cls_dense(cls)))

    def pool(self, hidden_last, normalize: bool = False):
        return hidden_last.float()

    def _map(self, n):
        if n.startswith("classifier."):
            return n.replace("classifier.dense.", "cls_dense.").replace("classifier.out_proj.", "cls_out.")
        return super()._map(n)


class Contriever(BertModel):
    """facebook/contriever (reference ``models/bert.py`` Contriever): BERT encoder, the sentence vector is the MEAN over the token rows,
    not normalised"""

    def pooled(self, h: torch.Tensor, lens: List[int]) -> torch.Tensor:
        out, s = [], 0
        for n in lens:
            out.append(h[s:s + n].float().mean(0).to(h.dtype))
            s += n
        return torch.stack(out) if out else h[:0]

    def pool(self, hidden_last: torch.Tensor, normalize: bool = True) -> torch.Tensor:
        return hidden_last.float()
