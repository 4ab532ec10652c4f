"""GGUF checkpoints: container parser, block de-quantisers, llama.cpp -> canonical name mapping, config + tokenizer from metadata,
and a writer (used by the tests and by ``tools/`` to produce fixtures offline).

Parity: reference ``python/sglang/srt/model_loader/loader.py`` (``GGUFModelLoader``), ``weight_utils.py`` (``gguf_quant_weights_iterator``)
and ``layers/quantization/gguf.py``.  The reference keeps GGUF blocks quantised on the GPU and calls ggml-derived CUDA kernels; here a
GGUF file is an *input format*: tensors are de-quantised once at load (vectorised NumPy, all legacy and K-quant block types that
llama.cpp's converters emit by default) into the serving dtype, so every sm_100a kernel in the tree sees ordinary bf16 weights (or they
can be re-quantised to the engine's own fp8 / W4A16 formats with ``--quantization``).

Format (GGUF v2/v3, little endian):  ``"GGUF" u32 version, u64 n_tensors, u64 n_kv``, then ``n_kv`` (string key, u32 type, value) pairs,
then ``n_tensors`` (string name, u32 n_dims, u64 dims[n_dims] innermost first, u32 ggml_type, u64 offset), then the data section aligned
to ``general.alignment`` (default 32).
"""
from __future__ import annotations

import os
import struct
from typing import Any, BinaryIO, Dict, Iterator, List, Optional, Tuple

import numpy as np
import torch

GGUF_MAGIC = 0x46554747
# ggml_type -> (name, block elements, block bytes)
GGML_TYPES: Dict[int, Tuple[str, int, int]] = {
    0: ("F32", 1, 4), 1: ("F16", 1, 2), 2: ("Q4_0", 32, 18), 3: ("Q4_1", 32, 20), 6: ("Q5_0", 32, 22), 7: ("Q5_1", 32, 24),
    8: ("Q8_0", 32, 34), 10: ("Q2_K", 256, 84), 11: ("Q3_K", 256, 110), 12: ("Q4_K", 256, 144), 13: ("Q5_K", 256, 176),
    14: ("Q6_K", 256, 210), 30: ("BF16", 1, 2),
}
_NAME_TO_TYPE = {v[0]: k for k, v in GGML_TYPES.items()}
_SCALAR = {0: "<B", 1: "<b", 2: "<H", 3: "<h", 4: "<I", 5: "<i", 6: "<f", 7: "<?", 10: "<Q", 11: "<q", 12: "<d"}


# ------------------------------------------------------------------------------------------------ container
def _rd(f: BinaryIO, fmt: str):
    return struct.unpack(fmt, f.read(struct.calcsize(fmt)))[0]


def _rd_str(f: BinaryIO) -> str:
    return f.read(_rd(f, "<Q")).decode("utf-8", errors="replace")


def _rd_value(f: BinaryIO, t: int):
    if t in _SCALAR:
        return _rd(f, _SCALAR[t])
    if t == 8:
        return _rd_str(f)
    if t == 9:
        et, n = _rd(f, "<I"), _rd(f, "<Q")
        if et in _SCALAR and et != 7:
            dt = np.dtype(_SCALAR[et])
            return np.frombuffer(f.read(n * dt.itemsize), dtype=dt).tolist()
        return [_rd_value(f, et) for _ in range(n)]
    raise ValueError(f"gguf: unknown metadata value type {t}")


class GGUFFile:
    """memory-mapped GGUF file: ``metadata`` dict, ``tensors`` {name: (shape_torch_order, ggml_type, offset)}"""

    def __init__(self, path: str):
        self.path = path
        with open(path, "rb") as f:
            if _rd(f, "<I") != GGUF_MAGIC:
                raise ValueError(f"{path}: not a GGUF file")
            self.version = _rd(f, "<I")
            if self.version not in (2, 3):
                raise ValueError(f"{path}: GGUF version {self.version} not supported")
            nt, nkv = _rd(f, "<Q"), _rd(f, "<Q")
            self.metadata: Dict[str, Any] = {}
            for _ in range(nkv):
                k = _rd_str(f)
                self.metadata[k] = _rd_value(f, _rd(f, "<I"))
            self.tensors: Dict[str, Tuple[Tuple[int, ...], int, int]] = {}
            for _ in range(nt):
                name = _rd_str(f)
                nd = _rd(f, "<I")
                dims = [_rd(f, "<Q") for _ in range(nd)]
                typ, off = _rd(f, "<I"), _rd(f, "<Q")
                self.tensors[name] = (tuple(reversed(dims)), typ, off)
            align = int(self.metadata.get("general.alignment", 32))
            self.data_start = (f.tell() + align - 1) // align * align
        self._mm = np.memmap(path, dtype=np.uint8, mode="r")

    @property
    def architecture(self) -> str:
        return self.metadata.get("general.architecture", "llama")

    def raw(self, name: str) -> np.ndarray:
        shape, typ, off = self.tensors[name]
        if typ not in GGML_TYPES:
            raise NotImplementedError(f"gguf: tensor {name} has unsupported ggml type {typ}")
        _, be, bb = GGML_TYPES[typ]
        n = int(np.prod(shape))
        assert n % be == 0, (name, shape, typ)
        a = self.data_start + off
        return self._mm[a:a + n // be * bb]

    def get(self, name: str) -> torch.Tensor:
        """de-quantised fp32 (F32 / quantised types) or native (F16 / BF16) tensor in torch (row-major, outermost-first) shape"""
        shape, typ, _ = self.tensors[name]
        return dequantize(self.raw(name), typ, shape)


# ------------------------------------------------------------------------------------------------ de-quantisers
def _f16(b: np.ndarray) -> np.ndarray:
    return np.ascontiguousarray(b).view(np.float16).astype(np.float32)


def _scale_min_k4(sc: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """the 12-byte packed 6-bit (scale, min) x 8 of Q4_K / Q5_K: sc [nb, 12] u8 -> ([nb, 8], [nb, 8])"""
    s = np.empty(sc.shape[:-1] + (8,), np.float32)
    m = np.empty_like(s)
    s[..., :4] = sc[..., 0:4] & 63
    m[..., :4] = sc[..., 4:8] & 63
    s[..., 4:] = (sc[..., 8:12] & 0xF) | ((sc[..., 0:4] >> 6) << 4)
    m[..., 4:] = (sc[..., 8:12] >> 4) | ((sc[..., 4:8] >> 6) << 4)
    return s, m


def dequantize(raw: np.ndarray, typ: int, shape: Tuple[int, ...]) -> torch.Tensor:
    name, be, bb = GGML_TYPES[typ]
    if name == "F32":
        return torch.from_numpy(np.ascontiguousarray(raw).view(np.float32).reshape(shape).copy())
    if name == "F16":
        return torch.from_numpy(np.ascontiguousarray(raw).view(np.float16).reshape(shape).copy())
    if name == "BF16":
        return torch.from_numpy(np.ascontiguousarray(raw).view(np.int16).reshape(shape).copy()).view(torch.bfloat16)
    b = np.ascontiguousarray(raw).reshape(-1, bb)
    nb = b.shape[0]
    if name == "Q8_0":
        y = _f16(b[:, 0:2]) * b[:, 2:].view(np.int8).astype(np.float32)
    elif name in ("Q4_0", "Q4_1"):
        o = 2 if name == "Q4_0" else 4
        d = _f16(b[:, 0:2])
        qs = b[:, o:o + 16]
        q = np.concatenate([qs & 0xF, qs >> 4], axis=1).astype(np.float32)
        y = (q - 8.0) * d if name == "Q4_0" else q * d + _f16(b[:, 2:4])
    elif name in ("Q5_0", "Q5_1"):
        o = 2 if name == "Q5_0" else 4
        d = _f16(b[:, 0:2])
        qh = np.ascontiguousarray(b[:, o:o + 4]).view(np.uint32)                         # [nb, 1]
        bits = ((qh >> np.arange(32, dtype=np.uint32)[None, :]) & 1).astype(np.uint8)    # [nb, 32]
        qs = b[:, o + 4:o + 20]
        q = (np.concatenate([qs & 0xF, qs >> 4], axis=1) | (bits << 4)).astype(np.float32)
        y = (q - 16.0) * d if name == "Q5_0" else q * d + _f16(b[:, 2:4])
    elif name in ("Q4_K", "Q5_K"):
        d, dmin = _f16(b[:, 0:2]), _f16(b[:, 2:4])
        sc, mn = _scale_min_k4(b[:, 4:16])
        if name == "Q4_K":
            qs = b[:, 16:144].reshape(nb, 4, 32)
            q = np.stack([qs & 0xF, qs >> 4], axis=2).reshape(nb, 8, 32).astype(np.float32)
        else:
            qh = b[:, 16:48]                                                              # [nb, 32]
            qs = b[:, 48:176].reshape(nb, 4, 32)
            hb = ((qh[:, None, :] >> np.arange(8, dtype=np.uint8)[None, :, None]) & 1).astype(np.uint8)    # [nb, 8, 32]: bit j of qh[l]
            q = (np.stack([qs & 0xF, qs >> 4], axis=2).reshape(nb, 8, 32) | (hb << 4)).astype(np.float32)
        y = (d[:, :, None] * sc[:, :, None]) * q - (dmin[:, :, None] * mn[:, :, None])
    elif name == "Q6_K":
        ql = b[:, 0:128].reshape(nb, 2, 64)
        qh = b[:, 128:192].reshape(nb, 2, 32)
        sc = b[:, 192:208].view(np.int8).astype(np.float32).reshape(nb, 2, 8)
        d = _f16(b[:, 208:210])
        lo = np.stack([ql[:, :, :32] & 0xF, ql[:, :, 32:] & 0xF, ql[:, :, :32] >> 4, ql[:, :, 32:] >> 4], axis=2)    # [nb, 2, 4, 32]
        hi = np.stack([(qh >> s) & 3 for s in (0, 2, 4, 6)], axis=2)                                                  # [nb, 2, 4, 32]
        q = (lo | (hi << 4)).astype(np.float32) - 32.0
        # scale index for (half n, quarter k, lane l) = 8 n + 2 k + l // 16
        scl = sc.reshape(nb, 2, 4, 2)[:, :, :, :, None].repeat(16, axis=4).reshape(nb, 2, 4, 32)
        y = d[:, :, None, None] * scl * q
    elif name == "Q2_K":
        scales = b[:, 0:16]
        qs = b[:, 16:80].reshape(nb, 2, 32)
        d, dmin = _f16(b[:, 80:82]), _f16(b[:, 82:84])
        q = np.stack([(qs >> s) & 3 for s in (0, 2, 4, 6)], axis=2).astype(np.float32)        # [nb, 2, 4, 32]
        scl = scales.reshape(nb, 2, 4, 2)[:, :, :, :, None].repeat(16, axis=4).reshape(nb, 2, 4, 32)
        y = d[:, :, None, None] * (scl & 0xF) * q - dmin[:, :, None, None] * (scl >> 4)
    elif name == "Q3_K":
        hm = b[:, 0:32]
        qs = b[:, 32:96].reshape(nb, 2, 32)
        s12 = b[:, 96:108].astype(np.uint32)
        d = _f16(b[:, 108:110])
        lo4 = np.concatenate([s12[:, 0:8] & 0xF, s12[:, 0:8] >> 4], axis=1)                   # scales 0..15 low 4 bits
        hi2 = np.concatenate([(s12[:, 8:12] >> s) & 3 for s in (0, 2, 4, 6)], axis=1)          # scales 0..15 high 2 bits
        sc = ((lo4 | (hi2 << 4)).astype(np.int32) - 32).astype(np.float32)                     # [nb, 16]
        q = np.stack([(qs >> s) & 3 for s in (0, 2, 4, 6)], axis=2).astype(np.int32)           # [nb, 2, 4, 32]
        bit = np.arange(8, dtype=np.uint8).reshape(1, 2, 4, 1)
        h = ((hm[:, None, None, :] >> bit) & 1).astype(np.int32)                               # mask bit (4 n + k) of hm[l]
        q = (q - 4 * (1 - h)).astype(np.float32)
        scl = sc.reshape(nb, 2, 4, 2)[:, :, :, :, None].repeat(16, axis=4).reshape(nb, 2, 4, 32)
        y = d[:, :, None, None] * scl * q
    else:
        raise NotImplementedError(f"gguf: de-quantiser for {name} not implemented")
    return torch.from_numpy(np.ascontiguousarray(y, dtype=np.float32).reshape(shape))


# ------------------------------------------------------------------------------------------------ quantisers (writer side)
def quantize(x: torch.Tensor, name: str) -> bytes:
    x = x.detach().cpu()
    if name == "F32":
        return x.float().contiguous().numpy().tobytes()
    if name == "F16":
        return x.to(torch.float16).contiguous().numpy().tobytes()
    if name == "BF16":
        return x.to(torch.bfloat16).contiguous().view(torch.int16).numpy().tobytes()
    a = x.float().contiguous().numpy().reshape(-1, 32)
    if name == "Q8_0":
        d = np.abs(a).max(axis=1, keepdims=True) / 127.0
        q = np.where(d > 0, np.round(a / np.where(d > 0, d, 1)), 0).astype(np.int8)
        return np.concatenate([d.astype(np.float16).view(np.uint8), q.view(np.uint8)], axis=1).tobytes()
    if name == "Q4_0":
        idx = np.abs(a).argmax(axis=1)
        mx = a[np.arange(a.shape[0]), idx][:, None]
        d = mx / -8.0
        q = np.clip(np.round(np.where(d != 0, a / np.where(d != 0, d, 1), 0) + 8.0), 0, 15).astype(np.uint8)
        packed = q[:, :16] | (q[:, 16:] << 4)
        return np.concatenate([d.astype(np.float16).view(np.uint8), packed], axis=1).tobytes()
    if name == "Q4_1":
        mn, mx = a.min(axis=1, keepdims=True), a.max(axis=1, keepdims=True)
        d = (mx - mn) / 15.0
        q = np.clip(np.round(np.where(d > 0, (a - mn) / np.where(d > 0, d, 1), 0)), 0, 15).astype(np.uint8)
        packed = q[:, :16] | (q[:, 16:] << 4)
        return np.concatenate([d.astype(np.float16).view(np.uint8), mn.astype(np.float16).view(np.uint8), packed], axis=1).tobytes()
    raise NotImplementedError(f"gguf writer: quantiser for {name} not implemented")


def write_gguf(path: str, metadata: Dict[str, Any], tensors: Dict[str, Tuple[torch.Tensor, str]], alignment: int = 32) -> None:
    """``tensors``: name -> (tensor in torch shape, ggml type name)"""
    def w_str(f, s: str):
        b = s.encode("utf-8")
        f.write(struct.pack("<Q", len(b)) + b)

    def w_val(f, v):
        if isinstance(v, bool):
            f.write(struct.pack("<I?", 7, v))
        elif isinstance(v, int):
            f.write(struct.pack("<II", 4, v) if 0 <= v < 2 ** 32 else struct.pack("<Iq", 11, v))
        elif isinstance(v, float):
            f.write(struct.pack("<If", 6, v))
        elif isinstance(v, str):
            f.write(struct.pack("<I", 8))
            w_str(f, v)
        elif isinstance(v, (list, tuple)):
            et = 8 if (v and isinstance(v[0], str)) else (6 if (v and isinstance(v[0], float)) else 5)
            f.write(struct.pack("<IIQ", 9, et, len(v)))
            for e in v:
                if et == 8:
                    w_str(f, e)
                else:
                    f.write(struct.pack("<f" if et == 6 else "<i", e))
        else:
            raise TypeError(type(v))

    blobs, infos, off = [], [], 0
    for name, (t, tname) in tensors.items():
        blob = quantize(t, tname)
        infos.append((name, tuple(reversed(t.shape)), _NAME_TO_TYPE[tname], off))
        blobs.append(blob)
        off = (off + len(blob) + alignment - 1) // alignment * alignment
    md = {"general.alignment": alignment, **metadata}
    with open(path, "wb") as f:
        f.write(struct.pack("<IIQQ", GGUF_MAGIC, 3, len(infos), len(md)))
        for k, v in md.items():
            w_str(f, k)
            w_val(f, v)
        for name, dims, typ, o in infos:
            w_str(f, name)
            f.write(struct.pack("<I", len(dims)))
            f.write(struct.pack(f"<{len(dims)}Q", *dims))
            f.write(struct.pack("<IQ", typ, o))
        f.write(b"\0" * ((-f.tell()) % alignment))
        for blob in blobs:
            f.write(blob)
            f.write(b"\0" * ((-len(blob)) % alignment))


# ------------------------------------------------------------------------------------------------ names, config, tokenizer
_ARCH = {"llama": "LlamaForCausalLM", "mistral": "LlamaForCausalLM", "qwen2": "Qwen2ForCausalLM", "qwen3": "Qwen3ForCausalLM",
         "qwen2moe": "Qwen2MoeForCausalLM", "qwen3moe": "Qwen3MoeForCausalLM", "gemma": "GemmaForCausalLM", "gemma2": "Gemma2ForCausalLM",
         "phi3": "Phi3ForCausalLM", "internlm2": "InternLM2ForCausalLM", "granite": "GraniteForCausalLM", "olmo2": "Olmo2ForCausalLM",
         "starcoder2": "Starcoder2ForCausalLM", "gpt2": "GPT2LMHeadModel", "gptneox": "GPTNeoXForCausalLM"}
_TOP = {"token_embd": "model.embed_tokens", "output_norm": "model.norm", "output": "lm_head", "position_embd": "model.embed_positions"}
_BLK = {"attn_norm": "input_layernorm", "ffn_norm": "post_attention_layernorm", "attn_q": "self_attn.q_proj", "attn_k": "self_attn.k_proj",
        "attn_v": "self_attn.v_proj", "attn_output": "self_attn.o_proj", "attn_q_norm": "self_attn.q_norm", "attn_k_norm": "self_attn.k_norm",
        "attn_qkv": "self_attn.qkv_proj", "ffn_gate": "mlp.gate_proj", "ffn_up": "mlp.up_proj", "ffn_down": "mlp.down_proj",
        "ffn_gate_inp": "mlp.gate", "post_attention_norm": "post_attention_layernorm", "post_ffw_norm": "post_feedforward_layernorm",
        "ffn_gate_inp_shexp": "mlp.shared_expert_gate", "ffn_gate_shexp": "mlp.shared_expert.gate_proj",
        "ffn_up_shexp": "mlp.shared_expert.up_proj", "ffn_down_shexp": "mlp.shared_expert.down_proj"}
_EXPS = {"ffn_gate_exps": "gate_proj", "ffn_up_exps": "up_proj", "ffn_down_exps": "down_proj"}
_PERMUTED_QK = {"llama", "mistral"}          # convert_hf_to_gguf permutes q/k rows for these (interleaved-pair RoPE in ggml)


def _unpermute(w: torch.Tensor, n_head: int) -> torch.Tensor:
    """inverse of llama.cpp's ``permute``: rows come as [head][pair i][half r], HF wants [head][half r][pair i]"""
    r = w.shape[0]
    return w.reshape(n_head, r // n_head // 2, 2, *w.shape[1:]).transpose(1, 2).reshape(w.shape)


def permute_qk(w: torch.Tensor, n_head: int) -> torch.Tensor:
    """what the llama.cpp converter applies (writer side / tests)"""
    r = w.shape[0]
    return w.reshape(n_head, 2, r // n_head // 2, *w.shape[1:]).transpose(1, 2).reshape(w.shape)


def config_from_gguf(g: GGUFFile) -> Dict[str, Any]:
    """HF-style config dict from GGUF metadata (keys: ``<arch>.block_count`` ...)"""
    a, md = g.architecture, g.metadata
    get = lambda k, d=None: md.get(f"{a}.{k}", d)
    heads = int(get("attention.head_count"))
    hidden = int(get("embedding_length"))
    vocab = get("vocab_size") or (len(md["tokenizer.ggml.tokens"]) if "tokenizer.ggml.tokens" in md else g.tensors["token_embd.weight"][0][0])
    cfg = {"architectures": [_ARCH.get(a, "LlamaForCausalLM")], "num_hidden_layers": int(get("block_count")), "hidden_size": hidden,
           "intermediate_size": int(get("feed_forward_length", 4 * hidden)), "num_attention_heads": heads,
           "num_key_value_heads": int(get("attention.head_count_kv", heads)), "vocab_size": int(vocab),
           "rms_norm_eps": float(get("attention.layer_norm_rms_epsilon", get("attention.layer_norm_epsilon", 1e-5))),
           "rope_theta": float(get("rope.freq_base", 10000.0)), "max_position_embeddings": int(get("context_length", 4096)),
           "tie_word_embeddings": "output.weight" not in g.tensors, "torch_dtype": "bfloat16"}
    if get("attention.key_length"):
        cfg["head_dim"] = int(get("attention.key_length"))
    if get("expert_count"):
        cfg["num_local_experts"] = cfg["num_experts"] = int(get("expert_count"))
        cfg["num_experts_per_tok"] = int(get("expert_used_count", 2))
        if get("expert_feed_forward_length"):
            cfg["moe_intermediate_size"] = int(get("expert_feed_forward_length"))
    if get("attention.sliding_window"):
        cfg["sliding_window"] = int(get("attention.sliding_window"))
    for k_src, k_dst in (("tokenizer.ggml.bos_token_id", "bos_token_id"), ("tokenizer.ggml.eos_token_id", "eos_token_id")):
        if k_src in md:
            cfg[k_dst] = int(md[k_src])
    return cfg


def iter_gguf(path: str, dtype: Optional[torch.dtype] = None) -> Iterator[Tuple[str, torch.Tensor]]:
    """(canonical HF-style name, tensor) pairs; stacked expert tensors are split per expert"""
    g = GGUFFile(path)
    a = g.architecture
    heads = int(g.metadata.get(f"{a}.attention.head_count", 1))
    kv_heads = int(g.metadata.get(f"{a}.attention.head_count_kv", heads))
    for name in g.tensors:
        t = g.get(name)
        if dtype is not None and t.dtype != dtype:
            t = t.to(dtype)
        base, _, suffix = name.rpartition(".")
        if base in _TOP:
            yield f"{_TOP[base]}.{suffix}", t
            continue
        parts = base.split(".")
        if len(parts) == 3 and parts[0] == "blk":
            i, key = parts[1], parts[2]
            if key in _EXPS:
                for e in range(t.shape[0]):
                    yield f"model.layers.{i}.mlp.experts.{e}.{_EXPS[key]}.{suffix}", t[e]
                continue
            if key in _BLK:
                if a in _PERMUTED_QK and suffix in ("weight", "bias") and key in ("attn_q", "attn_k"):
                    t = _unpermute(t, heads if key == "attn_q" else kv_heads)
                yield f"model.layers.{i}.{_BLK[key]}.{suffix}", t
                continue
        if base in ("rope_freqs",):
            continue
        yield name, t


class GGUFTokenizer:
    """tokenizer rebuilt from ``tokenizer.ggml.*`` metadata with the ``tokenizers`` library (BPE for ``gpt2``-style vocabularies,
    Unigram + byte fallback for SentencePiece ``llama``-style ones); exposes the subset of the HF interface the engine uses"""

    def __init__(self, md: Dict[str, Any]):
        from tokenizers import Tokenizer, decoders, models, pre_tokenizers

        toks: List[str] = md["tokenizer.ggml.tokens"]
        self.vocab_size = len(toks)
        self.bos_token_id = md.get("tokenizer.ggml.bos_token_id")
        self.eos_token_id = md.get("tokenizer.ggml.eos_token_id")
        self.add_bos = bool(md.get("tokenizer.ggml.add_bos_token", md.get("tokenizer.ggml.model", "llama") == "llama"))
        self.chat_template = md.get("tokenizer.chat_template")
        kind = md.get("tokenizer.ggml.model", "llama")
        if kind == "gpt2":
            merges = [tuple(m.split(" ", 1)) for m in md.get("tokenizer.ggml.merges", [])]
            tk = Tokenizer(models.BPE(vocab={t: i for i, t in enumerate(toks)}, merges=merges))
            tk.pre_tokenizer = pre_tokenizers.ByteLevel(add_prefix_space=False)
            tk.decoder = decoders.ByteLevel()
        else:
            scores = md.get("tokenizer.ggml.scores") or [0.0] * len(toks)
            unk = md.get("tokenizer.ggml.unknown_token_id", 0)
            tk = Tokenizer(models.Unigram([(t, float(s)) for t, s in zip(toks, scores)], unk_id=unk, byte_fallback=True))
            tk.pre_tokenizer = pre_tokenizers.Metaspace(replacement="▁", prepend_scheme="first")
            tk.decoder = decoders.Sequence([decoders.Replace("▁", " "), decoders.ByteFallback(), decoders.Fuse(), decoders.Strip(" ", 1, 0)])
        self._tk = tk
        self._toks = toks

    def encode(self, text: str, add_special_tokens: bool = True) -> List[int]:
        ids = self._tk.encode(text, add_special_tokens=False).ids
        if add_special_tokens and self.add_bos and self.bos_token_id is not None:
            ids = [int(self.bos_token_id)] + ids
        return ids

    def decode(self, ids, skip_special_tokens: bool = True, **kw) -> str:
        ids = [int(i) for i in ids]
        if skip_special_tokens:
            ids = [i for i in ids if i not in (self.bos_token_id, self.eos_token_id)]
        return self._tk.decode(ids, skip_special_tokens=False)

    def convert_ids_to_tokens(self, ids):
        return [self._toks[int(i)] for i in ids]

    def __call__(self, text, **kw):
        if isinstance(text, str):
            return {"input_ids": self.encode(text)}
        return {"input_ids": [self.encode(t) for t in text]}

    def __len__(self) -> int:
        return self.vocab_size


def is_gguf(path: Optional[str]) -> bool:
    return bool(path) and os.path.isfile(path) and path.endswith(".gguf")
