"""Tokenizer access.  HF tokenizers when a local path exists; otherwise a deterministic byte-level tokenizer so the
whole stack (chat templates, stop strings, incremental detokenisation, grammar masks) runs offline with
``dummy:`` models.  Reference: ``python/sglang/srt/utils/hf_transformers_utils.py`` ``get_tokenizer``.
"""
from __future__ import annotations

import os
from typing import Any, Dict, List, Optional, Sequence


class ByteTokenizer:
    """ids 0..255 = raw bytes; 256 = <bos>, 257 = <eos>, 258 = <pad>, 259 = <image>; higher ids decode to ''.
    For vocabularies smaller than 260 the specials are folded to the top of the range."""

    def __init__(self, vocab_size: int = 512, eos_token_id: Optional[int] = None, bos_token_id: Optional[int] = None):
        self.vocab_size = vocab_size
        base = min(256, vocab_size - 4)
        self._nbytes = base
        self.bos_token_id = bos_token_id if bos_token_id is not None else base
        self.eos_token_id = eos_token_id if eos_token_id is not None else base + 1
        self.pad_token_id = base + 2
        self.image_token_id = base + 3
        self.bos_token, self.eos_token, self.pad_token = "<bos>", "<eos>", "<pad>"
        self.chat_template = None
        self.name_or_path = "byte"
        self.all_special_ids = [self.bos_token_id, self.eos_token_id, self.pad_token_id, self.image_token_id]

    def __len__(self) -> int:
        return self.vocab_size

    def encode(self, text: str, add_special_tokens: bool = False) -> List[int]:
        sp = set(self.all_special_ids)
        ids = [b % self._nbytes for b in text.encode("utf-8")]
        ids = [i if i not in sp else 63 for i in ids]   # a byte that collides with a special id becomes '?'
        return ([self.bos_token_id] + ids) if add_special_tokens else ids

    def __call__(self, text, add_special_tokens: bool = False, **kw) -> Dict[str, Any]:
        if isinstance(text, str):
            return {"input_ids": self.encode(text, add_special_tokens)}
        return {"input_ids": [self.encode(t, add_special_tokens) for t in text]}

    def decode(self, ids: Sequence[int], skip_special_tokens: bool = True, **kw) -> str:
        out = bytearray()
        sp = set(self.all_special_ids)
        for i in ids:
            i = int(i)
            if i < self._nbytes and i not in sp:
                out.append(i)
            elif not skip_special_tokens:
                name = {self.bos_token_id: "<bos>", self.eos_token_id: "<eos>", self.pad_token_id: "<pad>",
                        self.image_token_id: "<image>"}.get(i, "")
                out.extend(name.encode())
        return out.decode("utf-8", errors="replace")

    def batch_decode(self, batch, **kw) -> List[str]:
        return [self.decode(x, **kw) for x in batch]

    def convert_ids_to_tokens(self, ids):
        if isinstance(ids, int):
            return self.decode([ids], skip_special_tokens=False)
        return [self.decode([i], skip_special_tokens=False) for i in ids]

    def get_vocab(self) -> Dict[str, int]:
        v = {}
        for i in range(self.vocab_size):
            if i < self._nbytes:
                v[bytes([i]).decode("latin-1")] = i
            else:
                v[f"<extra_{i}>"] = i
        return v

    def apply_chat_template(self, messages, tokenize: bool = False, add_generation_prompt: bool = True, **kw):
        text = "".join(f"<|{m['role']}|>\n{_content_text(m.get('content'))}\n" for m in messages)
        if add_generation_prompt:
            text += "<|assistant|>\n"
        return self.encode(text) if tokenize else text


def _content_text(c: Any) -> str:
    if c is None:
        return ""
    if isinstance(c, str):
        return c
    parts = []
    for p in c:
        if isinstance(p, dict):
            if p.get("type") == "text":
                parts.append(p.get("text", ""))
            elif p.get("type") in ("image_url", "image"):
                parts.append("<image>")
        else:
            parts.append(str(p))
    return "".join(parts)


def get_tokenizer(path: Optional[str], vocab_size: int = 512, eos_token_id: Any = None, bos_token_id: Optional[int] = None,
                  trust_remote_code: bool = False):
    if path and not path.startswith("dummy:") and os.path.isdir(path):
        from transformers import AutoTokenizer

        return AutoTokenizer.from_pretrained(path, trust_remote_code=trust_remote_code, local_files_only=True)
    if path and os.path.isfile(path) and path.endswith(".gguf"):
        from ..model_loader.gguf import GGUFFile, GGUFTokenizer

        md = GGUFFile(path).metadata
        if "tokenizer.ggml.tokens" in md:
            return GGUFTokenizer(md)
    eos = eos_token_id[0] if isinstance(eos_token_id, (list, tuple)) else eos_token_id
    return ByteTokenizer(vocab_size, eos, bos_token_id)
