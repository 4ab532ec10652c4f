"""Vision features -> rows of the LLM input embedding (reference ``managers/mm_utils.py:609`` general_mm_embed_routine).
Only image tokens that fall inside the *current* extend range are encoded/spliced (chunked prefill and prefix-cache hits
skip what is already in the KV cache); an encoder-output cache keyed by image hash avoids re-encoding across chunks."""
from __future__ import annotations

from typing import Any, Dict, Optional, Tuple

import numpy as np
import torch

_ENC_CACHE: Dict[Tuple[int, int], torch.Tensor] = {}
_ENC_CACHE_MAX = 64


def build_mm_embeds(model: Any, sb: Any, device: torch.device, dtype: torch.dtype) -> Optional[Tuple[torch.Tensor, torch.Tensor]]:
    rows, feats = [], []
    cu = np.concatenate([[0], np.cumsum(sb.q_lens)])
    for i, r in enumerate(sb.reqs):
        mm = r.mm_inputs
        if not mm:
            continue
        s, e = int(sb.prefix_lens[i]), int(sb.prefix_lens[i] + sb.q_lens[i])
        n_tok = mm["num_image_tokens"]
        need = [k for k, off in enumerate(mm["image_offsets"]) if off < e and off + n_tok > s]
        if not need:
            continue
        key = (id(mm), 0)
        emb = _ENC_CACHE.get(key)
        if emb is None and mm.get("image_embeds") is not None:      # EPD: an encoder server already produced the LM-space rows
            emb = torch.as_tensor(np.asarray(mm["image_embeds"]), device=device).to(dtype).view(len(mm["image_offsets"]), n_tok, -1)
            _ENC_CACHE[key] = emb
        if emb is None:
            if not hasattr(model, "encode_images") or getattr(model, "vision_tower", None) is None:
                raise RuntimeError("this server was started --language-only: images need --encoder-urls (no local vision tower)")
            px = torch.tensor(mm["pixel_values"], dtype=torch.float32, device=device)
            emb = model.encode_images(px.to(dtype)).view(len(mm["image_offsets"]), n_tok, -1)
            if len(_ENC_CACHE) >= _ENC_CACHE_MAX:
                _ENC_CACHE.pop(next(iter(_ENC_CACHE)))
            _ENC_CACHE[key] = emb
        for k in need:
            off = mm["image_offsets"][k]
            a, b = max(off, s), min(off + n_tok, e)
            rows.append(torch.arange(int(cu[i]) + a - s, int(cu[i]) + b - s, device=device))
            feats.append(emb[k, a - off:b - off])
    if not rows:
        return None
    return torch.cat(rows), torch.cat(feats)
