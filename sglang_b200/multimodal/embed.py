"""Vision features -> rows of the LLM input embedding (reference ``managers/mm_utils.py:609`` general_mm_embed_routine).
Only image tokens that fall inside the *current* extend range are encoded/spliced (chunked prefill and prefix-cache hits
skip what is already in the KV cache); an encoder-output cache keyed by image hash avoids re-encoding across chunks."""
from __future__ import annotations

from typing import Any, Dict, Optional, Tuple

import numpy as np
import torch

import itertools

_ENC_CACHE: Dict[Tuple[int, int], torch.Tensor] = {}
_ENC_CACHE_MAX = 64
_ENC_SERIAL = itertools.count(1)


def _encode(model: Any, mm: Dict[str, Any], device: torch.device, dtype: torch.dtype):
    """-> list of [tokens_k, hidden] tensors, one per image"""
    counts = mm.get("image_token_counts") or [mm["num_image_tokens"]] * len(mm["image_offsets"])
    if mm.get("image_embeds") is not None:      # EPD: an encoder server already produced the LM-space rows
        e = mm["image_embeds"]
        flat = torch.as_tensor(np.asarray(e), device=device).to(dtype)
        flat = flat.reshape(-1, flat.shape[-1])
    else:
        if not (hasattr(model, "encode_images") or hasattr(model, "encode_audio")) or getattr(model, "vision_tower", None) is None:
            raise RuntimeError("this server was started --language-only: images need --encoder-urls (no local vision tower)")
        if mm.get("audio_features") is not None:    # audio-understanding LLMs: padded log-mel windows + true frame counts
            px = torch.as_tensor(np.asarray(mm["audio_features"], dtype=np.float32), device=device)
            flat = model.encode_audio(px.to(dtype), mm["audio_frames"])
        elif mm.get("pixel_patches") is not None:     # dynamic-resolution towers (Qwen2-VL): flattened patches + grids
            px = torch.as_tensor(np.asarray(mm["pixel_patches"], dtype=np.float32), device=device)
            flat = model.encode_images(px.to(dtype), mm["grid_thw"])
        else:
            px = torch.as_tensor(np.asarray(mm["pixel_values"], dtype=np.float32), device=device)
            flat = model.encode_images(px.to(dtype))
        flat = flat.reshape(-1, flat.shape[-1])
    assert flat.shape[0] == sum(counts), (flat.shape, counts)
    return list(flat.split(counts, 0))


def build_mm_embeds(model: Any, sb: Any, device: torch.device, dtype: torch.dtype) -> Optional[Tuple[torch.Tensor, torch.Tensor]]:
    rows, feats = [], []
    cu = np.concatenate([[0], np.cumsum(sb.q_lens)])
    for i, r in enumerate(sb.reqs):
        mm = r.mm_inputs
        if not mm or not mm.get("image_offsets"):
            continue
        s, e = int(sb.prefix_lens[i]), int(sb.prefix_lens[i] + sb.q_lens[i])
        counts = mm.get("image_token_counts") or [mm["num_image_tokens"]] * len(mm["image_offsets"])
        need = [k for k, off in enumerate(mm["image_offsets"]) if off < e and off + counts[k] > s]
        if not need:
            continue
        # a serial stamped on the request's own dict: ``id(mm)`` can be reused by a later request once this one is collected, which handed
        # that request another image's rows (seen as a size mismatch in the EPD test)
        if "_enc_serial" not in mm:
            mm["_enc_serial"] = next(_ENC_SERIAL)
        key = (mm["_enc_serial"], 0)
        emb = _ENC_CACHE.get(key)
        if emb is None:
            emb = _encode(model, mm, device, dtype)
            if len(_ENC_CACHE) >= _ENC_CACHE_MAX:
                _ENC_CACHE.pop(next(iter(_ENC_CACHE)))
            _ENC_CACHE[key] = emb
        for k in need:
            off = mm["image_offsets"][k]
            a, b = max(off, s), min(off + counts[k], e)
            rows.append(torch.arange(int(cu[i]) + a - s, int(cu[i]) + b - s, device=device))
            feats.append(emb[k][a - off:b - off])
    if not rows:
        return None
    return torch.cat(rows), torch.cat(feats)
