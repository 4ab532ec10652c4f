"""Image preprocessing + prompt expansion for LLaVA-style models.

Parity: reference ``python/sglang/srt/multimodal/processors/*`` (base processor: load image from path / URL / base64 /
PIL / ndarray, CLIP preprocessing, expand the single ``<image>`` placeholder to ``num_image_tokens`` pad tokens) and
``managers/schedule_batch.py:220,392`` (pad values derived from the image hash so that the radix cache keys differ
between images and match for identical ones).  The Rust ``sglang-mm`` crate's resize/normalise is the same math
(``rust/sglang-mm``); here it is numpy/torch on the host, batched.
"""
from __future__ import annotations

import base64
import hashlib
import io
import os
from typing import Any, Dict, List, Optional, Tuple

import numpy as np
import torch

CLIP_MEAN = np.array([0.48145466, 0.4578275, 0.40821073], dtype=np.float32)
CLIP_STD = np.array([0.26862954, 0.26130258, 0.27577711], dtype=np.float32)


def load_image(data: Any) -> np.ndarray:
    """-> uint8 HWC RGB"""
    from PIL import Image

    if isinstance(data, np.ndarray):
        arr = data
    elif isinstance(data, torch.Tensor):
        arr = data.cpu().numpy()
    elif isinstance(data, Image.Image):
        arr = np.asarray(data.convert("RGB"))
    elif isinstance(data, bytes):
        arr = np.asarray(Image.open(io.BytesIO(data)).convert("RGB"))
    elif isinstance(data, str):
        if data.startswith("data:"):
            arr = np.asarray(Image.open(io.BytesIO(base64.b64decode(data.split(",", 1)[1]))).convert("RGB"))
        elif os.path.exists(data):
            arr = np.asarray(Image.open(data).convert("RGB"))
        elif data.startswith(("http://", "https://")):
            raise ValueError("remote image URLs are not reachable from this offline deployment; pass a path or base64")
        else:
            arr = np.asarray(Image.open(io.BytesIO(base64.b64decode(data))).convert("RGB"))
    else:
        raise TypeError(f"unsupported image input {type(data)}")
    if arr.ndim == 2:
        arr = np.stack([arr] * 3, -1)
    if arr.dtype != np.uint8:
        arr = (np.clip(arr, 0, 1) * 255).astype(np.uint8) if arr.max() <= 1.0 else arr.astype(np.uint8)
    return arr[..., :3]


def clip_preprocess(img: np.ndarray, size: int) -> np.ndarray:
    """resize shortest side (bicubic) + center crop + normalise -> float32 CHW"""
    from PIL import Image

    im = Image.fromarray(img)
    w, h = im.size
    s = size / min(w, h)
    im = im.resize((max(size, round(w * s)), max(size, round(h * s))), Image.BICUBIC)
    w, h = im.size
    l, t = (w - size) // 2, (h - size) // 2
    arr = np.asarray(im.crop((l, t, l + size, t + size)), dtype=np.float32) / 255.0
    arr = (arr - CLIP_MEAN) / CLIP_STD
    return arr.transpose(2, 0, 1)


def _native():
    """the C++ image pipeline (csrc/runtime/image_ops.cpp: Pillow-exact fixed-point resize, crop+normalise, patchify, hashing), or None"""
    try:
        from .. import sglb_rt

        return sglb_rt if hasattr(sglb_rt, "clip_preprocess_batch") else None
    except Exception:
        return None


def clip_preprocess_batch(imgs: List[np.ndarray], size: int, threads: int = 4) -> np.ndarray:
    """[N, 3, size, size] float32; native thread pool when the runtime is built, Pillow otherwise (bit-identical resize)"""
    rt = _native()
    if rt is not None and imgs:
        return rt.clip_preprocess_batch([np.ascontiguousarray(a) for a in imgs], size, CLIP_MEAN, CLIP_STD, 3, threads)
    return np.stack([clip_preprocess(a, size) for a in imgs])


def patchify(chw: np.ndarray, patch: int, merge: int = 1, temporal: int = 1) -> np.ndarray:
    """[3, H, W] -> [(H/p)(W/p), 3*t*p*p] patch rows in (merge x merge)-block order (Qwen2-VL flattening)"""
    rt = _native()
    if rt is not None:
        return rt.patchify(np.ascontiguousarray(chw, dtype=np.float32), patch, merge, temporal)
    C, H, W = chw.shape
    gh, gw = H // patch, W // patch
    x = np.repeat(chw[:, None], temporal, 1).reshape(C, temporal, gh // merge, merge, patch, gw // merge, merge, patch)
    return np.ascontiguousarray(x.transpose(2, 5, 3, 6, 0, 1, 4, 7)).reshape(gh * gw, C * temporal * patch * patch)


def image_hash(img: np.ndarray) -> int:
    rt = _native()
    if rt is not None:
        return int(rt.hash_bytes(np.ascontiguousarray(img))) & ((1 << 63) - 1)
    return int.from_bytes(hashlib.sha256(img.tobytes()).digest()[:8], "little") & ((1 << 63) - 1)


def process_images(input_ids: List[int], image_data: Any, cfg: Any, tokenizer: Any, encoder: Any = None) -> Tuple[List[int], Dict[str, Any]]:
    """returns (expanded ids, mm_inputs).  mm_inputs: pixel_values [N,3,S,S] float32 (nested lists for IPC),
    image_offsets (start index of every image's pad run), pad_values, num_image_tokens."""
    v = cfg.vision
    size, patch = v["image_size"], v["patch_size"]
    n_tok = (size // patch) ** 2
    imgs = image_data if isinstance(image_data, (list, tuple)) else [image_data]
    arrs = [load_image(x) for x in imgs]
    pixels = clip_preprocess_batch(arrs, size)
    img_tok = cfg.image_token_id if cfg.image_token_id is not None else getattr(tokenizer, "image_token_id", None)
    if img_tok is None:
        raise ValueError("model config has no image_token_id")
    # locate placeholders (a literal "<image>" tokenised as bytes also counts for the byte tokenizer)
    ids = list(input_ids)
    marker = tokenizer.encode("<image>") if tokenizer is not None else []
    pos: List[int] = []
    i = 0
    out: List[int] = []
    k = 0
    offsets, pads = [], []
    V = cfg.vocab_size
    while i < len(ids):
        hit = ids[i] == img_tok or (marker and ids[i:i + len(marker)] == marker)
        if hit and k < len(arrs):
            h = image_hash(arrs[k])
            # pad value: a token id derived from the image hash (kept inside the vocab so embedding lookups stay valid)
            pad = int(h % max(1, V - 1))
            offsets.append(len(out))
            pads.append(pad)
            # the run mixes two hash-derived values so that different images essentially never share a radix key
            run = [pad if (j & 1) == 0 else int((h >> 17) % max(1, V - 1)) for j in range(n_tok)]
            out.extend(run)
            i += 1 if ids[i] == img_tok else len(marker)
            k += 1
        else:
            out.append(ids[i])
            i += 1
    while k < len(arrs):   # images without an explicit placeholder are prepended (reference behaviour for raw prompts)
        h = image_hash(arrs[k])
        pad = int(h % max(1, V - 1))
        run = [pad if (j & 1) == 0 else int((h >> 17) % max(1, V - 1)) for j in range(n_tok)]
        offsets = [0] + [o + n_tok for o in offsets]
        pads = [pad] + pads
        out = run + out
        k += 1
    if encoder is not None:       # EPD: ship the pixel blocks to the encoder fleet now, carry embeddings instead of pixels
        emb = encoder.encode(pixels, [image_hash(a) for a in arrs])
        mm = {"image_embeds": emb, "image_offsets": offsets, "pad_values": pads, "num_image_tokens": n_tok}
    else:
        mm = {"pixel_values": pixels.tolist(), "image_offsets": offsets, "pad_values": pads, "num_image_tokens": n_tok}
    return out, mm
