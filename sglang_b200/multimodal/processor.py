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
from typing import Any, Dict, List, Tuple

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


def squash_preprocess_batch(imgs: List[np.ndarray], size: int, mean, std, filt: int = 2) -> np.ndarray:
    """[N, 3, size, size] float32: resize straight to size x size (filter 2 = bilinear, 3 = bicubic), rescale, normalise"""
    rt = _native()
    mean, std = np.asarray(mean, np.float32), np.asarray(std, np.float32)
    out = []
    for a in imgs:
        if rt is not None:
            r = rt.image_resize(np.ascontiguousarray(a), size, size, filt)
        else:
            from PIL import Image

            r = np.asarray(Image.fromarray(a).resize((size, size), Image.BILINEAR if filt == 2 else Image.BICUBIC))
        out.append(((r.astype(np.float32) / 255.0 - mean) / std).transpose(2, 0, 1))
    return np.stack(out) if out else np.zeros((0, 3, size, size), np.float32)


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


def smart_resize(h: int, w: int, factor: int = 28, min_pixels: int = 56 * 56, max_pixels: int = 14 * 14 * 4 * 1280) -> Tuple[int, int]:
    """Qwen2-VL dynamic resolution: both sides multiples of ``factor``, area inside [min_pixels, max_pixels], aspect ratio kept"""
    import math

    if max(h, w) / min(h, w) > 200:
        raise ValueError("image aspect ratio must be below 200")
    hb, wb = max(factor, round(h / factor) * factor), max(factor, round(w / factor) * factor)
    if hb * wb > max_pixels:
        beta = math.sqrt((h * w) / max_pixels)
        hb, wb = math.floor(h / beta / factor) * factor, math.floor(w / beta / factor) * factor
    elif hb * wb < min_pixels:
        beta = math.sqrt(min_pixels / (h * w))
        hb, wb = math.ceil(h * beta / factor) * factor, math.ceil(w * beta / factor) * factor
    return hb, wb


def qwen_vl_patches(img: np.ndarray, patch: int, merge: int, temporal: int, min_pixels: int, max_pixels: int, mean=None,
                    std=None) -> Tuple[np.ndarray, Tuple[int, int, int]]:
    """uint8 HWC -> (flattened patches [gh*gw, 3*t*p*p] in merge-block order, (1, gh, gw))"""
    h, w = img.shape[:2]
    nh, nw = smart_resize(h, w, patch * merge, min_pixels, max_pixels)
    rt = _native()
    if rt is not None:
        r = rt.image_resize(np.ascontiguousarray(img), nh, nw, 3)
    else:
        from PIL import Image

        r = np.asarray(Image.fromarray(img).resize((nw, nh), Image.BICUBIC))
    mean = CLIP_MEAN if mean is None else np.asarray(mean, np.float32)
    std = CLIP_STD if std is None else np.asarray(std, np.float32)
    chw = ((r.astype(np.float32) / 255.0 - mean) / std).transpose(2, 0, 1)
    return patchify(np.ascontiguousarray(chw), patch, merge, temporal), (1, nh // patch, nw // patch)


def process_images_qwen_vl(input_ids: List[int], image_data: Any, cfg: Any, tokenizer: Any, encoder: Any = None):
    """Qwen2-VL: every image placeholder token expands to ``gh * gw / merge^2`` pad tokens (hash-derived ids for the radix cache); returns the
    patches, grids, M-RoPE position table and decode delta (reference multimodal/processors/qwen_vl.py)"""
    from ..models.qwen2_vl import mrope_index

    v = cfg.vision
    patch, merge, tps = v["patch_size"], v["spatial_merge_size"], v.get("temporal_patch_size", 2)
    imgs = image_data if isinstance(image_data, (list, tuple)) else [image_data]
    arrs = [load_image(x) for x in imgs]
    pp = [qwen_vl_patches(a, patch, merge, tps, v.get("min_pixels", 56 * 56), v.get("max_pixels", 14 * 14 * 4 * 1280), v.get("image_mean"),
                          v.get("image_std")) for a in arrs]
    img_tok, V = cfg.image_token_id, cfg.vocab_size
    out, offsets, counts, pads, k = [], [], [], [], 0
    for t in input_ids:
        if t == img_tok and k < len(arrs):
            g = pp[k][1]
            n = g[0] * g[1] * g[2] // (merge * merge)
            h = image_hash(arrs[k])
            offsets.append(len(out))
            counts.append(n)
            pads.append(int(h % max(1, V - 1)))
            out.extend([img_tok] * n)
            k += 1
        else:
            out.append(t)
    if k < len(arrs):
        raise ValueError(f"{len(arrs)} images but only {k} image placeholder tokens in the prompt")
    pos, delta = mrope_index(out, [g for _, g in pp], img_tok, merge)
    # the placeholder ids themselves never reach the model (their embedding rows are overwritten): replace them with hash-derived ids so
    # that different images never share a radix-cache key
    for off, n, k_ in zip(offsets, counts, range(len(arrs))):
        h = image_hash(arrs[k_])
        out[off:off + n] = [pads[k_] if (j & 1) == 0 else int((h >> 17) % max(1, V - 1)) for j in range(n)]
    mm: Dict[str, Any] = {"image_offsets": offsets, "image_token_counts": counts, "pad_values": pads, "num_image_tokens": 0,
                          "grid_thw": [list(g) for _, g in pp], "mrope_positions": pos, "mrope_delta": int(delta)}
    patches = np.concatenate([p for p, _ in pp]) if pp else np.zeros((0, 3 * tps * patch * patch), np.float32)
    if encoder is not None:       # EPD: one item per image (its patches + grid), embeddings come back per image
        embs = encoder.encode_items([{"hash": image_hash(a), "array": p_, "grid": list(g_)} for a, (p_, g_) in zip(arrs, pp)])
        mm["image_embeds"] = np.concatenate(embs) if embs else np.zeros((0, cfg.hidden_size), np.float16)
    else:
        mm["pixel_patches"] = patches
    return out, mm


def pixtral_resize_shape(h: int, w: int, longest: int, unit: int) -> Tuple[int, int]:
    """Pixtral / Mistral3 processor: shrink (never enlarge) so that both sides fit ``longest``, then round each side UP to a multiple of
    ``unit`` (= patch * spatial merge)"""
    import math

    ratio = max(h / longest, w / longest)
    if ratio > 1:
        h, w = int(math.floor(h / ratio)), int(math.floor(w / ratio))
    return ((h - 1) // unit + 1) * unit, ((w - 1) // unit + 1) * unit


def process_images_pixtral(input_ids: List[int], image_data: Any, cfg: Any, tokenizer: Any, encoder: Any = None):
    """Mistral3: the single ``[IMG]`` placeholder of each image expands to ``cols`` image tokens per merged row followed by ``[IMG_BREAK]``
    (``[IMG_END]`` after the last row); only the image tokens receive vision features, so every row is its own splice run (reference
    multimodal/processors/pixtral.py)"""
    v = cfg.vision
    patch, merge = v["patch_size"], v["spatial_merge_size"]
    unit = patch * merge
    imgs = image_data if isinstance(image_data, (list, tuple)) else [image_data]
    arrs = [load_image(x) for x in imgs]
    rt = _native()
    pp = []
    for a in arrs:
        nh, nw = pixtral_resize_shape(a.shape[0], a.shape[1], v.get("longest_edge", v["image_size"]), unit)
        if rt is not None:
            r = rt.image_resize(np.ascontiguousarray(a), nh, nw, 3)
        else:
            from PIL import Image

            r = np.asarray(Image.fromarray(a).resize((nw, nh), Image.BICUBIC))
        chw = ((r.astype(np.float32) / 255.0 - CLIP_MEAN) / CLIP_STD).transpose(2, 0, 1)
        pp.append((patchify(np.ascontiguousarray(chw), patch, 1, 1), (1, nh // patch, nw // patch)))
    img_tok, brk, end, V = cfg.image_token_id, v.get("image_break_token_id", 12), v.get("image_end_token_id", 13), cfg.vocab_size
    out, offsets, counts, k = [], [], [], 0
    for t in input_ids:
        if t == img_tok and k < len(arrs):
            _, gh, gw = pp[k][1]
            rows, cols = gh // merge, gw // merge
            h = image_hash(arrs[k])
            pads = (int(h % max(1, V - 1)), int((h >> 17) % max(1, V - 1)))
            for r_ in range(rows):
                offsets.append(len(out))
                counts.append(cols)
                out.extend(pads[(r_ + j) & 1] for j in range(cols))       # hash-derived ids: different images never share a radix key
                out.append(end if r_ == rows - 1 else brk)
            k += 1
        else:
            out.append(t)
    if k < len(arrs):
        raise ValueError(f"{len(arrs)} images but only {k} image placeholder tokens in the prompt")
    mm: Dict[str, Any] = {"image_offsets": offsets, "image_token_counts": counts, "num_image_tokens": 0, "grid_thw": [list(g) for _, g in pp]}
    if encoder is not None:
        embs = encoder.encode_items([{"hash": image_hash(a), "array": p_, "grid": list(g_)} for a, (p_, g_) in zip(arrs, pp)])
        mm["image_embeds"] = np.concatenate(embs) if embs else np.zeros((0, cfg.hidden_size), np.float16)
    else:
        mm["pixel_patches"] = np.concatenate([p for p, _ in pp]) if pp else np.zeros((0, 3 * patch * patch), np.float32)
    return out, mm


IMAGENET_MEAN = np.array([0.485, 0.456, 0.406], dtype=np.float32)
IMAGENET_STD = np.array([0.229, 0.224, 0.225], dtype=np.float32)


def internvl_tiles(img: np.ndarray, size: int = 448, min_tiles: int = 1, max_tiles: int = 12, thumbnail: bool = True) -> np.ndarray:
    """InternVL dynamic tiling: pick the (cols, rows) grid with cols*rows in [min, max] whose aspect ratio is closest to the image's (ties go
    to the larger grid when the image has enough pixels), resize to the grid, cut ``size`` tiles, append a thumbnail -> [n, 3, size, size]"""
    h, w = img.shape[:2]
    ar = w / h
    grids = sorted({(i, j) for n in range(min_tiles, max_tiles + 1) for i in range(1, n + 1) for j in range(1, n + 1) if min_tiles <= i * j <= max_tiles},
                   key=lambda g: g[0] * g[1])
    best, best_d = (1, 1), float("inf")
    for g in grids:
        d = abs(ar - g[0] / g[1])
        if d < best_d or (d == best_d and w * h > 0.5 * size * size * g[0] * g[1]):
            best, best_d = g, d
    rt = _native()

    def resize(a, hh, ww):
        if rt is not None:
            return rt.image_resize(np.ascontiguousarray(a), hh, ww, 3)
        from PIL import Image

        return np.asarray(Image.fromarray(a).resize((ww, hh), Image.BICUBIC))

    big = resize(img, size * best[1], size * best[0])
    tiles = [big[r * size:(r + 1) * size, c * size:(c + 1) * size] for r in range(best[1]) for c in range(best[0])]
    if thumbnail and len(tiles) > 1:
        tiles.append(resize(img, size, size))
    arr = np.stack(tiles).astype(np.float32) / 255.0
    return np.ascontiguousarray(((arr - IMAGENET_MEAN) / IMAGENET_STD).transpose(0, 3, 1, 2))


def process_images_internvl(input_ids: List[int], image_data: Any, cfg: Any, tokenizer: Any, encoder: Any = None):
    v = cfg.vision
    per_tile = int((v["image_size"] // v["patch_size"]) ** 2 * v.get("downsample_ratio", 0.5) ** 2)
    imgs = image_data if isinstance(image_data, (list, tuple)) else [image_data]
    arrs = [load_image(x) for x in imgs]
    tiles = [internvl_tiles(a, v["image_size"], max_tiles=v.get("max_tiles", 12)) for a in arrs]
    img_tok, V = cfg.image_token_id, cfg.vocab_size
    out, offsets, counts, pads, k = [], [], [], [], 0
    for t in input_ids:
        if t == img_tok and k < len(arrs):
            n = tiles[k].shape[0] * per_tile
            h = image_hash(arrs[k])
            offsets.append(len(out))
            counts.append(n)
            pads.append(int(h % max(1, V - 1)))
            out.extend([pads[-1] if (j & 1) == 0 else int((h >> 17) % max(1, V - 1)) for j in range(n)])
            k += 1
        else:
            out.append(t)
    if k < len(arrs):
        raise ValueError(f"{len(arrs)} images but only {k} image placeholder tokens in the prompt")
    mm = {"image_offsets": offsets, "image_token_counts": counts, "pad_values": pads, "num_image_tokens": per_tile}
    if encoder is not None:       # EPD: one item per image = all of its tiles
        embs = encoder.encode_items([{"hash": image_hash(a), "array": t_, "grid": None} for a, t_ in zip(arrs, tiles)])
        mm["image_embeds"] = np.concatenate(embs) if embs else np.zeros((0, cfg.hidden_size), np.float16)
    else:
        mm["pixel_values"] = np.concatenate(tiles) if tiles else np.zeros((0, 3, v["image_size"], v["image_size"]), np.float32)
    return out, mm


def process_images(input_ids: List[int], image_data: Any, cfg: Any, tokenizer: Any, encoder: Any = None) -> Tuple[List[int], Dict[str, Any]]:
    if cfg.architecture.startswith("InternVL"):
        return process_images_internvl(input_ids, image_data, cfg, tokenizer, encoder)
    if cfg.architecture.startswith(("Mistral3", "Pixtral", "LightOnO")):
        return process_images_pixtral(input_ids, image_data, cfg, tokenizer, encoder)
    if cfg.architecture.startswith(("Qwen2VL", "Qwen2_5_VL", "Qwen3VL", "Glm4v")):
        return process_images_qwen_vl(input_ids, image_data, cfg, tokenizer, encoder)
    return _process_images_clip(input_ids, image_data, cfg, tokenizer, encoder)


def _process_images_clip(input_ids: List[int], image_data: Any, cfg: Any, tokenizer: Any, encoder: Any = None) -> Tuple[List[int], Dict[str, Any]]:
    """returns (expanded ids, mm_inputs).  mm_inputs: pixel_values [N,3,S,S] float32 (nested lists for IPC),
    image_offsets (start index of every image's pad run), pad_values, num_image_tokens."""
    v = cfg.vision
    size, patch = v["image_size"], v["patch_size"]
    n_tok = int(v.get("tokens_per_image") or (size // patch) ** 2)
    imgs = image_data if isinstance(image_data, (list, tuple)) else [image_data]
    arrs = [load_image(x) for x in imgs]
    if v.get("resize_mode") == "squash":          # SigLIP / Gemma-3: plain resize to the square input (aspect ratio not kept), own mean / std
        pixels = squash_preprocess_batch(arrs, size, v.get("image_mean", CLIP_MEAN), v.get("image_std", CLIP_STD), int(v.get("resize_filter", 2)))
    else:
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
