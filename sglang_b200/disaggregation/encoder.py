"""EPD: encoder (vision tower) disaggregation.  Encoder servers run only the vision tower + projector of a VLM and answer with the
language-model-space embeddings of every image; language servers (``--encoder-urls``, optionally ``--language-only`` to skip loading
the vision weights) splice those rows into the prompt exactly where a co-located encoder would have.

Parity: reference ``python/sglang/srt/disaggregation/encoder/*`` (encoder-only server ``launch_server.py:16-52`` dispatch, HTTP/gRPC
embedding transport, MM cache) and ``--encoder-only`` / ``--language-only`` / ``--encoder-urls``.

Transport: one HTTP POST per request with the *preprocessed* pixel tensors (``.npy`` bytes: the resize/crop/normalise work stays on the
language server's CPU thread pool, csrc/runtime/image_ops.cpp) and one ``.npy`` response ``[n_images, n_tokens, hidden]`` in fp16.  Images are
content-addressed (hash of the pixel block): the encoder keeps an LRU of embeddings, the client sends hashes first and pixels only for
misses, and picks the encoder by ``hash % len(urls)`` so repeats land on the server that has them cached.
"""
from __future__ import annotations

import io
import json
import logging
import threading
import urllib.request
from collections import OrderedDict
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

try:      # module level: FastAPI resolves handler annotations against module globals (``from __future__ import annotations``)
    from fastapi import FastAPI, Request, Response
except Exception:      # pragma: no cover
    FastAPI = Request = Response = None

from ..configs.model_config import ModelConfig, load_model_config

logger = logging.getLogger(__name__)


def _npy(a: np.ndarray) -> bytes:
    b = io.BytesIO()
    np.save(b, a, allow_pickle=False)
    return b.getvalue()


def _from_npy(b: bytes) -> np.ndarray:
    return np.load(io.BytesIO(b), allow_pickle=False)


class VisionEncoder:
    """the vision half of a VLM checkpoint: the model class is instantiated with ZERO decoder layers, so any registered multimodal architecture
    (LLaVA, Qwen2-VL / 2.5-VL, InternVL) brings its own tower / projector / checkpoint mapping and no language-model weights are allocated"""

    def __init__(self, model_path: str, device: str = "cpu", dtype: Optional[torch.dtype] = None, cache_items: int = 256, seed: int = 0):
        import dataclasses

        from ..model_loader.loader import build_model, init_dummy_weights, iter_safetensors
        from ..parallel import parallel_state as ps

        self.cfg: ModelConfig = load_model_config(model_path)
        if not self.cfg.is_multimodal:
            raise ValueError(f"{model_path}: not a multimodal model, nothing to encode")
        self.device = torch.device(device)
        cfg0 = dataclasses.replace(self.cfg, num_hidden_layers=0, layer_types=None)
        if dtype is not None:
            cfg0.dtype = {torch.float32: "float32", torch.float16: "float16"}.get(dtype, "bfloat16")
        ps.get_state()
        self.model = build_model(cfg0, self.device, None)
        self.dtype = next(self.model.parameters()).dtype
        if model_path.startswith("dummy:"):
            init_dummy_weights(self.model, seed, cfg=None)      # name-seeded: identical to the tower inside the full dummy model
        else:
            self.model.load_weights((n, w) for n, w in iter_safetensors(model_path) if ".layers." not in n or "vision" in n or "visual" in n)
        self.model.eval()
        self.num_image_tokens = getattr(self.model, "num_image_tokens", 0)
        self._cache: "OrderedDict[int, np.ndarray]" = OrderedDict()
        self._cache_items = cache_items
        self._lock = threading.Lock()
        self.stats = {"encoded": 0, "cache_hits": 0, "requests": 0}

    @torch.no_grad()
    def _encode_one(self, arr: np.ndarray, grid: Optional[Sequence[int]]) -> np.ndarray:
        t = torch.from_numpy(np.array(arr, dtype=np.float32, copy=True)).to(self.device, self.dtype)
        if grid is not None:                      # dynamic-resolution towers: flattened patches + (t, h, w)
            e = self.model.encode_images(t, [list(grid)])
        else:                                     # fixed-size tiles [n, 3, S, S] (LLaVA: n = 1, InternVL: n tiles)
            e = self.model.encode_images(t if t.dim() == 4 else t[None])
        return e.reshape(-1, e.shape[-1]).to(torch.float16).cpu().numpy()

    def encode_items(self, items: Sequence[dict]) -> List[np.ndarray]:
        """items: {"hash": int, "array": ndarray | None (None = expected in the cache), "grid": [t, h, w] | None} -> per-item [tokens, hidden]"""
        out: List[Optional[np.ndarray]] = [None] * len(items)
        with self._lock:
            for i, it in enumerate(items):
                h = it.get("hash")
                if h is not None and h in self._cache:
                    self._cache.move_to_end(h)
                    out[i] = self._cache[h]
                    self.stats["cache_hits"] += 1
        for i, it in enumerate(items):
            if out[i] is not None:
                continue
            if it.get("array") is None:
                raise KeyError(it.get("hash"))
            emb = self._encode_one(it["array"], it.get("grid"))
            self.stats["encoded"] += 1
            out[i] = emb
            if it.get("hash") is not None:
                with self._lock:
                    self._cache[it["hash"]] = emb
                    while len(self._cache) > self._cache_items:
                        self._cache.popitem(last=False)
        self.stats["requests"] += 1
        return out          # type: ignore[return-value]

    def encode(self, pixels: np.ndarray, hashes: Optional[Sequence[int]] = None) -> np.ndarray:
        """fixed-size convenience form: pixels [N, 3, S, S] -> [N, n_tokens, hidden]"""
        hs = list(hashes) if hashes is not None else [None] * pixels.shape[0]
        return np.stack(self.encode_items([{"hash": h, "array": pixels[i], "grid": None} for i, h in enumerate(hs)]))

    def cached(self, hashes: Sequence[int]) -> List[bool]:
        with self._lock:
            return [h in self._cache for h in hashes]


def build_encoder_app(enc: VisionEncoder):
    app = FastAPI()

    @app.get("/health")
    async def health():
        return {"status": "ok", "role": "encoder"}

    @app.get("/get_server_info")
    async def info():
        return {"role": "encoder", "num_image_tokens": enc.num_image_tokens, "hidden_size": enc.cfg.hidden_size, **enc.stats}

    @app.post("/encode/probe")
    async def probe(req: Request):
        body = await req.json()
        return {"cached": enc.cached([int(h) for h in body.get("hashes", [])])}

    @app.post("/encode")
    async def encode(req: Request):
        """body: u32 header length, JSON header {"items": [{"hash", "have", "shape", "grid"}]}, then the raw float32 arrays of the items with
        ``have == false`` back to back.  Reply: u32 header length, JSON {"counts": [...]}, float16 embeddings of all items concatenated."""
        raw = await req.body()
        hl = int.from_bytes(raw[:4], "little")
        hdr = json.loads(raw[4:4 + hl])
        off, items = 4 + hl, []
        for it in hdr["items"]:
            arr = None
            if not it.get("have"):
                n = int(np.prod(it["shape"])) * 4
                arr = np.frombuffer(raw[off:off + n], dtype=np.float32).reshape(it["shape"])
                off += n
            items.append({"hash": int(it["hash"]), "array": arr, "grid": it.get("grid")})
        try:
            embs = enc.encode_items(items)
        except KeyError:
            return Response(status_code=409, content=b"cache miss for a hash the client assumed present")
        out_hdr = json.dumps({"counts": [int(e.shape[0]) for e in embs], "hidden": int(embs[0].shape[1]) if embs else 0}).encode()
        return Response(content=len(out_hdr).to_bytes(4, "little") + out_hdr + b"".join(np.ascontiguousarray(e).tobytes() for e in embs),
                        media_type="application/octet-stream")

    return app


class EncoderClient:
    """language-server side: image arrays -> embeddings via the encoder fleet"""

    def __init__(self, urls: Sequence[str], timeout: float = 60.0):
        self.urls = [u.rstrip("/") for u in urls]
        self.timeout = timeout
        self.stats = {"requests": 0, "pixels_sent": 0, "pixels_skipped": 0}

    def _post(self, url: str, data: bytes, ctype: str = "application/octet-stream") -> bytes:
        req = urllib.request.Request(url, data=data, headers={"Content-Type": ctype}, method="POST")
        with urllib.request.urlopen(req, timeout=self.timeout) as r:
            return r.read()

    def encode_items(self, items: Sequence[dict]) -> List[np.ndarray]:
        """items: {"hash", "array" float32, "grid" | None}; one encoder per item by ``hash % len(urls)`` (repeats meet their cache)"""
        out: List[Optional[np.ndarray]] = [None] * len(items)
        by_url: Dict[str, List[int]] = {}
        for i, it in enumerate(items):
            by_url.setdefault(self.urls[int(it["hash"]) % len(self.urls)], []).append(i)
        for url, idx in by_url.items():
            hs = [int(items[i]["hash"]) for i in idx]
            try:
                have = json.loads(self._post(url + "/encode/probe", json.dumps({"hashes": hs}).encode(), "application/json"))["cached"]
            except Exception:
                have = [False] * len(hs)
            for attempt in range(2):
                meta = [{"hash": hs[j], "have": bool(have[j]), "shape": list(items[i]["array"].shape), "grid": items[i].get("grid")}
                        for j, i in enumerate(idx)]
                hdr = json.dumps({"items": meta}).encode()
                body = len(hdr).to_bytes(4, "little") + hdr + b"".join(
                    np.ascontiguousarray(items[i]["array"], dtype=np.float32).tobytes() for j, i in enumerate(idx) if not have[j])
                try:
                    raw = self._post(url + "/encode", body)
                    break
                except urllib.error.HTTPError as e:      # evicted between probe and encode: resend everything once
                    if e.code == 409 and attempt == 0:
                        have = [False] * len(hs)
                        continue
                    raise
            hl = int.from_bytes(raw[:4], "little")
            rh = json.loads(raw[4:4 + hl])
            flat = np.frombuffer(raw[4 + hl:], dtype=np.float16).reshape(-1, rh["hidden"]) if rh["hidden"] else np.zeros((0, 0), np.float16)
            o = 0
            for j, i in enumerate(idx):
                out[i] = flat[o:o + rh["counts"][j]]
                o += rh["counts"][j]
            self.stats["requests"] += 1
            self.stats["pixels_sent"] += sum(1 for hv in have if not hv)
            self.stats["pixels_skipped"] += sum(1 for hv in have if hv)
        return out          # type: ignore[return-value]

    def encode(self, pixels: np.ndarray, hashes: Sequence[int]) -> np.ndarray:
        return np.stack(self.encode_items([{"hash": h, "array": pixels[i], "grid": None} for i, h in enumerate(hashes)]))


def launch_encoder_server(model_path: str, host: str = "127.0.0.1", port: int = 30010, device: str = "auto", **kw) -> None:
    import uvicorn

    dev = ("cuda" if torch.cuda.is_available() else "cpu") if device == "auto" else device
    uvicorn.run(build_encoder_app(VisionEncoder(model_path, dev, **kw)), host=host, port=port, log_level="warning")
