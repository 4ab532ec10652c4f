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
    """the vision half of a VLM checkpoint (no language model weights are allocated)"""

    def __init__(self, model_path: str, device: str = "cpu", dtype: Optional[torch.dtype] = None, cache_items: int = 256, seed: int = 0):
        from ..models.llava import CLIPVisionTower

        self.cfg: ModelConfig = load_model_config(model_path)
        if not self.cfg.is_multimodal:
            raise ValueError(f"{model_path}: not a multimodal model, nothing to encode")
        self.device = torch.device(device)
        self.dtype = dtype or {"float32": torch.float32, "float16": torch.float16}.get(self.cfg.dtype, torch.bfloat16)
        v = self.cfg.vision
        self.tower = CLIPVisionTower(v, self.dtype)
        H = self.cfg.hidden_size
        self.projector = torch.nn.Sequential(torch.nn.Linear(v["hidden_size"], H, dtype=self.dtype), torch.nn.GELU(),
                                             torch.nn.Linear(H, H, dtype=self.dtype))
        self.num_image_tokens = self.tower.num_patches
        self._load(model_path, seed)
        self.tower.to(self.device).eval()
        self.projector.to(self.device).eval()
        self._cache: "OrderedDict[int, np.ndarray]" = OrderedDict()
        self._cache_items = cache_items
        self._lock = threading.Lock()
        self.stats = {"encoded": 0, "cache_hits": 0, "requests": 0}

    def _load(self, model_path: str, seed: int) -> None:
        own = {("vision_tower." + n): p for n, p in self.tower.named_parameters()}
        own.update({("projector." + n): p for n, p in self.projector.named_parameters()})
        if model_path.startswith("dummy:"):
            # the same name-seeded synthetic initialisation the full model gets (model_loader.loader.init_dummy_weights)
            from ..model_loader.loader import init_dummy_weights

            holder = torch.nn.Module()
            holder.vision_tower, holder.projector = self.tower, self.projector
            init_dummy_weights(holder, seed, cfg=None)
            return
        from ..model_loader.loader import iter_safetensors

        for name, w in iter_safetensors(model_path):
            if name.startswith(("model.vision_tower", "model.multi_modal_projector")):
                name = name[len("model."):]
            if not (name.startswith("vision_tower") or name.startswith("multi_modal_projector")):
                continue
            n = name.replace("vision_tower.vision_model.embeddings.", "vision_tower.").replace(
                "vision_tower.vision_model.encoder.layers.", "vision_tower.layers.").replace(
                "vision_tower.vision_model.pre_layrnorm", "vision_tower.pre_layrnorm").replace(".self_attn.", ".").replace(
                ".mlp.fc", ".fc").replace("multi_modal_projector.linear_1", "projector.0").replace("multi_modal_projector.linear_2", "projector.2")
            if n in own:
                own[n].data.copy_(w.reshape(own[n].shape))

    @torch.no_grad()
    def encode(self, pixels: np.ndarray, hashes: Optional[Sequence[int]] = None) -> np.ndarray:
        """pixels [N, 3, S, S] float32 -> [N, n_tokens, hidden] float16; rows whose hash is cached are not recomputed"""
        n = pixels.shape[0]
        hashes = list(hashes) if hashes is not None else [None] * n
        out: List[Optional[np.ndarray]] = [None] * n
        with self._lock:
            for i, h in enumerate(hashes):
                if h is not None and h in self._cache:
                    self._cache.move_to_end(h)
                    out[i] = self._cache[h]
                    self.stats["cache_hits"] += 1
        todo = [i for i in range(n) if out[i] is None]
        if todo:
            px = torch.from_numpy(np.ascontiguousarray(pixels[todo])).to(self.device, self.dtype)
            emb = self.projector(self.tower(px)).to(torch.float16).cpu().numpy()
            self.stats["encoded"] += len(todo)
            with self._lock:
                for j, i in enumerate(todo):
                    out[i] = emb[j]
                    if hashes[i] is not None:
                        self._cache[hashes[i]] = emb[j]
                        while len(self._cache) > self._cache_items:
                            self._cache.popitem(last=False)
        self.stats["requests"] += 1
        return np.stack(out)

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
        """body: 4-byte little-endian header length, JSON header {"hashes": [...], "have": [bool...]}, then the .npy pixel block of the
        images with ``have == false`` (in order)"""
        raw = await req.body()
        hl = int.from_bytes(raw[:4], "little")
        hdr = json.loads(raw[4:4 + hl])
        hashes = [int(h) for h in hdr["hashes"]]
        have = hdr.get("have") or [False] * len(hashes)
        sent = _from_npy(raw[4 + hl:]) if len(raw) > 4 + hl else np.zeros((0, 3, 1, 1), np.float32)
        S = sent.shape[-1] if sent.shape[0] else enc.cfg.vision["image_size"]
        pixels = np.zeros((len(hashes), 3, S, S), np.float32)
        k = 0
        for i, hv in enumerate(have):
            if not hv:
                pixels[i] = sent[k]
                k += 1
        if any(have) and not all(c or not hv for c, hv in zip(enc.cached(hashes), have)):
            return Response(status_code=409, content=b"cache miss for a hash the client assumed present")
        return Response(content=_npy(enc.encode(pixels, hashes)), media_type="application/octet-stream")

    return app


class EncoderClient:
    """language-server side: pixel blocks -> embeddings via the encoder fleet"""

    def __init__(self, urls: Sequence[str], timeout: float = 60.0):
        self.urls = [u.rstrip("/") for u in urls]
        self.timeout = timeout
        self.stats = {"requests": 0, "pixels_sent": 0, "pixels_skipped": 0}

    def _post(self, url: str, data: bytes, ctype: str = "application/octet-stream") -> bytes:
        req = urllib.request.Request(url, data=data, headers={"Content-Type": ctype}, method="POST")
        with urllib.request.urlopen(req, timeout=self.timeout) as r:
            return r.read()

    def encode(self, pixels: np.ndarray, hashes: Sequence[int]) -> np.ndarray:
        out: List[Optional[np.ndarray]] = [None] * len(hashes)
        by_url: Dict[str, List[int]] = {}
        for i, h in enumerate(hashes):
            by_url.setdefault(self.urls[int(h) % len(self.urls)], []).append(i)
        for url, idx in by_url.items():
            hs = [int(hashes[i]) for i in idx]
            try:
                have = json.loads(self._post(url + "/encode/probe", json.dumps({"hashes": hs}).encode(), "application/json"))["cached"]
            except Exception:
                have = [False] * len(hs)
            for attempt in range(2):
                hdr = json.dumps({"hashes": hs, "have": have}).encode()
                miss = [i for i, hv in zip(idx, have) if not hv]
                body = len(hdr).to_bytes(4, "little") + hdr + (_npy(np.ascontiguousarray(pixels[miss], dtype=np.float32)) if miss else b"")
                try:
                    emb = _from_npy(self._post(url + "/encode", body))
                    break
                except urllib.error.HTTPError as e:      # evicted between probe and encode: resend everything once
                    if e.code == 409 and attempt == 0:
                        have = [False] * len(hs)
                        continue
                    raise
            self.stats["requests"] += 1
            self.stats["pixels_sent"] += len(miss)
            self.stats["pixels_skipped"] += len(idx) - len(miss)
            for j, i in enumerate(idx):
                out[i] = emb[j]
        return np.stack(out)


def launch_encoder_server(model_path: str, host: str = "127.0.0.1", port: int = 30010, device: str = "auto", **kw) -> None:
    import uvicorn

    dev = ("cuda" if torch.cuda.is_available() else "cpu") if device == "auto" else device
    uvicorn.run(build_encoder_app(VisionEncoder(model_path, dev, **kw)), host=host, port=port, log_level="warning")
