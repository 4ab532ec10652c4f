"""Text -> image / video pipeline and its request loop (reference: ``multimodal_gen/runtime/pipelines_core`` stages —
text encode, latent preparation, denoise with classifier-free guidance, decode — and ``runtime/managers`` request loop)."""
from __future__ import annotations

import base64
import io
import json
import os
import queue
import threading
import time
from dataclasses import dataclass, field
from typing import List, Optional

import torch

from .dit import DiT, DiTConfig, LatentDecoder, TextEncoder
from .schedulers import TeaCache, make_scheduler
from .usp import UspGroup

_PRESETS = {"tiny-dit": DiTConfig(), "tiny-dit-eps": DiTConfig(prediction="epsilon"),
            "dit-xl-2": DiTConfig(hidden_size=1152, num_heads=16, num_layers=28, text_len=64, text_layers=4, vae_channels=128),
            "mmdit-2b": DiTConfig(hidden_size=1536, num_heads=24, num_layers=24, in_channels=16, text_len=128, text_layers=6, vae_channels=128)}


@dataclass
class GenArgs:
    height: int = 64
    width: int = 64
    num_frames: int = 1
    num_inference_steps: int = 8
    guidance_scale: float = 4.0
    negative_prompt: str = ""
    seed: int = 0
    shift: float = 1.0
    teacache_thresh: float = 0.0
    vae_tile: int = 0


@dataclass
class DiffusionRequest:
    prompt: str
    args: GenArgs = field(default_factory=GenArgs)
    rid: str = ""
    result: Optional[torch.Tensor] = None           # uint8 [F, H, W, 3]
    error: Optional[str] = None
    done: threading.Event = field(default_factory=threading.Event)
    timings: dict = field(default_factory=dict)


class DiffusionPipeline:
    def __init__(self, model_path: str = "dummy:tiny-dit", device: str = "cpu", dtype: torch.dtype = torch.float32,
                 sp: Optional[UspGroup] = None, seed: int = 0):
        self.device, self.dtype, self.sp = torch.device(device), dtype, sp
        if model_path.startswith("dummy:"):
            self.cfg = _PRESETS[model_path.split(":", 1)[1]]
            g = torch.Generator().manual_seed(seed)
            self.text, self.dit, self.vae = TextEncoder(self.cfg), DiT(self.cfg), LatentDecoder(self.cfg)
            for m in (self.text, self.dit, self.vae):                        # name-independent but rank-independent init
                for p in m.parameters():
                    with torch.no_grad():
                        p.copy_(torch.randn(p.shape, generator=g) * (0.02 if p.ndim > 1 else 0.0) + (1.0 if p.ndim == 1 and "norm" in str(type(p)) else 0.0))
            for m in (self.text, self.dit, self.vae):
                for mod in m.modules():
                    if isinstance(mod, (torch.nn.LayerNorm, torch.nn.GroupNorm, torch.nn.RMSNorm)) and getattr(mod, "weight", None) is not None:
                        torch.nn.init.ones_(mod.weight)
        else:
            self.cfg = DiTConfig(**json.load(open(os.path.join(model_path, "model_index.json")))["config"])
            self.text, self.dit, self.vae = TextEncoder(self.cfg), DiT(self.cfg), LatentDecoder(self.cfg)
            from safetensors.torch import load_file

            for name, m in (("text_encoder", self.text), ("transformer", self.dit), ("vae", self.vae)):
                m.load_state_dict(load_file(os.path.join(model_path, name, "model.safetensors")))
        for m in (self.text, self.dit, self.vae):
            m.to(self.device, self.dtype).eval()

    def save_pretrained(self, path: str) -> None:
        from safetensors.torch import save_file

        os.makedirs(path, exist_ok=True)
        json.dump({"_class_name": "DiffusionPipeline", "config": self.cfg.to_dict()}, open(os.path.join(path, "model_index.json"), "w"))
        for name, m in (("text_encoder", self.text), ("transformer", self.dit), ("vae", self.vae)):
            os.makedirs(os.path.join(path, name), exist_ok=True)
            save_file({k: v.contiguous().cpu() for k, v in m.state_dict().items()}, os.path.join(path, name, "model.safetensors"))

    # -- stages -----------------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def encode_prompt(self, prompts: List[str]):
        ids = self.text.tokenize(prompts).to(self.device)
        return self.text(ids)

    def prepare_latents(self, a: GenArgs, batch: int = 1) -> torch.Tensor:
        c, s = self.cfg, self.cfg.vae_scale
        assert a.height % (s * c.patch_size) == 0 and a.width % (s * c.patch_size) == 0, f"height/width must be multiples of {s * c.patch_size}"
        g = torch.Generator().manual_seed(a.seed)                         # CPU generator: identical on every SP rank and device
        z = torch.randn(batch, c.in_channels, a.num_frames, a.height // s, a.width // s, generator=g)
        return z.to(self.device, self.dtype)

    @torch.no_grad()
    def denoise(self, z, txt, pooled, utxt, upooled, a: GenArgs, timings: Optional[dict] = None):
        sch = make_scheduler(self.cfg.prediction, a.shift)
        sch.set_timesteps(a.num_inference_steps)
        z = z * sch.init_noise_sigma()
        cfg_on = a.guidance_scale > 1.0
        cache = TeaCache(a.teacache_thresh)
        B = z.shape[0]
        for i, t in enumerate(sch.timesteps):
            tb = torch.full((B * (2 if cfg_on else 1),), float(t), device=self.device)
            sig = self.dit.t_mlp[0].weight[:, :8].float().sum() * 0 + torch.cat([z.float().flatten()[:4096], tb[:1] / 1000])  # cheap input signature
            if cache.should_skip(sig, i, i == len(sch.timesteps) - 1):
                out = cache.residual
            else:
                if cfg_on:
                    o = self.dit(torch.cat([z, z]), tb, torch.cat([txt, utxt]), torch.cat([pooled, upooled]), self.sp)
                    oc, ou = o.chunk(2)
                    out = ou + a.guidance_scale * (oc - ou)
                else:
                    out = self.dit(z, tb, txt, pooled, self.sp)
                cache.residual = out
            z = sch.step(out, i, z)
        if timings is not None:
            timings["skipped_steps"] = list(cache.skipped)
        return z

    @torch.no_grad()
    def decode(self, z: torch.Tensor, tile: int = 0) -> torch.Tensor:
        B, C, Fr, H, W = z.shape
        img = self.vae(z.permute(0, 2, 1, 3, 4).reshape(B * Fr, C, H, W), tile)
        img = ((img.float().clamp(-1, 1) + 1) * 127.5).round().to(torch.uint8)
        return img.reshape(B, Fr, 3, img.shape[-2], img.shape[-1]).permute(0, 1, 3, 4, 2).cpu()

    def generate(self, prompt, args: Optional[GenArgs] = None, timings: Optional[dict] = None, **kw) -> torch.Tensor:
        """prompt: str or list of str.  Returns uint8 [B, F, H, W, 3] (B dropped for a single prompt)."""
        a = args or GenArgs(**kw)
        prompts = [prompt] if isinstance(prompt, str) else list(prompt)
        t0 = time.perf_counter()
        txt, pooled = self.encode_prompt(prompts)
        utxt, upooled = self.encode_prompt([a.negative_prompt] * len(prompts))
        t1 = time.perf_counter()
        z = self.denoise(self.prepare_latents(a, len(prompts)), txt, pooled, utxt, upooled, a, timings)
        t2 = time.perf_counter()
        out = self.decode(z, a.vae_tile)
        if timings is not None:
            timings.update(encode_s=t1 - t0, denoise_s=t2 - t1, decode_s=time.perf_counter() - t2)
        return out[0] if isinstance(prompt, str) else out


def to_png(frame: torch.Tensor) -> bytes:
    from PIL import Image

    buf = io.BytesIO()
    Image.fromarray(frame.numpy()).save(buf, format="PNG")
    return buf.getvalue()


def to_gif(frames: torch.Tensor, fps: int = 8) -> bytes:
    from PIL import Image

    ims = [Image.fromarray(f.numpy()) for f in frames]
    buf = io.BytesIO()
    ims[0].save(buf, format="GIF", save_all=True, append_images=ims[1:], duration=int(1000 / fps), loop=0)
    return buf.getvalue()


class DiffusionEngine:
    """Request loop: one worker thread owns the pipeline; requests with identical shapes / step counts / guidance are
    batched into one denoise loop (the reference's scheduler batches the same way: same-shape requests only)."""

    def __init__(self, pipeline: DiffusionPipeline, max_batch: int = 4):
        self.pipe, self.max_batch = pipeline, max_batch
        self.q: "queue.Queue[Optional[DiffusionRequest]]" = queue.Queue()
        self.num_served = 0
        self._th = threading.Thread(target=self._loop, daemon=True)
        self._th.start()

    def submit(self, prompt: str, **kw) -> DiffusionRequest:
        r = DiffusionRequest(prompt, GenArgs(**kw), rid=f"img-{time.monotonic_ns():x}")
        self.q.put(r)
        return r

    def generate(self, prompt: str, timeout: float = 600.0, **kw) -> torch.Tensor:
        r = self.submit(prompt, **kw)
        if not r.done.wait(timeout):
            raise TimeoutError("diffusion request timed out")
        if r.error:
            raise RuntimeError(r.error)
        return r.result

    def _key(self, r: DiffusionRequest):
        a = r.args
        return (a.height, a.width, a.num_frames, a.num_inference_steps, a.guidance_scale, a.negative_prompt, a.shift, a.teacache_thresh, a.seed)

    def _loop(self) -> None:
        hold: List[DiffusionRequest] = []
        while True:
            r = hold.pop(0) if hold else self.q.get()
            if r is None:
                return
            batch = [r]
            while len(batch) < self.max_batch:
                try:
                    n = self.q.get_nowait()
                except queue.Empty:
                    break
                if n is None:
                    self.q.put(None)
                    break
                (batch if self._key(n) == self._key(r) else hold).append(n)
            try:
                tm: dict = {}
                out = self.pipe.generate([b.prompt for b in batch], r.args, tm)
                for b, o in zip(batch, out):
                    b.result, b.timings = o, dict(tm, batch=len(batch))
            except Exception as e:  # noqa: BLE001 — reported to the caller
                for b in batch:
                    b.error = f"{type(e).__name__}: {e}"
            self.num_served += len(batch)
            for b in batch:
                b.done.set()

    def shutdown(self) -> None:
        self.q.put(None)
        self._th.join(timeout=5)


def build_app(engine: DiffusionEngine):
    """OpenAI-style image / video endpoints (reference ``multimodal_gen/runtime/entrypoints/openai``)."""
    import asyncio

    from fastapi import FastAPI
    from fastapi.responses import JSONResponse

    app = FastAPI()

    async def _run(body: dict, video: bool):
        size = str(body.get("size", "64x64")).lower().split("x")
        kw = dict(width=int(size[0]), height=int(size[1]), num_inference_steps=int(body.get("num_inference_steps", 8)),
                  guidance_scale=float(body.get("guidance_scale", 4.0)), negative_prompt=body.get("negative_prompt", ""),
                  num_frames=int(body.get("num_frames", 8 if video else 1)), teacache_thresh=float(body.get("teacache_thresh", 0.0)))
        n, seed = int(body.get("n", 1)), int(body.get("seed", 0))
        reqs = [engine.submit(body["prompt"], seed=seed + i, **kw) for i in range(n)]
        loop = asyncio.get_running_loop()
        for r in reqs:
            await loop.run_in_executor(None, r.done.wait, 600.0)
        if any(r.error for r in reqs):
            return JSONResponse({"error": {"message": next(r.error for r in reqs if r.error), "type": "invalid_request_error"}}, status_code=400)
        enc = (lambda r: to_gif(r.result)) if video else (lambda r: to_png(r.result[0]))
        return {"created": int(time.time()), "data": [{"b64_json": base64.b64encode(enc(r)).decode(), "timings": r.timings} for r in reqs]}

    @app.post("/v1/images/generations")
    async def images(body: dict):
        return await _run(body, False)

    @app.post("/v1/videos/generations")
    async def videos(body: dict):
        return await _run(body, True)

    @app.get("/health")
    async def health():
        return {"status": "ok", "served": engine.num_served}

    @app.get("/v1/models")
    async def models():
        return {"object": "list", "data": [{"id": "diffusion", "object": "model"}]}

    return app


def main(argv=None) -> None:
    """``python -m sglang_b200.multimodal_gen.pipeline --model-path dummy:tiny-dit [--port P | --prompt ... --output x.png]``.
    Under ``torchrun`` with ``--ulysses-degree/--ring-degree`` every rank runs the same SPMD loop (rank 0 serves / writes)."""
    import argparse

    ap = argparse.ArgumentParser("sglang_b200 diffusion")
    ap.add_argument("--model-path", default="dummy:tiny-dit")
    ap.add_argument("--device", default="cuda" if torch.cuda.is_available() else "cpu")
    ap.add_argument("--dtype", default="float32")
    ap.add_argument("--ulysses-degree", type=int, default=1)
    ap.add_argument("--ring-degree", type=int, default=1)
    ap.add_argument("--port", type=int, default=0)
    ap.add_argument("--host", default="127.0.0.1")
    ap.add_argument("--prompt")
    ap.add_argument("--output", default="out.png")
    for f, t, d in (("height", int, 64), ("width", int, 64), ("num-frames", int, 1), ("num-inference-steps", int, 8),
                    ("guidance-scale", float, 4.0), ("seed", int, 0), ("teacache-thresh", float, 0.0)):
        ap.add_argument("--" + f, type=t, default=d)
    ns = ap.parse_args(argv)
    sp = None
    if ns.ulysses_degree * ns.ring_degree > 1:
        import torch.distributed as dist

        if not dist.is_initialized():
            dist.init_process_group("nccl" if ns.device.startswith("cuda") else "gloo")
        if ns.device.startswith("cuda"):
            torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
        sp = UspGroup.build(ns.ulysses_degree, ns.ring_degree)
    pipe = DiffusionPipeline(ns.model_path, ns.device, getattr(torch, ns.dtype), sp)
    if ns.prompt is not None:
        out = pipe.generate(ns.prompt, GenArgs(ns.height, ns.width, ns.num_frames, ns.num_inference_steps, ns.guidance_scale, "", ns.seed,
                                               teacache_thresh=ns.teacache_thresh))
        if sp is None or sp.sp_rank == 0:
            open(ns.output, "wb").write(to_gif(out) if ns.num_frames > 1 else to_png(out[0]))
            print(json.dumps({"output": ns.output, "shape": list(out.shape)}))
        return
    import uvicorn

    uvicorn.run(build_app(DiffusionEngine(pipe)), host=ns.host, port=ns.port or 30010, log_level="warning")


if __name__ == "__main__":
    main()
