"""Diffusion (image / video generation) runtime — the counterpart of the reference's separate ``multimodal_gen`` product
(``python/sglang/multimodal_gen/``): a DiT denoiser, flow-matching / DDIM schedulers, a latent decoder, and unified sequence
parallelism (Ulysses all-to-all x ring attention) so one long latent sequence is split over GPUs."""
from .pipeline import DiffusionPipeline, DiffusionRequest, GenArgs  # noqa: F401
from .usp import UspGroup, ring_attention, ulysses_attention, usp_attention  # noqa: F401
