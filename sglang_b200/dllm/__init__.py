"""Diffusion-LLM (block diffusion) decoding — reference ``python/sglang/srt/dllm`` (``algorithm/low_confidence.py``,
``algorithm/joint_threshold.py``, ``mixin/scheduler.py``; ``ForwardMode.DLLM_EXTEND``).

A decode step emits a whole BLOCK: ``block_size`` positions start as ``mask_id`` and are denoised by repeated forwards in
which the block attends to itself bidirectionally and to the committed prefix; an algorithm decides which positions to
commit after each forward.  The block's KV rows are rewritten by every forward, so the last one (all positions final)
leaves the cache exact.

Design: the worker rides the speculative-decoding seam of the scheduler (a decode step that returns several tokens per
request), and the block-bidirectional pattern is just a custom attention mask on the verify path — no new forward mode,
no model changes; any registered causal LM whose checkpoint was trained for block diffusion can be served this way.
"""
from .algorithms import ALGORITHMS, DllmAlgorithm, JointThreshold, LowConfidence, register_algorithm  # noqa: F401
from .worker import DllmConfig, DllmWorker  # noqa: F401
