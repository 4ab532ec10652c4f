"""SM-partitioned CUDA streams (green contexts).

Parity: reference ``sgl-kernel/csrc/spatial/greenctx_stream.cu`` + ``python/sglang/srt/multiplex/pdmux_context.py`` (PD-multiplexing runs prefill
and decode on disjoint SM groups of one GPU).  ``split_sm_streams`` wraps the native op (``csrc/kernels/green_ctx.cu``): work launched on stream
``i`` is confined to its own group of SMs, so a bandwidth-bound decode step and a compute-bound prefill chunk can share the GPU without the
prefill's CTAs occupying every SM.  B200: 148 SMs, the driver hands out groups in multiples of 8.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch


def split_sm_streams(sm_counts: Sequence[int], device: int = 0) -> List[Tuple["torch.cuda.Stream", int]]:
    """``sm_counts``: SMs requested for every partition but the last (which receives the remainder).  Returns ``[(stream, sms), ...]`` with the
    SM count the driver actually provisioned.  Raises ``RuntimeError`` when the driver has no green-context support."""
    from .. import ops

    if not ops.load_native():
        raise RuntimeError("green contexts need the native extension")
    flat = torch.ops.sglb.green_ctx_streams(int(device), [int(c) for c in sm_counts])
    return [(torch.cuda.ExternalStream(int(flat[i]), device=torch.device("cuda", device)), int(flat[i + 1])) for i in range(0, len(flat), 2)]


def pd_partition(prefill_fraction: float = 0.75, device: int = 0):
    """the PD-multiplex split: (prefill stream, decode stream) with ``prefill_fraction`` of the SMs (rounded to the driver granularity)
    for the compute-bound side"""
    total = torch.cuda.get_device_properties(device).multi_processor_count
    want = max(8, min(total - 8, int(round(total * prefill_fraction / 8)) * 8))
    (ps, pn), (ds, dn) = split_sm_streams([want], device)
    return (ps, pn), (ds, dn)
