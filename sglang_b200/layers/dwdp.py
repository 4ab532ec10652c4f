"""DWDP — distributed-weight data parallelism for expert-parallel MoE prefill.

Reference behaviour: ``python/sglang/srt/layers/moe/dwdp/`` (``dwdp_manager.py``, ``weight_manager.py``, ``transport.py``): during large
prefills under DP-attention the TOKENS stay on their rank and the peers' EXPERT WEIGHTS travel instead, prefetched one MoE layer ahead
over NVLink, so a layer costs one weight all-gather (independent of the token count) rather than a dispatch + combine all-to-all whose
volume grows with the batch.

Design here: two ping-pong buffers hold the full expert set of one layer each.  ``weights(i)`` waits for layer i's gather and immediately
posts layer i+1's into the other buffer on a side stream (``async_op`` with gloo on the CPU tier), so the transfer of the next layer
overlaps this layer's grouped GEMMs.  Whether a forward uses DWDP is decided from the DP-wide token count (one tiny all-gather per
forward, cached in the DP-attention context), so every rank takes the same branch and collectives stay matched.
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Tuple

import torch
import torch.distributed as dist


class DwdpManager:
    def __init__(self, layers: List[Any], group: Any, min_tokens: int = 1024):
        assert layers, "DWDP needs MoE layers"
        self.layers, self.group, self.min_tokens = layers, group, int(min_tokens)
        self.index = {id(m): i for i, m in enumerate(layers)}
        m0 = layers[0]
        W = group.world_size
        self.bufs = [(torch.empty((W * m0.w13.shape[0],) + tuple(m0.w13.shape[1:]), dtype=m0.w13.dtype, device=m0.w13.device),
                      torch.empty((W * m0.w2.shape[0],) + tuple(m0.w2.shape[1:]), dtype=m0.w2.dtype, device=m0.w2.device)) for _ in range(2)]
        self.side = torch.cuda.Stream() if m0.w13.is_cuda else None
        self.inflight: Dict[int, Any] = {}           # layer index -> async handles / CUDA event
        self.num_prefetches = 0
        self.num_forwards = 0
        self.bytes_moved = 0

    # ---- decision ---------------------------------------------------------------------------------------------------------------
    def use(self, ctx: Any, x: torch.Tensor) -> bool:
        return sum(ctx.counts("dwdp", x.shape[0], x.device)) >= self.min_tokens

    def reset(self) -> None:
        """New forward: weights may have been hot-swapped, so nothing gathered earlier is trusted."""
        self._drain()

    def _drain(self) -> None:
        for i in list(self.inflight):
            self._wait(i)
        self.inflight.clear()

    # ---- transport --------------------------------------------------------------------------------------------------------------
    def _post(self, i: int) -> None:
        m, (b13, b2) = self.layers[i], self.bufs[i % 2]
        self.num_prefetches += 1
        self.bytes_moved += (b13.numel() - m.w13.numel()) * b13.element_size() + (b2.numel() - m.w2.numel()) * b2.element_size()
        if self.side is not None:
            self.side.wait_stream(torch.cuda.current_stream())          # the buffer's previous reader (layer i-2) is ordered before us
            with torch.cuda.stream(self.side):
                dist.all_gather_into_tensor(b13, m.w13.data, group=self.group.device_group)
                dist.all_gather_into_tensor(b2, m.w2.data, group=self.group.device_group)
                ev = torch.cuda.Event()
                ev.record(self.side)
            self.inflight[i] = ev
        else:
            self.inflight[i] = [dist.all_gather_into_tensor(b13, m.w13.data.contiguous(), group=self.group.device_group, async_op=True),
                                dist.all_gather_into_tensor(b2, m.w2.data.contiguous(), group=self.group.device_group, async_op=True)]

    def _wait(self, i: int) -> None:
        h = self.inflight.get(i)
        if h is None or h is True:
            return
        if self.side is not None:
            torch.cuda.current_stream().wait_event(h)
        else:
            for w in h:
                w.wait()
        self.inflight[i] = True                                           # gathered and still resident in its buffer

    def weights(self, i: int) -> Tuple[torch.Tensor, torch.Tensor]:
        if i not in self.inflight:
            self.inflight.pop(i - 2, None)                                # same buffer
            self._post(i)
        self._wait(i)
        if i + 1 < len(self.layers) and (i + 1) not in self.inflight:
            self.inflight.pop(i - 1, None)                                # layer i-1 is done with the other buffer
            self._post(i + 1)
        return self.bufs[i % 2]

    def forward(self, layer: Any, x: torch.Tensor, w: torch.Tensor, ids: torch.Tensor) -> torch.Tensor:
        from ..ops import moe_ops

        self.num_forwards += 1
        w13, w2 = self.weights(self.index[id(layer)])
        if x.shape[0] == 0:
            return x
        return moe_ops.fused_moe(x, w13, w2, w, ids, 0)


_MGR: Optional[DwdpManager] = None


def get_dwdp_manager() -> Optional[DwdpManager]:
    return _MGR


def set_dwdp_manager(m: Optional[DwdpManager]) -> None:
    global _MGR
    _MGR = m


def attach(runner: Any, min_tokens: int) -> Optional[DwdpManager]:
    """Build the manager for a loaded model (whole experts per rank: ep_size == tp_size, no TP inside an expert)."""
    layers = [m for m in runner.model.modules() if m.__class__.__name__ == "FusedMoE"]
    if not layers or any(m.moe_tp != 1 or m.ep_size != m.tp.world_size for m in layers):
        raise ValueError("--enable-dwdp needs an MoE model with ep_size == tp_size")
    mgr = DwdpManager(layers, layers[0].tp, min_tokens)
    set_dwdp_manager(mgr)
    runner.dwdp = mgr
    return mgr
