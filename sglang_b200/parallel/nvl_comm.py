"""NvlComm: symmetric-memory arena over NVLink 5 / NVSwitch + the collectives built on it.

Stands in for the reference's custom all-reduce v1/v2 (``distributed/device_communicators/custom_all_reduce_v2.py:98``),
torch-symm-mem AR (``torch_symm_mem.py:34-185``) and FlashInfer AR+RMSNorm fusion (``layers/flashinfer_comm_fusion.py``).
Plumbing (allocation, handle exchange, peer + multicast mapping) is delegated to ``torch.distributed._symmetric_memory``
(CUDA VMM + fabric/posix-fd handles underneath); every kernel that touches the arena is ours
(``csrc/comm/nvl_allreduce.cu``, ``csrc/comm/gemm_allreduce.cu``).

Arena layout per rank (identical on all ranks):  slot[parity] = [ input region | result region ], parity = epoch & 1.
Hazard argument for double buffering: call k writes slot k%2 *before* its flag round; a slow peer may still be reading
slot (k-1)%2 from call k-1, never slot k%2, because having passed the flag round of call k-1 implies it finished all
reads of call k-2.
"""
from __future__ import annotations

import logging
from typing import Any, Optional

import torch

from ..environ import envs

logger = logging.getLogger(__name__)

CH_ONESHOT, CH_TWOSHOT, CH_FUSED_NORM, CH_GEMM, CH_RS_AG = 0, 1, 2, 3, 4
NUM_CHANNELS = 8
MAX_RANKS = 8          # must match NVL_MAX_RANKS in csrc/comm/nvl_allreduce.cu
FLAG_OFFSET = 1024


class NvlComm:
    def __init__(self, group: Any, buf: torch.Tensor, hdl: Any, slot_bytes: int):
        self.group = group
        self.rank, self.world = group.rank, group.world_size
        self.buf, self.hdl = buf, hdl
        self.slot_bytes = slot_bytes
        dev = buf.device
        self.buf_ptrs = torch.tensor([int(p) for p in hdl.buffer_ptrs], dtype=torch.int64)
        # our flag words live FLAG_OFFSET bytes into each rank's signal pad (the head of the pad is left to torch's own ops)
        self.flag_ptrs = torch.tensor([int(p) + FLAG_OFFSET for p in hdl.signal_pad_ptrs], dtype=torch.int64)
        self.mc_ptr = int(getattr(hdl, "multicast_ptr", 0) or 0)
        # per-channel epoch counters + completion counters (device memory; graph-replay safe)
        self.epochs = torch.zeros(NUM_CHANNELS, dtype=torch.int32, device=dev)
        self.done = torch.zeros(NUM_CHANNELS, 2, dtype=torch.int32, device=dev)
        self.host_epoch = [0] * NUM_CHANNELS   # host mirror used only to pick the slot parity
        self.graph_epoch = [0] * NUM_CHANNELS
        self.capturing = False
        # one-shot -> two-shot crossover: 512 KB on 8 GPUs (reference k3_ar_fusion.py:36-38); with two ranks a reduce-scatter +
        # all-gather moves exactly the same bytes as the one-shot exchange, so two-shot never wins (measured: profiles/mgpu_tp2_r1.md)
        self.max_oneshot_bytes = (1 << 40) if self.world == 2 else 512 * 1024
        self.calls = 0
        self.use_nvls = bool(self.mc_ptr) and envs.SGLB_NVLS.get()      # NVSwitch in-network reduction when a multicast alias exists
        # large-T row-parallel outputs: reduce-scatter + norm + all-gather with a ROW-SHARDED residual stream (csrc/comm/nvl_allreduce.cu
        # nvl_rs_add_rmsnorm_ag_kernel).  The model runner switches it on per forward (``rs_active``) for model families whose every residual
        # update goes through a fused consume_norm; with two ranks the one-shot exchange moves the same bytes, so it stays off there.
        _rs = envs.SGLB_RS_AG_MIN_BYTES.get()
        # validated at 4 ranks only: the 8-rank run of tools/mgpu_test.py died in this kernel (launch failure, run N) -> refused above 4 ranks
        self.rs_min_bytes = _rs if (2 < self.world <= 4 and _rs > 0) else (1 << 60)
        self.rs_active = False
        _lib = envs.SGLB_TP_LIBRARY_AR_MIN_BYTES.get()
        self.library_ar_min_bytes = _lib if (self.world > 2 and _lib > 0) else (1 << 60)
        self.rs_calls = 0

    @classmethod
    def create(cls, group: Any) -> "NvlComm":
        import torch.distributed._symmetric_memory as symm_mem

        total = envs.SGLB_NVL_MAX_BYTES.get()
        slot = total // 4          # {eager, graph} x {even, odd}
        dev = group.device
        gname = group.device_group.group_name
        try:   # older torch needs an explicit opt-in per group; newer versions made it a no-op / deprecated
            symm_mem.enable_symm_mem_for_group(gname)
        except Exception:
            pass
        buf = symm_mem.empty(total, dtype=torch.uint8, device=dev)
        hdl = symm_mem.rendezvous(buf, gname)
        buf.zero_()
        try:   # zero our flag words in the local signal pad
            pad = hdl.get_signal_pad(hdl.rank)
            pad.view(torch.uint8)[FLAG_OFFSET:FLAG_OFFSET + NUM_CHANNELS * MAX_RANKS * 4].zero_()
        except Exception:
            pass
        torch.cuda.synchronize(dev)
        group.barrier()
        c = cls(group, buf, hdl, slot)
        logger.info("NVL arena: %d MiB/rank, world=%d, multicast=%s", total >> 20, c.world, bool(c.mc_ptr))
        return c

    # ------------------------------------------------------------------ helpers
    def _slot(self, channel: int) -> int:
        """byte offset of the slot used by the NEXT call on this channel, and advance the host mirror.
        The arena has two regions x two parities: eager calls alternate inside region 0; calls recorded into a CUDA
        graph alternate inside region 1 starting from parity 0 at every capture (a forward pass stages an even number of
        times per channel, so back-to-back replays of any graphs keep alternating, and an eager step followed by a graph
        replay — or vice versa — never reuses the slot of the immediately preceding collective)."""
        ctr = self.graph_epoch if self.capturing else self.host_epoch
        ctr[channel] += 1
        region = 1 if self.capturing else 0
        # the reduce-scatter channel owns the upper half of every slot: its slots never alias the one-shot channels' slots
        sub = self.slot_bytes // 2 if channel == CH_RS_AG else 0
        return region * 2 * self.slot_bytes + (ctr[channel] & 1) * self.slot_bytes + sub

    def begin_capture(self) -> None:
        self.capturing = True
        self.graph_epoch = [0] * NUM_CHANNELS

    def end_capture(self) -> None:
        self.capturing = False

    def slot_view(self, offset: int, shape, dtype=torch.bfloat16) -> torch.Tensor:
        n = 1
        for s in shape:
            n *= s
        nbytes = n * torch.empty((), dtype=dtype).element_size()
        return self.buf[offset:offset + nbytes].view(dtype).view(*shape)

    def can_all_reduce(self, x: torch.Tensor) -> bool:
        nbytes = x.numel() * x.element_size()
        return (x.is_cuda and x.dtype == torch.bfloat16 and x.is_contiguous() and x.numel() % 8 == 0
                and 4 * nbytes <= self.slot_bytes)

    # ------------------------------------------------------------------ collectives
    def all_reduce(self, x: torch.Tensor, algo: Optional[int] = None) -> torch.Tensor:
        nbytes = x.numel() * x.element_size()
        if algo is None:
            # algo 0/1: peer pull one-/two-shot; 2/3: NVLS (switch-side reduction: multimem.ld_reduce / multimem.st)
            small = nbytes <= self.max_oneshot_bytes
            algo = (2 if small else 3) if self.use_nvls else (0 if small else 1)
        ch = CH_ONESHOT if algo in (0, 2, 4) else CH_TWOSHOT
        off = self._slot(ch)
        self.slot_view(off, x.shape).copy_(x)      # stage into the arena (the fused GEMM epilogue writes here directly)
        out = torch.empty_like(x)
        torch.ops.sglb.nvl_allreduce(self.buf_ptrs, self.flag_ptrs, self.mc_ptr, off, self.rank, self.world, out, self.epochs[ch:ch + 1],
                                     self.done[ch], ch, algo)
        self.calls += 1
        return out

    def stage(self, channel: int, shape, dtype=torch.bfloat16):
        """reserve this call's arena slot; returns (offset, view).  The producer (GEMM) writes its partial output there."""
        off = self._slot(channel)
        return off, self.slot_view(off, shape, dtype)

    def allreduce_from_slot(self, off: int, out: torch.Tensor, channel: int = CH_FUSED_NORM) -> None:
        """one-shot all-reduce of partials that a producer already wrote into slot ``off`` (no staging copy)"""
        torch.ops.sglb.nvl_allreduce(self.buf_ptrs, self.flag_ptrs, self.mc_ptr, off, self.rank, self.world, out,
                                     self.epochs[channel:channel + 1], self.done[channel], channel, 2 if self.use_nvls else 0)
        self.calls += 1

    def fused_allreduce_add_rmsnorm(self, off: int, residual: Optional[torch.Tensor], weight: torch.Tensor, out: torch.Tensor,
                                    eps: float, gemma: bool = False) -> None:
        """consumer side of a row-parallel GEMM whose partial output already sits in arena slot ``off``"""
        T, H = out.shape
        ch = CH_FUSED_NORM
        torch.ops.sglb.nvl_ar_add_rmsnorm(self.buf_ptrs, self.flag_ptrs, self.mc_ptr if self.use_nvls else 0, off, self.rank, self.world, residual, weight, out,
                                          float(eps), gemma, self.epochs[ch:ch + 1], self.done[ch], ch)
        self.calls += 1

    def wants_rs_bytes(self, nbytes: int) -> bool:
        return nbytes >= self.rs_min_bytes and 8 * nbytes <= self.slot_bytes

    def wants_rs(self, nbytes: int) -> bool:
        """row-parallel producers ask this when they stage their partial output: large enough and allowed for this forward?"""
        return self.rs_active and nbytes >= self.rs_min_bytes and 8 * nbytes <= self.slot_bytes

    def rs_add_rmsnorm_ag(self, off: int, T: int, H: int, residual: Optional[torch.Tensor], weight: torch.Tensor, eps: float,
                          gemma: bool = False) -> torch.Tensor:
        """consumer of partials staged with ``stage(CH_RS_AG, ...)``: returns the normalised [T, H] activations as a VIEW of this rank's arena
        (result region of the same slot); ``residual`` is updated on this rank's rows (t % world == rank) only"""
        nbytes = T * H * 2
        out_off = (nbytes + 255) // 256 * 256
        ch = CH_RS_AG
        torch.ops.sglb.nvl_rs_add_rmsnorm_ag(self.buf_ptrs, self.flag_ptrs, self.mc_ptr if self.use_nvls else 0, off, out_off, self.rank, self.world,
                                             residual, weight, T, H, float(eps), gemma, self.epochs[ch:ch + 1], self.done[ch], ch)
        self.calls += 1
        self.rs_calls += 1
        return self.slot_view(off + out_off, (T, H))

    # ------------------------------------------------------------------ fused GEMM + all-reduce (push)
    def gemm_push(self, x: torch.Tensor, w: torch.Tensor, bn: int = 0) -> int:
        """x[T,K] @ w[N,K]^T computed by the tcgen05 GEMM whose epilogue stores every tile into slot [rank] of EVERY peer's
        arena and finally publishes this call's epoch to all peers.  Returns the slot offset for ``push_reduce_add_rmsnorm``."""
        T, N = x.shape[0], w.shape[0]
        assert self.world * T * N * 2 <= self.slot_bytes, "NVL arena slot too small for this GEMM output"
        ch = CH_GEMM
        off = self._slot(ch)
        slots = self.buf_ptrs + off
        flags = self.flag_ptrs + ch * MAX_RANKS * 4
        torch.ops.sglb.gemm_push(x, w, slots, flags, self.rank, self.world, self.epochs[ch:ch + 1], self.done[ch, 1:], bn)
        self.calls += 1
        return off

    def moe_push(self, x: torch.Tensor, w13: torch.Tensor, w2: torch.Tensor, topk_w: torch.Tensor, topk_ids: torch.Tensor,
                 expert_offset: int) -> int:
        """expert MLP block (align -> gather -> grouped GEMMs -> combine) whose combine kernel stores this rank's partial
        [T,H] straight into slot [rank] of every peer's arena: the expert-parallel reduction costs no kernel of its own."""
        T, H = x.shape
        assert self.world * T * H * 2 <= self.slot_bytes, "NVL arena slot too small for this MoE output"
        ch = CH_GEMM
        off = self._slot(ch)
        torch.ops.sglb.moe_grouped_mlp_push(x, w13, w2, topk_w, topk_ids, expert_offset, self.buf_ptrs + off,
                                            self.flag_ptrs + ch * MAX_RANKS * 4, self.rank, self.world, self.epochs[ch:ch + 1],
                                            self.done[ch, 1:])
        self.calls += 1
        return off

    def gemm_push_skinny(self, x: torch.Tensor, w: torch.Tensor, max_splits: int = 4):
        """decode-shape variant of ``gemm_push`` (csrc/kernels/gemm_skinny.cu): W streams as the UMMA M operand, K is split so
        ~one CTA per SM is busy, and every (rank, split) partial lands as slab [rank*splits + split] in every peer's arena.
        Returns (slot offset, splits) for ``push_reduce_add_rmsnorm``."""
        T, K = x.shape
        N = w.shape[0]
        splits = int(torch.ops.sglb.gemm_skinny_splits(T, N, K, max_splits))
        while splits > 1 and self.world * splits * T * N * 2 > self.slot_bytes:
            splits = int(torch.ops.sglb.gemm_skinny_norm_splits(K, splits - 1))
        assert self.world * splits * T * N * 2 <= self.slot_bytes, "NVL arena slot too small for this GEMM output"
        ch = CH_GEMM
        off = self._slot(ch)
        torch.ops.sglb.gemm_skinny_push(x, w, splits, self.buf_ptrs + off, self.flag_ptrs + ch * MAX_RANKS * 4, self.rank, self.world,
                                        self.epochs[ch:ch + 1], self.done[ch, 1:])
        self.calls += 1
        return off, splits

    def push_reduce_add_rmsnorm(self, off: int, T: int, H: int, residual: Optional[torch.Tensor], weight: torch.Tensor, out: torch.Tensor,
                                eps: float, gemma: bool = False, splits: int = 1) -> None:
        ch = CH_GEMM
        nslabs = self.world * splits
        slot = self.slot_view(off, (nslabs, T, H))
        torch.ops.sglb.nvl_local_reduce_add_rmsnorm(slot, int(self.flag_ptrs[self.rank]) + ch * MAX_RANKS * 4, residual, weight, out,
                                                    self.world, nslabs, float(eps), gemma, self.epochs[ch:ch + 1], self.done[ch])
        self.calls += 1
