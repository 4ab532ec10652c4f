"""ForwardBatch: the device-side view of one scheduler step.

Parity: reference ``python/sglang/srt/model_executor/forward_batch_info.py`` (``ForwardMode``, ``init_new:705``,
``compute_position:1771``) and SURVEY Appendix A.5.

B200-first: ``ForwardBatch.build`` packs every per-step integer (input ids, positions, KV slot ids, request
slots, sequence lengths, query offsets, page-table deltas, last-token indices) into ONE pinned int32 buffer and
issues ONE host->device copy; views into that device buffer are handed to the kernels.  Sampling parameters ride
in a second (float) region of the same buffer.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, List, Optional

import numpy as np
import torch

from ..managers.schedule_batch import ForwardMode, ScheduleBatch


@dataclass
class SamplingTensors:
    temperature: torch.Tensor   # f32 [B]
    top_k: torch.Tensor         # i32 [B]
    top_p: torch.Tensor         # f32 [B]
    min_p: torch.Tensor         # f32 [B]
    all_greedy: bool = True
    need_filter: bool = False
    seeds: Optional[torch.Tensor] = None
    has_penalty: bool = False


@dataclass
class ForwardBatch:
    mode: ForwardMode
    batch_size: int
    num_tokens: int
    input_ids: Optional[torch.Tensor]     # i32 [T] (extend) / None (decode gathers from next_token_buf)
    positions: torch.Tensor               # i32 [T]
    out_cache_loc: torch.Tensor           # i32 [T]
    req_slots: torch.Tensor               # i32 [B]
    seq_lens: torch.Tensor                # i32 [B]
    cu_q: Optional[torch.Tensor]          # i32 [B+1] (extend)
    q_lens: Optional[torch.Tensor]        # i32 [B]
    last_idx: Optional[torch.Tensor]      # i32 [Bs] rows of the hidden states that feed the LM head
    sample_slots: Optional[torch.Tensor]  # i32 [Bs] request slots of the sampled rows
    max_q_len: int = 1
    max_kv_len: int = 0
    page_table: Optional[torch.Tensor] = None   # i32 [max_reqs+1, max_pages] (device, persistent)
    page_size: int = 1
    kv_pool: Any = None
    num_splits: int = 1
    attn_workspace: Any = None
    sampling: Optional[SamplingTensors] = None
    next_token_buf: Optional[torch.Tensor] = None
    mm_embeds: Any = None            # (row indices, embeddings) to splice into the input embeddings
    lora_ids: Optional[List[Optional[str]]] = None
    return_logprob: bool = False
    return_hidden: bool = False
    spec_info: Any = None
    custom_mask: Optional[torch.Tensor] = None
    custom_mask_sliding: Optional[torch.Tensor] = None      # the sliding-window layers' variant (Gemma-3 bidirectional image blocks)
    tree_anc: Optional[torch.Tensor] = None      # i32 [T] ancestor bits per query node (speculative tree verify on the tcgen05 kernel)
    h2d_bytes: int = 0
    # host mirrors (no sync needed)
    seq_lens_cpu: Optional[np.ndarray] = None
    q_lens_cpu: Optional[np.ndarray] = None
    pp_hidden: Any = None
    capture_aux: Any = None          # EAGLE-3: set of (global) layer indices whose input hidden state is captured ...
    aux_hidden: Any = None           # ... into this [T, n_aux * H] tensor by the model
    rope_override: Any = None        # (cos_sin rows [T, rot], row index [T]): per-token rotary table (M-RoPE)
    mrope_positions: Any = None      # i64 [3, T]
    cp: Any = None                   # layers/cp.py CpMeta inside a context-parallel prefill
    mamba_snap: Any = None           # {request index: (tokens into this chunk, snapshot row)}: recurrent-state snapshots to take in this forward
    out_cache_loc_swa: Optional[torch.Tensor] = None   # i32 [T] slots in the sliding-window pool (hybrid SWA models, mem_cache/swa.py)
    swa_pt_offset: int = 0           # the sliding page table is columns [offset, 2 * offset) of ``page_table``
    dcp: Any = None                  # layers/dcp.py DcpMeta when decode context parallelism is on


class BatchPacker:
    """Reusable pinned staging buffer + packing logic."""

    def __init__(self, device: torch.device, capacity_ints: int = 1 << 20):
        self.device = device
        self.cap = capacity_ints
        self.pinned = None
        self.is_cuda = device.type == "cuda"
        if self.is_cuda:
            self.pinned = torch.empty(capacity_ints, dtype=torch.int32).pin_memory()
            self._ring = [torch.empty(capacity_ints, dtype=torch.int32).pin_memory() for _ in range(3)]
            self._ring_i = 0

    def _stage(self, arr: np.ndarray) -> torch.Tensor:
        n = arr.size
        if not self.is_cuda:
            return torch.from_numpy(arr.copy())
        if n > self.cap:
            self.cap = int(n * 1.5)
            self._ring = [torch.empty(self.cap, dtype=torch.int32).pin_memory() for _ in range(3)]
        buf = self._ring[self._ring_i]
        self._ring_i = (self._ring_i + 1) % len(self._ring)
        buf[:n].numpy()[:] = arr
        return buf[:n].to(self.device, non_blocking=True)

    def build(self, sb: ScheduleBatch, page_table: torch.Tensor, page_size: int, kv_pool: Any, next_token_buf: torch.Tensor,
              apply_delta) -> ForwardBatch:
        B = len(sb.reqs)
        decode = sb.forward_mode.is_decode()
        T = B if decode else int(sb.q_lens.sum())
        parts: List[np.ndarray] = []
        offs = {}

        def add(name: str, a: np.ndarray) -> None:
            offs[name] = (sum(p.size for p in parts), a.size)
            parts.append(np.ascontiguousarray(a, dtype=np.int32).reshape(-1))

        add("positions", sb.positions)
        add("out_cache_loc", sb.out_cache_loc)
        if sb.out_cache_loc_swa is not None:
            add("out_cache_loc_swa", sb.out_cache_loc_swa)
        add("req_slots", sb.req_slots)
        add("seq_lens", sb.seq_lens)
        if not decode:
            add("input_ids", sb.input_ids)
            cu = np.zeros(B + 1, dtype=np.int32)
            np.cumsum(sb.q_lens, out=cu[1:])
            add("cu_q", cu)
            add("q_lens", sb.q_lens)
            samp = np.flatnonzero(sb.sample_mask)
            add("last_idx", (cu[1:][samp] - 1).astype(np.int32))
            add("sample_slots", sb.req_slots[samp])
        else:
            samp = np.arange(B)
        nd = 0 if sb.pt_rows is None else sb.pt_rows.size
        if nd:
            add("pt_rows", sb.pt_rows)
            add("pt_cols", sb.pt_cols)
            add("pt_vals", sb.pt_vals)
        # sampling parameters (float32 bit patterns ride in the same int32 buffer)
        reqs = [sb.reqs[i] for i in samp.tolist()]
        Bs = len(reqs)
        sp = [r.sampling_params for r in reqs]
        temp = np.asarray([p.temperature for p in sp], dtype=np.float32)
        topp = np.asarray([p.top_p for p in sp], dtype=np.float32)
        minp = np.asarray([p.min_p for p in sp], dtype=np.float32)
        topk = np.asarray([min(p.top_k, 1 << 30) for p in sp], dtype=np.int32)
        add("temperature", temp.view(np.int32))
        add("top_p", topp.view(np.int32))
        add("min_p", minp.view(np.int32))
        add("top_k", topk)
        packed = np.concatenate(parts) if parts else np.empty(0, dtype=np.int32)
        dev = self._stage(packed)

        def v(name: str) -> Optional[torch.Tensor]:
            if name not in offs:
                return None
            o, n = offs[name]
            return dev[o:o + n]

        if nd:
            apply_delta(page_table, v("pt_rows"), v("pt_cols"), v("pt_vals"))
        seeds = None
        if any(p.sampling_seed is not None for p in sp):
            # (seed, the request's own output position): scheduling-independent streams; -1 = unseeded row (global stream).  The position
            # is exact because the scheduler drains the pipeline before a batch that needs host-side sampling state (needs_host_state)
            seeds = torch.tensor([[r.sampling_params.sampling_seed, len(r.output_ids)] if r.sampling_params.sampling_seed is not None
                                  else [0, -1] for r in reqs], dtype=torch.int64, device=self.device)
        sampling = SamplingTensors(
            temperature=v("temperature").view(torch.float32), top_k=v("top_k"), top_p=v("top_p").view(torch.float32),
            min_p=v("min_p").view(torch.float32), all_greedy=bool((topk == 1).all()) if Bs else True,
            need_filter=bool(((topk != 1) & ((topk < (1 << 30)) | (topp < 1.0) | (minp > 0.0))).any()) if Bs else False,
            seeds=seeds,
            has_penalty=any(p.frequency_penalty != 0 or p.presence_penalty != 0 or p.repetition_penalty != 1.0 or p.logit_bias for p in sp))
        fb = ForwardBatch(
            mode=sb.forward_mode, batch_size=B, num_tokens=T, input_ids=v("input_ids"), positions=v("positions"),
            out_cache_loc=v("out_cache_loc"), req_slots=v("req_slots"), seq_lens=v("seq_lens"), cu_q=v("cu_q"),
            q_lens=v("q_lens"), last_idx=v("last_idx"), sample_slots=v("sample_slots") if not decode else v("req_slots"),
            max_q_len=1 if decode else int(sb.q_lens.max()), max_kv_len=int(sb.seq_lens.max()) if B else 0,
            page_table=page_table, page_size=page_size, kv_pool=kv_pool, sampling=sampling, next_token_buf=next_token_buf,
            return_logprob=sb.return_logprob, h2d_bytes=int(packed.size * 4), seq_lens_cpu=sb.seq_lens,
            q_lens_cpu=sb.q_lens, spec_info=sb.spec_info, out_cache_loc_swa=v("out_cache_loc_swa"), mamba_snap=getattr(sb, "mamba_snap", None))
        return fb
