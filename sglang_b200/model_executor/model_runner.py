"""ModelRunner: owns the model, the KV pool, device-side request state and the CUDA-graph decode path.

Parity: reference ``python/sglang/srt/model_executor/model_runner.py:285`` (``initialize:642``, ``load_model:1072``,
pool sizing ``pool_configurator.py``, ``forward:1520``, ``sample:1792``), ``managers/tp_worker.py:299``
(``forward_batch_generation:574``), ``runner/decode_cuda_graph_runner.py:200`` (capture per batch-size bucket,
``load_batch:1240``, ``execute:1386``) and ``managers/overlap_utils.py:246`` (FutureMap).

B200-first notes
* one H2D copy per step (``BatchPacker``), page table kept on device and patched by deltas;
* sampled tokens are scattered into ``next_token_buf[req_slot]`` by the sampling kernel and gathered by the next
  decode's embedding kernel: the "future map" is a device array, the host only reads results one step late;
* decode runs as ONE ``cudaGraphLaunch`` per step (embedding -> layers -> LM head -> sampling);
* everything is enqueued on a private stream; results are copied to pinned memory and fenced with an event.
"""
from __future__ import annotations

import json
import logging
import time
from typing import Any, Dict, List, Optional

import numpy as np
import torch

from .. import ops
from ..configs.model_config import ModelConfig, load_model_config
from ..environ import envs
from ..managers import io_struct as ios
from ..managers.schedule_batch import ForwardMode, ScheduleBatch
from ..mem_cache.memory_pool import KVPool
from ..model_loader.loader import load_model
from ..parallel.parallel_state import get_pp_group, get_state, get_tp_group
from ..server_args import ServerArgs
from .forward_batch import BatchPacker, ForwardBatch, SamplingTensors

logger = logging.getLogger(__name__)


def _llama_forward():
    from ..models.llama import LlamaForCausalLM

    return LlamaForCausalLM.forward


_LLAMA_FORWARD = _llama_forward()
_DT = {"bfloat16": torch.bfloat16, "float16": torch.float16, "float32": torch.float32}


class _null_ctx:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


class BatchResult:
    """Asynchronous result of one step; ``wait()`` blocks until the device->host copy has landed."""

    wait_s = 0.0      # class-wide: seconds the host spent blocked in event.synchronize()

    def __init__(self) -> None:
        self.event: Optional[torch.cuda.Event] = None
        self._tok_host: Optional[torch.Tensor] = None
        self._lp_host: Optional[torch.Tensor] = None
        self.next_token_ids: Optional[np.ndarray] = None   # aligned with batch.reqs (non-sampled rows = -1)
        self.next_token_logprobs: Optional[np.ndarray] = None
        self.top_logprobs: Any = None
        self.input_logprobs: Any = None
        self.embeddings: Any = None
        self.hidden_states: Any = None
        self.sample_rows: Optional[np.ndarray] = None
        self.batch_size = 0
        self.d2h_bytes = 0
        self.accept_lens = None
        self.accepted_tokens = None

    def wait(self) -> None:
        if self.next_token_ids is not None or self._tok_host is None:
            return
        if self.event is not None:
            t0 = time.perf_counter()
            self.event.synchronize()
            BatchResult.wait_s += time.perf_counter() - t0      # host blocked on the device (accounting for bench.py's host/device split)
        toks = self._tok_host.numpy()
        out = np.full(self.batch_size, -1, dtype=np.int64)
        out[self.sample_rows] = toks[: self.sample_rows.size]
        self.next_token_ids = out
        if self._lp_host is not None:
            lp = np.zeros(self.batch_size, dtype=np.float32)
            lp[self.sample_rows] = self._lp_host.numpy()[: self.sample_rows.size]
            self.next_token_logprobs = lp


class ModelRunner:
    def __init__(self, server_args: ServerArgs, model_config: Optional[ModelConfig] = None):
        self.args = server_args
        self.cfg = model_config or load_model_config(server_args.model_path, json.loads(server_args.json_model_override_args or "{}"),
                                                     server_args.context_length)
        st = get_state()
        self.tp, self.pp = st.tp, st.pp
        self.is_cuda = server_args.device == "cuda"
        self.device = torch.device("cuda", torch.cuda.current_device()) if self.is_cuda else torch.device("cpu")
        self.dtype = _DT[self.cfg.dtype] if self.is_cuda or self.cfg.dtype != "bfloat16" else torch.float32
        if not self.is_cuda and self.cfg.dtype == "bfloat16":
            self.cfg.dtype = "float32"      # CPU tier runs fp32
            self.dtype = torch.float32
        if self.is_cuda:
            ops.load_native()
            if server_args.gemm_backend in ("tc", "lib"):
                envs.SGLB_GEMM_IMPL.set(server_args.gemm_backend)
            if server_args.attention_backend in ("tc", "simt"):
                envs.SGLB_ATTN_IMPL.set(server_args.attention_backend)
        self.page_size = server_args.page_size
        from ..parallel.parallel_state import get_dcp_group

        self.dcp = get_dcp_group()
        from ..parallel.parallel_state import get_cp_group, get_state as _gs

        self.cp = get_cp_group()
        self.cp_rank = _gs().attn_dp_rank if self.cp is not None else 0
        self.cp_min_tokens = max(2 * (self.cp.world_size if self.cp is not None else 1), envs.SGLB_CP_MIN_TOKENS.get())
        self.cp_forwards = 0
        assert self.cp is None or self.dcp is None, "prefill CP and decode CP are separate modes"
        self.local_page_size = self.page_size // (self.dcp.world_size if self.dcp is not None else 1)
        torch.manual_seed(server_args.random_seed)
        t0 = time.perf_counter()
        if self.cfg.is_moe:
            from ..models import mixtral

            mixtral.set_ep_size(server_args.ep_size)
        if server_args.language_only and self.cfg.vision:
            self.cfg.vision = {**self.cfg.vision, "skip_tower": True}
        remote = None
        if server_args.load_format == "remote_instance":
            remote = {"seed_url": f"http://{server_args.remote_instance_weight_loader_seed_instance_ip}:"
                                  f"{server_args.remote_instance_weight_loader_seed_instance_service_port}",
                      "port": server_args.remote_instance_weight_loader_send_weights_group_port, "address": server_args.host
                      if server_args.host not in ("0.0.0.0", "") else "127.0.0.1"}
        self.model = load_model(self.cfg, server_args.model_path, server_args.load_format, self.device, server_args.quantization,
                                server_args.random_seed, remote)
        self.load_time = time.perf_counter() - t0
        self.weight_bytes = sum(p.numel() * p.element_size() for p in self.model.parameters())
        self.stream = torch.cuda.Stream(self.device) if self.is_cuda else None
        self.num_sms = torch.cuda.get_device_properties(self.device).multi_processor_count if self.is_cuda else 1
        self.kv_manager = None
        self._init_memory()
        self.packer = BatchPacker(self.device)
        # pinned staging rings: the host may run up to two steps ahead of the device (overlap scheduling), so a
        # staging buffer is only reused after >= 4 later steps
        self._stage_ring: List[torch.Tensor] = []
        self._res_ring: List[Any] = []
        self._ring_i = 0
        if self.is_cuda:
            n_stage = max(8 * max(self.max_running, 512), 3 * 4096 + 10 * 128 + 8) + 2 + 3 * 8192
            self._stage_ring = [torch.zeros(n_stage, dtype=torch.int32).pin_memory() for _ in range(4)]
            self._res_ring = [(torch.zeros(self.max_running + 8, dtype=torch.int32).pin_memory(),
                               torch.zeros(self.max_running + 8, dtype=torch.float32).pin_memory()) for _ in range(8)]
        self.graphs: Dict[int, Any] = {}
        self.graphs_greedy: Dict[int, Any] = {}      # TP > 1: all-greedy batches skip the [B, V] logits all-gather (vocab-parallel argmax)
        self.prefill_graphs: Dict[int, Any] = {}     # token-bucketed prefill graphs (small extends are launch-bound: ~10 kernels per layer)
        self.prefill_graph_tokens: List[int] = []
        self.prefill_graph_hits = 0
        self.graph_bs: List[int] = []
        self.graph_launches = 0       # kernels replayed through CUDA graphs (for gpu_launches accounting)
        self.step_counter = 0
        self.forward_ct = 0
        self.h2d_bytes = 0
        self.d2h_bytes = 0
        self.lora_manager = None
        self.hidden_tap = None          # callable(ScheduleBatch, hidden[T,H]) installed by feature-based draft workers (EAGLE)
        self.last_verify_hidden = None
        self.aux_layer_ids = None       # EAGLE-3: capture the hidden states entering these layers instead of the final one
        self.profiler = None
        self.expert_recorder = None
        self.dwdp = None
        if getattr(server_args, "enable_dwdp", False):
            if not server_args.enable_dp_attention:
                raise ValueError("--enable-dwdp rides on DP-attention (tokens stay on their rank)")
            from ..layers import dwdp as _dwdp

            _dwdp.attach(self, server_args.dwdp_min_tokens)
        if (self.is_cuda and not server_args.disable_cuda_graph and not envs.SGLB_DISABLE_CUDA_GRAPH.get() and self.pp.world_size == 1
                and getattr(self.model, "supports_cuda_graph", True) and not server_args.enable_dp_attention and self.dcp is None):
            with torch.cuda.stream(self.stream):
                self._capture_graphs()
            self.stream.synchronize()

    # ------------------------------------------------------------------ memory
    def _init_memory(self) -> None:
        cfg, a = self.cfg, self.args
        n_layers = len(self.model.model.layers)
        a0 = self.model.model.layers[0].self_attn.attn if n_layers else None
        kv_heads = a0.num_kv_heads if a0 is not None else cfg.num_key_value_heads
        kv_dim = getattr(a0, "kv_head_dim", None) or cfg.head_dim      # MLA caches one 576-wide latent row per token
        kv_dtype = self.dtype
        if a.kv_cache_dtype in ("fp8_e4m3", "fp8"):
            kv_dtype = torch.float8_e4m3fn
        if self.dcp is not None:
            assert self.page_size % self.dcp.world_size == 0, "DCP: --page-size must be a multiple of --dcp-size"
            q0 = self.model.model.layers[0].self_attn.qkv_proj if n_layers and hasattr(self.model.model.layers[0].self_attn, "qkv_proj") else None
            assert q0 is not None and getattr(q0, "kv_replicas", 1) % self.dcp.world_size == 0, \
                "DCP: dcp_size must divide the KV-head replication factor tp_size / num_kv_heads (nothing to de-duplicate otherwise)"
        # hybrid sliding-window models (--swa-full-tokens-ratio > 0): the sliding layers move to a second, smaller pool (mem_cache/swa.py)
        sliding = [i for i, l in enumerate(self.model.model.layers) if getattr(getattr(getattr(l, "self_attn", None), "attn", None), "window", -1) > 0] \
            if n_layers else []
        self.swa = None
        ratio = float(getattr(a, "swa_full_tokens_ratio", 0.0) or 0.0)
        if ratio > 0 and sliding and len(sliding) < n_layers:
            if self.dcp is not None or self.cp is not None or self.pp.world_size > 1 or a.enable_hierarchical_cache or a.speculative_algorithm:
                raise ValueError("--swa-full-tokens-ratio: not combined with DCP / CP / PP / HiCache / speculative decoding yet")
            self.swa = {"layers": sliding, "ratio": min(1.0, ratio), "window": max(self.model.model.layers[i].self_attn.attn.window for i in sliding)}
        eff_layers = n_layers if self.swa is None else (n_layers - len(sliding)) + self.swa["ratio"] * len(sliding)
        latent_v = int(getattr(a0, "latent_v_dim", 0) or 0)       # MLA: one latent plane, the value is a column view of it
        per_page = KVPool.bytes_per_page(self.local_page_size, n_layers, kv_heads, kv_dim, kv_dtype, 1 if latent_v else 2)
        if self.swa is not None:
            per_page = int(per_page * eff_layers / n_layers)
        self.max_context_len = a.context_length or cfg.max_position_embeddings
        if self.is_cuda:
            free, total = torch.cuda.mem_get_info(self.device)
            budget = int(total * a.mem_fraction_static) - (total - free)
            # activation head-room for the largest prefill chunk (QKV + MLP intermediates), logits and graphs
            chunk = a.chunked_prefill_size or a.max_prefill_tokens
            act = chunk * (cfg.hidden_size * 6 + (cfg.intermediate_size // self.tp.world_size) * 3 * 2) * 2
            budget -= act
            # a model without attention layers (pure Mamba) has no KV bytes per page: the page pool is then only bookkeeping for the scheduler
            num_pages = max(16, budget // per_page) if per_page > 0 else (a.max_total_tokens or (1 << 20)) // self.page_size + 1
        else:
            num_pages = 4096 // self.page_size + 1
        if a.max_total_tokens:
            num_pages = min(num_pages, a.max_total_tokens // self.page_size + 1)
        if self.tp.world_size > 1:   # all ranks must agree (reference: min over TP ranks)
            t = torch.tensor([num_pages], dtype=torch.int64, device=self.device)
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MIN, group=self.tp.device_group)
            num_pages = int(t.item())
        self.num_pages = int(num_pages)
        if self.swa is None:
            self.kv_pool = KVPool(self.num_pages, self.local_page_size, n_layers, kv_heads, kv_dim, kv_dtype, str(self.device), latent_v)
        else:
            from ..mem_cache.memory_pool import HybridKVPool

            n_swa = max(8, int(self.num_pages * self.swa["ratio"]))
            self.swa["num_pages"] = n_swa
            self.kv_pool = HybridKVPool(KVPool(self.num_pages, self.local_page_size, n_layers - len(sliding), kv_heads, kv_dim, kv_dtype, str(self.device)),
                                        KVPool(n_swa, self.local_page_size, len(sliding), kv_heads, kv_dim, kv_dtype, str(self.device)), sliding)
            logger.info("sliding-window pool: %d of %d layers, %d pages (window %d) next to %d full pages", len(sliding), n_layers, n_swa,
                        self.swa["window"], self.num_pages)
        tokens = (self.num_pages - 1) * self.page_size
        self.max_running = a.max_running_requests or max(1, min(4096 if self.is_cuda else 64, tokens // 64))
        max_pages_per_req = (self.max_context_len + self.page_size - 1) // self.page_size + 1
        # hybrid sliding-window models: ONE device table, columns [0, P) full-attention pages, [P, 2P) sliding pages.  Sized HERE, before the
        # decode graphs are captured: the graphs hold this tensor's address and the column offset
        self.swa_pt_offset = max_pages_per_req if self.swa is not None else 0
        self.page_table = torch.zeros(self.max_running + 1, max_pages_per_req * (2 if self.swa is not None else 1), dtype=torch.int32,
                                      device=self.device)
        # recurrent (SSM) state of hybrid / Mamba decoders, indexed by request slot like the page table (models/mamba.py)
        # (+ --mamba-cache-slots snapshot rows after the request rows: mem_cache/mamba_radix.py)
        self.mamba_snapshot_rows: List[int] = []
        n_snap = int(getattr(a, "mamba_cache_slots", 0) or 0) if getattr(self.model, "supports_mamba_snapshots", False) else 0
        self.kv_pool.mamba = self.model.make_mamba_pool(self.max_running + 1 + n_snap, self.device) if hasattr(self.model, "make_mamba_pool") else None
        if self.kv_pool.mamba is not None and n_snap:
            self.mamba_snapshot_rows = list(range(self.max_running + 1, self.max_running + 1 + n_snap))
        self.next_token_buf = torch.zeros(self.max_running + 1, dtype=torch.int32, device=self.device)
        if getattr(self.model, "is_mrope", False):
            self.model.mrope_delta = torch.zeros(self.max_running + 1, dtype=torch.int32, device=self.device)
        logger.info("KV pool: %d pages x %d tokens (%.2f GB), max_running=%d, weights %.2f GB", self.num_pages, self.page_size,
                    self.num_pages * per_page / 2 ** 30, self.max_running, self.weight_bytes / 2 ** 30)

    def get_memory_info(self) -> Dict[str, int]:
        info = {"num_pages": self.num_pages, "max_context_len": self.max_context_len, "max_running": self.max_running}
        if self.mamba_snapshot_rows:
            info["mamba"] = {"rows": list(self.mamba_snapshot_rows), "granularity": int(getattr(self.model, "snapshot_granularity", 64))}
        if getattr(self, "swa", None) is not None:
            info["swa"] = {"num_pages": self.swa["num_pages"], "window": self.swa["window"]}
        return info

    def attach_kv_manager(self, kv: Any) -> None:
        self.kv_manager = kv
        if getattr(self, "swa", None) is not None:
            # captured graphs already point at the table and its sliding half: the manager's geometry has to match it exactly
            assert kv.page_table.shape[1] == self.swa_pt_offset and self.page_table.shape[1] == 2 * self.swa_pt_offset, \
                (kv.page_table.shape, self.page_table.shape, self.swa_pt_offset)
            return
        assert kv.page_table.shape[1] <= self.page_table.shape[1] + 1
        if kv.page_table.shape[1] != self.page_table.shape[1]:
            if self.graphs:
                raise RuntimeError("page-table geometry changed after CUDA-graph capture")
            self.page_table = torch.zeros(self.max_running + 1, kv.page_table.shape[1], dtype=torch.int32, device=self.device)

    def copy_mamba_state(self, src: int, dst: int) -> None:
        """recurrent state row ``src`` -> row ``dst`` in every layer (prefix-cache snapshot restore), ordered on the runner's stream"""
        pool = self.kv_pool.mamba
        ctx = torch.cuda.stream(self.stream) if self.is_cuda else _null_ctx()
        with ctx:
            pool.conv[:, dst].copy_(pool.conv[:, src])
            pool.ssm[:, dst].copy_(pool.ssm[:, src])

    def get_eos_token_ids(self) -> List[int]:
        return self.cfg.eos_token_ids

    def supports_overlap(self) -> bool:
        return True

    def flush(self) -> None:
        self.page_table.zero_()

    # ------------------------------------------------------------------ prefill context parallelism (layers/cp.py)
    def _cp_eligible(self, sb: ScheduleBatch, fb: ForwardBatch) -> bool:
        return (fb.cu_q is not None and fb.input_ids is not None and fb.num_tokens >= self.cp_min_tokens and fb.mm_embeds is None
                and self.aux_layer_ids is None
                and self.lora_manager is None and sb.spec_info is None and fb.custom_mask is None and self.pp.world_size == 1
                and not getattr(self.model, "is_mrope", False))

    def _forward_cp(self, sb: ScheduleBatch, fb: ForwardBatch) -> torch.Tensor:
        """run the decoder stack on this attention group's zig-zag slice of the rows; returns the hidden states of ALL rows, original order"""
        import dataclasses as _dc

        from ..layers.cp import build_meta
        from ..layers.dp_attention import get_cp_ctx

        st_rank = self.cp_rank
        meta = build_meta(st_rank, self.cp.world_size, fb.seq_lens_cpu, fb.q_lens_cpu, np.asarray(sb.req_slots), fb.out_cache_loc)
        rows = meta.rows
        lfb = _dc.replace(fb, batch_size=int(meta.v_slots.numel()), num_tokens=int(rows.numel()), input_ids=fb.input_ids[rows],
                          positions=fb.positions[rows], out_cache_loc=fb.out_cache_loc[rows], req_slots=meta.v_slots,
                          seq_lens=meta.v_seq_lens, cu_q=meta.v_cu_q, q_lens=meta.v_q_lens, max_q_len=meta.v_max_q, cp=meta, num_splits=1)
        ctx = get_cp_ctx()
        ctx.begin()
        ctx.cp_active = True
        try:
            hidden = self.model(lfb)
            full = ctx.gather(hidden, "tokens")
        finally:
            ctx.cp_active = False
        self.cp_forwards += 1
        return full[meta.inv_perm]

    # ------------------------------------------------------------------ two-batch overlap (reference batch_overlap/two_batch_overlap.py)
    def _sp_eligible(self, fb: ForwardBatch) -> bool:
        """sequence parallel prefill: plain Llama-shaped dense decoders (every layer runs ``LlamaDecoderLayer.forward`` with a dense MLP)"""
        from ..layers.dp_attention import get_dp_ctx
        from ..models.llama import LlamaDecoderLayer, LlamaMLP

        m = getattr(self.model, "model", None)
        ok = getattr(self, "_sp_static", None)
        if ok is None:
            ok = (self.tp.world_size > 1 and self.pp.world_size == 1 and m is not None and hasattr(m, "embed") and self.lora_manager is None
                  and self.cp is None and self.dcp is None and type(self.model).forward is _LLAMA_FORWARD
                  and all(type(l).forward is LlamaDecoderLayer.forward and isinstance(l.mlp, LlamaMLP) for l in m.layers))
            self._sp_static = ok
        return ok and get_dp_ctx() is None and fb.capture_aux is None and fb.num_tokens >= envs.SGLB_SP_MIN_TOKENS.get()

    def _tbo_static_ok(self) -> bool:
        m = getattr(self.model, "model", None)
        return (m is not None and hasattr(m, "embed") and self.pp.world_size == 1 and self.lora_manager is None and self.dcp is None
                and self.aux_layer_ids is None
                and self.cp is None and not getattr(self.model, "is_mrope", False) and type(self.model).forward is _LLAMA_FORWARD)

    def _tbo_eligible(self, sb: ScheduleBatch, fb: ForwardBatch) -> bool:
        from ..layers.dp_attention import get_dp_ctx

        if not self._tbo_static_ok():
            return False
        if get_dp_ctx() is not None:
            # DP-attention: every group must issue the same collective sequence -> the decision may only depend on static facts; a group
            # with fewer than two requests runs an EMPTY second micro-batch (and idle groups two empty ones, see forward_idle)
            return self.cfg.vision is None
        return len(sb.reqs) >= 2 and fb.mm_embeds is None and sb.spec_info is None and fb.custom_mask is None

    def _forward_tbo(self, sb: ScheduleBatch, fb: ForwardBatch) -> torch.Tensor:
        """Split the batch by requests into two micro-batches of near-equal token count and walk the decoder stack layer by layer, alternating
        between them on two streams: while one micro-batch waits for its MoE all-to-all / TP all-reduce the other one's attention and GEMMs
        keep the SMs busy.  On one stream (CPU tier) this is just a different evaluation order, so outputs are identical by construction."""
        import dataclasses as _dc

        from ..layers.dp_attention import get_dp_ctx

        decode = fb.cu_q is None
        B = fb.batch_size
        if decode:
            k = (B + 1) // 2
            cuts = [(0, k, 0, k), (k, B, k, B)]                                   # (req lo, req hi, row lo, row hi)
        elif B < 2:
            T = fb.num_tokens
            cuts = [(0, B, 0, T), (B, B, T, T)]                                   # second micro-batch empty (DP-attention lock-step)
        else:
            cu = np.concatenate([[0], np.cumsum(fb.q_lens_cpu)])
            k = int(np.argmin(np.abs(cu[1:-1] - cu[-1] / 2))) + 1
            cuts = [(0, k, 0, int(cu[k])), (k, B, int(cu[k]), int(cu[-1]))]
        parts = []
        for (r0, r1, t0, t1) in cuts:
            kw = dict(batch_size=r1 - r0, num_tokens=t1 - t0, positions=fb.positions[t0:t1], out_cache_loc=fb.out_cache_loc[t0:t1],
                      req_slots=fb.req_slots[r0:r1], seq_lens=fb.seq_lens[r0:r1], input_ids=None if fb.input_ids is None else fb.input_ids[t0:t1],
                      seq_lens_cpu=None if fb.seq_lens_cpu is None else fb.seq_lens_cpu[r0:r1], num_splits=1, attn_workspace=None)
            if not decode:
                ql = fb.q_lens[r0:r1]
                qc = fb.q_lens_cpu[r0:r1] if fb.q_lens_cpu is not None else np.zeros(0, np.int32)
                sc = fb.seq_lens_cpu[r0:r1] if fb.seq_lens_cpu is not None else np.zeros(0, np.int32)
                kw.update(q_lens=ql, cu_q=(fb.cu_q[r0:r1 + 1] - fb.cu_q[r0]).contiguous(), q_lens_cpu=qc,
                          max_q_len=int(qc.max()) if qc.size else 0, max_kv_len=int(sc.max()) if sc.size else 0)
            parts.append(_dc.replace(fb, **kw))
        m = self.model.model
        ctx = get_dp_ctx()
        streams = None
        if self.is_cuda:
            if getattr(self, "_tbo_streams", None) is None:
                self._tbo_streams = [torch.cuda.Stream(device=self.device), torch.cuda.Stream(device=self.device)]
            streams = self._tbo_streams
            cur = torch.cuda.current_stream()
            for s_ in streams:
                s_.wait_stream(cur)

        def on(i):
            if ctx is not None:
                ctx.mb = i
            return torch.cuda.stream(streams[i]) if streams else _null_ctx()

        st = []
        for i, f in enumerate(parts):
            with on(i):
                st.append((m.embed(f), None))
        for layer in m.layers:
            for i, f in enumerate(parts):
                with on(i):
                    st[i] = layer(st[i][0], st[i][1], f)
        outs = []
        for i, f in enumerate(parts):
            with on(i):
                h, r = st[i]
                h = h.resolve() if hasattr(h, "resolve") else h
                outs.append(m.norm(h) if r is None else m.norm(h, r)[0])
        if streams:
            for s_ in streams:
                torch.cuda.current_stream().wait_stream(s_)
        if ctx is not None:
            ctx.mb = 0
        self.tbo_forwards = getattr(self, "tbo_forwards", 0) + 1
        return torch.cat(outs, 0)

    def _attach_dcp(self, sb: ScheduleBatch, fb: ForwardBatch) -> None:
        if self.dcp is None:
            return
        from ..layers.dcp import build_meta

        decode = sb.forward_mode.is_decode() and fb.cu_q is None
        fb.dcp = build_meta(self.dcp, self.page_size, fb.out_cache_loc, fb.seq_lens, decode, fb.seq_lens_cpu, fb.q_lens_cpu,
                            np.asarray(sb.req_slots) if not decode else None)
        fb.page_size = self.local_page_size
        fb.num_splits = 1

    # ------------------------------------------------------------------ attention split heuristic
    def _pick_splits(self, bs: int, max_kv: int) -> int:
        if not self.is_cuda or getattr(self.args, "enable_deterministic_inference", False):
            return 1          # (deterministic mode: a request's softmax reduction tree must not follow the batch size)
        if not len(self.model.model.layers):
            return 1
        kvh = self.model.model.layers[0].self_attn.attn.num_kv_heads
        ctas = bs * kvh
        tiles = (max_kv + 127) // 128            # the decode kernel works on 128-key tiles, one persistent CTA per SM
        sms = self.num_sms
        if ctas >= sms or tiles <= 2:
            return 1
        return int(max(1, min(tiles, 8, (sms * 2) // max(1, ctas))))

    # ------------------------------------------------------------------ forward
    def forward_batch_generation(self, sb: ScheduleBatch) -> BatchResult:
        self.forward_ct += 1
        if self.is_cuda:
            with torch.cuda.stream(self.stream):
                return self._forward_guarded(sb)
        return self._forward_guarded(sb)

    @torch.no_grad()
    def forward_idle(self) -> None:
        """DP-attention: this group has no batch this iteration but the other groups do — run the same layer sequence on ZERO local
        tokens so that every gather / all-reduce of the step completes (reference scheduler_components/dp_attn.py idle batches)."""
        from ..layers.dp_attention import get_dp_ctx

        ctx = get_dp_ctx()
        if ctx is None:
            return
        self.forward_ct += 1
        ctx.begin()
        dev = self.device
        z = lambda: torch.zeros(0, dtype=torch.int32, device=dev)
        fb = ForwardBatch(mode=ForwardMode.EXTEND, batch_size=0, num_tokens=0, input_ids=z(), positions=z(), out_cache_loc=z(),
                          req_slots=z(), seq_lens=z(), cu_q=torch.zeros(1, dtype=torch.int32, device=dev), q_lens=z(), last_idx=z(),
                          sample_slots=z(), max_q_len=0, max_kv_len=0, page_table=self.page_table, page_size=self.page_size,
                          kv_pool=self.kv_pool, next_token_buf=self.next_token_buf)
        hidden = self._forward_tbo(None, fb) if (self.args.enable_two_batch_overlap and self._tbo_static_ok() and self.cfg.vision is None) \
            else self.model(fb)
        self.model.compute_logits(hidden[:0])        # the LM head gathers sampled rows across groups as well ...
        self.model.compute_logits(hidden[:0], key="logits_all")     # ... and the rows of a prompt-logprob pass

    def _forward_guarded(self, sb: ScheduleBatch) -> BatchResult:
        nvtx = self.is_cuda and envs.SGLB_NVTX.get()
        if nvtx:     # nsys / ncu --nvtx timelines: one range per forward, named by mode and batch shape
            torch.cuda.nvtx.range_push(f"{sb.forward_mode.name} bs={len(sb.reqs)}")
        try:
            return self._forward(sb)
        finally:
            self.set_rs(0)          # the row-sharded residual mode never outlives the forward that armed it
            if nvtx:
                torch.cuda.nvtx.range_pop()

    @torch.no_grad()
    def _forward(self, sb: ScheduleBatch) -> BatchResult:
        from ..layers.dp_attention import get_dp_ctx

        _dp = get_dp_ctx()
        if _dp is not None:
            _dp.begin()
        B = len(sb.reqs)
        decode = sb.forward_mode.is_decode()
        self.step_counter += 1
        # the captured sampler sees only the static tensors: anything that needs per-request host state (penalties, logit bias, active
        # min_new_tokens -> EOS mask, per-request seeds, hidden-state return) takes the eager path (ScheduleBatch.needs_host_state)
        use_graph = decode and self.graphs and not sb.return_logprob and not sb.has_grammar and sb.spec_info is None and \
            B <= self.graph_bs[-1] and not sb.needs_host_state() and (sb.out_cache_loc_swa is None or self.graphs[self.graph_bs[-1]].has_swa)
        if use_graph:
            return self._forward_graph(sb)
        if self.prefill_graphs and not decode and self._prefill_graph_eligible(sb):
            return self._forward_prefill_graph(sb)
        fb = self.packer.build(sb, self.page_table, self.page_size, self.kv_pool, self.next_token_buf, ops.apply_pt_delta)
        fb.swa_pt_offset = getattr(self, "swa_pt_offset", 0)
        self.h2d_bytes += fb.h2d_bytes
        self.set_rs(fb.num_tokens if sb.spec_info is None and not self.args.enable_two_batch_overlap else 0)
        fb.num_splits = self._pick_splits(B, fb.max_kv_len) if decode else 1
        self._attach_dcp(sb, fb)
        self._attach_mm(sb, fb)
        if self.lora_manager is not None:
            self.lora_manager.prepare_batch(sb, fb)
        if not self.pp.is_first:
            fb.pp_hidden = self._pp_recv(fb.num_tokens)
        if self.aux_layer_ids is not None and not decode:
            fb.capture_aux = self.aux_layer_ids
        if self.cp is not None and not decode and self._cp_eligible(sb, fb):
            hidden = self._forward_cp(sb, fb)
        elif self.args.enable_two_batch_overlap and self._tbo_eligible(sb, fb):
            hidden = self._forward_tbo(sb, fb)
        elif self.args.enable_sp and not decode and self._sp_eligible(fb):
            from ..layers.sp import sp_context

            sp = sp_context()
            sp.begin(self.tp, fb.num_tokens)
            try:
                hidden = self.model(fb)
            finally:
                sp.end()
        else:
            hidden = self.model(fb)
        if self.hidden_tap is not None and not decode and self.pp.is_last:
            self.hidden_tap(sb, hidden if fb.aux_hidden is None else fb.aux_hidden)      # EAGLE: target features of the prefilled positions
        res = BatchResult()
        res.batch_size = B
        if not self.pp.is_last:
            self._pp_send(hidden)
            return self._pp_result(sb, res)
        samp_rows = np.arange(B) if decode else np.flatnonzero(sb.sample_mask)
        res.sample_rows = samp_rows
        if _dp is not None:
            # DP-attention: the LM head is vocab-parallel over ALL groups -> every group calls it exactly twice per forward (sampled
            # rows, then the rows of a prompt-logprob pass), with zero rows when it has nothing to contribute
            self._dp_all_logits = None
            need_all = (not decode) and any(r.return_logprob and r.logprob_start_len >= 0 for r in sb.reqs)
            if samp_rows.size == 0:
                self.model.compute_logits(hidden[:0])
                al = self.model.compute_logits(hidden if need_all else hidden[:0], key="logits_all")
                res.next_token_ids = np.full(B, -1, dtype=np.int64)
                if need_all:
                    self._dp_all_logits = al
                    self._input_logprobs(sb, hidden, res)
                return res
        if samp_rows.size == 0:
            res.next_token_ids = np.full(B, -1, dtype=np.int64)
            if sb.return_logprob and _dp is None:
                self._input_logprobs(sb, hidden, res)
            return res
        h_last = hidden if decode else ops.gather_rows(hidden, fb.last_idx)
        if getattr(self.model, "is_embedding", False) or any(r.is_embedding for r in sb.reqs):
            emb = self.model.pool(h_last) if hasattr(self.model, "pool") else torch.nn.functional.normalize(h_last.float(), dim=-1)
            embs = emb.cpu().tolist()
            res.embeddings = [None] * B
            for j, i in enumerate(samp_rows.tolist()):
                res.embeddings[i] = embs[j]
            res.next_token_ids = np.full(B, -1, dtype=np.int64)
            return res
        logits = self.model.compute_logits(h_last)
        if _dp is not None:
            al = self.model.compute_logits(hidden if need_all else hidden[:0], key="logits_all")
            self._dp_all_logits = al if need_all else None
        if any(r.return_hidden_states for r in sb.reqs):
            hs = h_last.float().cpu().tolist()
            res.hidden_states = [None] * B
            for j, i in enumerate(samp_rows.tolist()):
                res.hidden_states[i] = hs[j]
        reqs = [sb.reqs[i] for i in samp_rows.tolist()]
        logits = self._preprocess_logits(logits, reqs, fb)
        lp_out = torch.empty(len(reqs), dtype=torch.float32, device=self.device) if sb.return_logprob else None
        toks = self.sample(logits, fb, lp_out)
        if self.pp.world_size > 1:      # earlier stages need the sampled ids as their next decode input (see _pp_result)
            t32 = toks.to(torch.int32).contiguous()
            self.pp.broadcast(t32, src=self.pp.world_size - 1)
        if sb.return_logprob:
            self._extra_logprobs(sb, fb, hidden, logits, res, samp_rows)
        self._finish_result(res, toks, lp_out)
        return res

    @torch.no_grad()
    # ---- expert placement (eplb/placement.py); reached via Engine.collective_rpc so every rank moves together
    def build_expert_backup(self):
        from ..eplb import placement

        return placement.build_backup(self)

    def eplb_rebalance(self, redundant_slots_per_rank: int = 0, load=None):
        from ..eplb import placement

        return placement.rebalance(self, redundant_slots_per_rank, load)

    def elastic_ep_recover(self, dead_ranks):
        from ..eplb import placement

        return placement.recover(self, dead_ranks)

    def forward_verify(self, sb: ScheduleBatch) -> np.ndarray:
        """Speculative verification: ragged extend over [last_token, drafts...] of every request; returns the target
        model's sample at EVERY fed position (int64 [T]).  Synchronous (speculative decoding disables overlap)."""
        self.forward_ct += 1
        self.step_counter += 1
        ctx = torch.cuda.stream(self.stream) if self.is_cuda else _null_ctx()
        with ctx:
            fb = self.packer.build(sb, self.page_table, self.page_size, self.kv_pool, self.next_token_buf, ops.apply_pt_delta)
            self.h2d_bytes += fb.h2d_bytes
            self._attach_dcp(sb, fb)
            cm = (sb.spec_info or {}).get("custom_mask") if isinstance(sb.spec_info, dict) else None
            if cm is not None:       # token-tree verification: every node attends to the prefix and its own ancestors only
                assert self.dcp is None, "tree drafts + DCP: not supported"
                fb.custom_mask = torch.from_numpy(cm).to(self.device)
            ta = (sb.spec_info or {}).get("tree_anc") if isinstance(sb.spec_info, dict) else None
            if ta is not None and (cm is None or self.supports_tree_anc()):
                assert self.dcp is None, "tree drafts + DCP: not supported"
                fb.tree_anc = torch.from_numpy(ta).to(self.device)
            if self.aux_layer_ids is not None:
                fb.capture_aux = self.aux_layer_ids
            hidden = self.model(fb)
            self.last_verify_hidden = (hidden if fb.aux_hidden is None else fb.aux_hidden) if self.hidden_tap is not None else None
            logits = self.model.compute_logits(hidden)          # [T, V]
            self.last_verify_logits = logits if isinstance(sb.spec_info, dict) and sb.spec_info.get("return_logits") else None
            reps = torch.from_numpy(sb.q_lens.astype(np.int64)).to(self.device)
            s = fb.sampling
            rep = lambda t: torch.repeat_interleave(t, reps)
            g_rows = [i for i, r in enumerate(sb.reqs) if r.grammar is not None]
            if g_rows:
                cu = np.concatenate([[0], np.cumsum(sb.q_lens)])
                V = logits.shape[1]
                bm = np.full((len(g_rows), (V + 31) // 32), -1, dtype=np.int32)
                for k, i in enumerate(g_rows):
                    sb.reqs[i].grammar.fill_vocab_mask(bm[k], V)
                rows = torch.tensor([int(cu[i]) for i in g_rows], dtype=torch.int32, device=self.device)
                ops.apply_token_bitmask(logits, torch.from_numpy(bm).to(self.device), rows)
            toks = ops.sample(logits, rep(s.temperature), rep(s.top_k), rep(s.top_p), rep(s.min_p), None, None, None, None,
                              global_seed=self.args.random_seed, step=self.step_counter)
            out = toks.cpu().numpy().astype(np.int64)
        self.d2h_bytes += out.size * 4
        return out

    def sample(self, logits: torch.Tensor, fb: ForwardBatch, lp_out: Optional[torch.Tensor]) -> torch.Tensor:
        s = fb.sampling
        if self.args.sampling_backend == "token_oracle" and self.token_oracle is not None:
            return self.token_oracle(fb, logits)
        return ops.sample(logits, s.temperature, s.top_k, s.top_p, s.min_p, s.seeds, fb.sample_slots, fb.next_token_buf,
                          lp_out, global_seed=self.args.random_seed, step=self.step_counter)

    token_oracle = None
    last_verify_logits = None

    def _finish_result(self, res: BatchResult, toks: torch.Tensor, lp: Optional[torch.Tensor]) -> None:
        n = toks.numel()
        if self.is_cuda:
            self._ring_i += 1
            th, lh = self._res_ring[self._ring_i % len(self._res_ring)]
            res._tok_host = th[:n]
            res._tok_host.copy_(toks, non_blocking=True)
            if lp is not None:
                res._lp_host = lh[:n]
                res._lp_host.copy_(lp, non_blocking=True)
            res.event = torch.cuda.Event()
            res.event.record(torch.cuda.current_stream())
        else:
            res._tok_host = toks.to(torch.int32).cpu()
            res._lp_host = lp.cpu() if lp is not None else None
        res.d2h_bytes = n * 4 + (n * 4 if lp is not None else 0)
        self.d2h_bytes += res.d2h_bytes

    # ---- logits post-processing (grammar mask, penalties, logit bias) — reference sampling_batch_info.py:240,301
    def _preprocess_logits(self, logits: torch.Tensor, reqs: List[Any], fb: ForwardBatch) -> torch.Tensor:
        need_f32 = fb.sampling.has_penalty
        if need_f32:
            logits = logits.float()
            V = logits.shape[1]
            for j, r in enumerate(reqs):
                sp = r.sampling_params
                if sp.logit_bias:
                    idx = torch.tensor([int(k) for k in sp.logit_bias], device=self.device)
                    logits[j, idx] += torch.tensor(list(sp.logit_bias.values()), device=self.device, dtype=logits.dtype)
                if sp.frequency_penalty or sp.presence_penalty or sp.repetition_penalty != 1.0:
                    counts = torch.zeros(1, V, dtype=torch.int32, device=self.device)
                    if r.output_ids:
                        out = torch.tensor(r.output_ids, device=self.device)
                        counts[0].index_add_(0, out, torch.ones_like(out, dtype=torch.int32))
                    pm = torch.zeros(1, V, dtype=torch.uint8, device=self.device)
                    if sp.repetition_penalty != 1.0:
                        pm[0, torch.tensor(r.origin_input_ids, device=self.device)] = 1
                    row = logits[j:j + 1]
                    ops.apply_penalties(row, counts, pm, torch.tensor([sp.frequency_penalty], device=self.device),
                                        torch.tensor([sp.presence_penalty], device=self.device),
                                        torch.tensor([sp.repetition_penalty], device=self.device))
        g_rows = [j for j, r in enumerate(reqs) if r.grammar is not None]
        if g_rows:
            V = logits.shape[1]
            words = (V + 31) // 32
            bm = np.full((len(g_rows), words), -1, dtype=np.int32)
            for k, j in enumerate(g_rows):
                reqs[j].grammar.fill_vocab_mask(bm[k], V)
            bm_t = torch.from_numpy(bm).to(self.device)
            ops.apply_token_bitmask(logits, bm_t, torch.tensor(g_rows, dtype=torch.int32, device=self.device))
        for j, r in enumerate(reqs):
            if len(r.output_ids) < r.sampling_params.min_new_tokens:
                stops = list(r.eos_token_ids) + list(r.sampling_params.stop_token_ids)
                if stops:
                    logits[j, torch.tensor(stops, device=self.device)] = float("-inf")
        return logits

    def _extra_logprobs(self, sb: ScheduleBatch, fb: ForwardBatch, hidden, logits, res: BatchResult, samp_rows) -> None:
        """top-k logprobs of the sampled position and prompt (input) logprobs — library ops, off the hot path"""
        lsm = torch.log_softmax(logits.float(), dim=-1)
        res.top_logprobs = [None] * len(sb.reqs)
        for j, i in enumerate(samp_rows.tolist()):
            k = sb.reqs[i].top_logprobs_num
            if k > 0:
                v, idx = lsm[j].topk(k)
                res.top_logprobs[i] = list(zip(v.tolist(), idx.tolist()))
        self._input_logprobs(sb, hidden, res)

    def _input_logprobs(self, sb: ScheduleBatch, hidden, res: BatchResult) -> None:
        """prompt (input) logprobs of THIS chunk: for every prompt position ``pos`` in ``(max(chunk_begin, logprob_start_len), chunk_end]``
        the logprob of token ``pos`` given everything before it (row ``pos - 1`` of the chunk) — the last one belongs to the first token of
        the NEXT chunk, so chunked prefill loses nothing.  The scheduler concatenates the chunks (reference logprob_processor input path)."""
        if not (sb.forward_mode.is_extend() and any(r.return_logprob and r.logprob_start_len >= 0 for r in sb.reqs)):
            return
        pre = getattr(self, "_dp_all_logits", None)      # DP-attention computed it collectively
        all_logits = torch.log_softmax((pre if pre is not None else self.model.compute_logits(hidden)).float(), dim=-1)
        res.input_logprobs = [None] * len(sb.reqs)
        cu = np.concatenate([[0], np.cumsum(sb.q_lens)])
        for i, r in enumerate(sb.reqs):
            if not (r.return_logprob and r.logprob_start_len >= 0):
                continue
            cb, ce = int(sb.prefix_lens[i]), int(sb.prefix_lens[i] + sb.q_lens[i])
            full = r.origin_input_ids
            lo, hi = max(cb, r.logprob_start_len) + 1, min(ce, len(full) - 1)
            if hi < lo:
                res.input_logprobs[i] = []
                continue
            rows = torch.arange(lo - 1 - cb, hi - cb, device=self.device) + int(cu[i])
            ids = torch.tensor(full[lo: hi + 1], dtype=torch.int64, device=self.device)
            lp = all_logits[rows].gather(1, ids[:, None]).squeeze(1).tolist()
            res.input_logprobs[i] = [(float(a), int(b)) for a, b in zip(lp, ids.tolist())]

    def _attach_mm(self, sb: ScheduleBatch, fb: ForwardBatch) -> None:
        if not sb.forward_mode.is_extend():
            return
        if getattr(self.model, "is_mrope", False):
            self._attach_mrope(sb, fb)
        if hasattr(self.model, "prepare_batch"):       # encoder-decoder models: run the encoder of requests that start here (models/whisper.py)
            self.model.prepare_batch(sb, fb, self)
            return
        if not any(r.mm_inputs for r in sb.reqs):
            return
        from ..multimodal.embed import build_mm_embeds

        fb.mm_embeds = build_mm_embeds(self.model, sb, self.device, self.dtype)

    def _attach_mrope(self, sb: ScheduleBatch, fb: ForwardBatch) -> None:
        """M-RoPE (Qwen2-VL): 3-axis rotary positions of this chunk's tokens + the per-request decode delta (models/qwen2_vl.py)"""
        m = self.model
        if m.mrope_delta is None:
            m.mrope_delta = torch.zeros(self.max_running + 1, dtype=torch.int32, device=self.device)
        cols, deltas = [], []
        for i, r in enumerate(sb.reqs):
            s, e = int(sb.prefix_lens[i]), int(sb.prefix_lens[i] + sb.q_lens[i])
            mm = r.mm_inputs or {}
            tab, d = mm.get("mrope_positions"), int(mm.get("mrope_delta", 0))
            p = np.arange(s, e, dtype=np.int64)
            if tab is None:
                cols.append(np.stack([p, p, p]))
            else:
                tab = np.asarray(tab, dtype=np.int64)
                n = tab.shape[1]
                cols.append(np.concatenate([tab[:, s:min(e, n)], np.stack([p[p >= n] + d] * 3)], 1))
            deltas.append(d)
        fb.mrope_positions = torch.from_numpy(np.concatenate(cols, 1)).to(self.device)
        slots = torch.from_numpy(np.asarray(sb.req_slots, dtype=np.int64)).to(self.device)
        m.mrope_delta[slots] = torch.tensor(deltas, dtype=torch.int32, device=self.device)

    # ------------------------------------------------------------------ pipeline parallel hand-off
    def _pp_recv(self, T: int):
        H = self.cfg.hidden_size
        h = torch.empty(T, H, dtype=self.dtype, device=self.device)
        r = torch.empty(T, H, dtype=self.dtype, device=self.device)
        self.pp.recv(h, self.pp.rank - 1)
        self.pp.recv(r, self.pp.rank - 1)
        return h, r

    def _pp_send(self, hr) -> None:
        h, r = hr
        self.pp.send(h.contiguous(), self.pp.rank + 1)
        self.pp.send(r.contiguous(), self.pp.rank + 1)

    def _pp_result(self, sb: ScheduleBatch, res: BatchResult) -> BatchResult:
        """non-last stages learn the sampled ids from the last stage (they need them for the next decode input)"""
        B = len(sb.reqs)
        decode = sb.forward_mode.is_decode()
        rows = np.arange(B) if decode else np.flatnonzero(sb.sample_mask)
        res.sample_rows = rows
        toks = torch.empty(max(1, rows.size), dtype=torch.int32, device=self.device)
        if rows.size:
            self.pp.broadcast(toks, src=self.pp.world_size - 1)
            slots = torch.from_numpy(sb.req_slots[rows].astype(np.int64)).to(self.device)
            self.next_token_buf[slots] = toks[: rows.size]
        self._finish_result(res, toks[: rows.size], None)
        return res

    # ------------------------------------------------------------------ CUDA graph decode path
    def _capture_graphs(self) -> None:
        a = self.args
        max_bs = min(a.cuda_graph_max_bs, self.max_running)
        bs_list = a.cuda_graph_bs or sorted({b for b in ([1, 2, 4, 8, 12, 16, 24, 32, 48, 64, 96, 128, 160, 192, 224, 256] +
                                                         list(range(320, 4097, 64))) if b <= max_bs} | {max_bs})
        self.graph_bs = sorted(bs_list)
        t0 = time.perf_counter()
        pool = None
        kvh = self.model.model.layers[0].self_attn.attn.num_kv_heads
        shard_greedy = self.greedy_head() is not None
        for bs in reversed(self.graph_bs):
            g = _DecodeGraph(self, bs)
            g.capture(pool)
            pool = g.graph.pool()
            self.graphs[bs] = g
            if shard_greedy:
                gg = _DecodeGraph(self, bs, greedy=True)
                gg.capture(pool)
                self.graphs_greedy[bs] = gg
        logger.info("captured %d decode graphs (bs %s..%s) in %.1fs", len(self.graphs), self.graph_bs[0], self.graph_bs[-1],
                    time.perf_counter() - t0)
        self._capture_prefill_graphs(pool)

    # ------------------------------------------------------------------ prefill graphs
    def _capture_prefill_graphs(self, pool) -> None:
        """Token-bucketed graphs of a whole ragged extend step (embed -> layers -> last-row gather -> LM head -> sampler).  Reference:
        model_executor/runner/prefill_cuda_graph_runner.py (bucketed by token count, padded).  Every per-step quantity lives in ONE static
        int32 buffer; padded requests have q_len 0 (their attention work items exit at once), padded tokens write their K/V to slot 0."""
        a = self.args
        cap_t = int(getattr(a, "prefill_cuda_graph_max_tokens", 0) or 0)
        if cap_t <= 0 or a.disable_cuda_graph or not self.is_cuda:
            return
        if self.pp.world_size > 1 or self.cp is not None or a.enable_sp or a.enable_two_batch_overlap or getattr(self, "dcp", None) is not None:
            return
        from ..layers.dp_attention import get_dp_ctx

        if get_dp_ctx() is not None or getattr(self.model, "is_embedding", False) or not hasattr(self.model, "compute_logits"):
            return
        buckets = [t for t in (64, 128, 192, 256, 384, 512, 768, 1024, 1536, 2048, 3072, 4096) if t <= cap_t]
        if not buckets:
            return
        t0 = time.perf_counter()
        for T in reversed(buckets):
            g = _PrefillGraph(self, T)
            g.capture(pool)
            pool = g.graph.pool()
            self.prefill_graphs[T] = g
        self.prefill_graph_tokens = sorted(self.prefill_graphs)
        logger.info("captured %d prefill graphs (%d..%d tokens) in %.1fs", len(buckets), buckets[0], buckets[-1], time.perf_counter() - t0)

    def _prefill_graph_eligible(self, sb: ScheduleBatch) -> bool:
        if sb.return_logprob or sb.has_grammar or sb.spec_info is not None or sb.needs_host_state() or sb.out_cache_loc_swa is not None:
            return False
        if self.lora_manager is not None or self.hidden_tap is not None or self.aux_layer_ids is not None:
            return False
        T = int(sb.q_lens.sum())
        if T > self.prefill_graph_tokens[-1]:
            return False
        import bisect

        g = self.prefill_graphs[self.prefill_graph_tokens[bisect.bisect_left(self.prefill_graph_tokens, T)]]
        if len(sb.reqs) > g.bcap:
            return False
        for r in sb.reqs:
            if r.is_embedding or r.return_hidden_states or getattr(r, "mm_inputs", None) is not None:
                return False
        return True

    def _forward_prefill_graph(self, sb: ScheduleBatch) -> BatchResult:
        import bisect

        B = len(sb.reqs)
        T = int(sb.q_lens.sum())
        g = self.prefill_graphs[self.prefill_graph_tokens[bisect.bisect_left(self.prefill_graph_tokens, T)]]
        nd = g.load(sb, self.step_counter)
        if nd:
            d = g.delta_views(nd)
            ops.apply_pt_delta(self.page_table, d[0], d[1], d[2])
        g.graph.replay()
        self.graph_launches += g.num_kernels
        self.prefill_graph_hits += 1
        res = BatchResult()
        res.batch_size = B
        samp = np.flatnonzero(sb.sample_mask)
        res.sample_rows = samp
        if samp.size == 0:
            res.next_token_ids = np.full(B, -1, dtype=np.int64)
            return res
        toks = g.out_tokens[:B]
        if samp.size != B:
            toks = toks[torch.from_numpy(samp).to(self.device, non_blocking=True)]
        self._finish_result(res, toks, None)
        return res

    # ------------------------------------------------------------------ row-sharded residual (large-T tensor parallel steps)
    def _rs_capable(self) -> bool:
        """may this model run with a row-sharded residual stream (NvlComm.rs_add_rmsnorm_ag)?  Only when EVERY residual update is a fused
        consume_norm of a row-parallel projection: the plain dense Llama-family stack, no PP / SP / DP-attention / aux-hidden capture."""
        c = getattr(self, "_rs_capable_cached", None)
        if c is None:
            from ..layers.dp_attention import get_dp_ctx
            from ..layers.quantization import UnquantizedMethod
            from ..models.llama import LlamaDecoderLayer, LlamaForCausalLM, LlamaMLP

            nvl = self.tp.nvl
            c = bool(nvl is not None and self.tp.world_size > 2 and self.pp.world_size == 1 and not self.args.enable_sp
                     and get_dp_ctx() is None and self.cp is None and getattr(self, "dcp", None) is None
                     and type(self.model) is LlamaForCausalLM
                     and all(type(l) is LlamaDecoderLayer and type(l.mlp) is LlamaMLP
                             and isinstance(getattr(l.mlp.down_proj, "quant", None), UnquantizedMethod)
                             and isinstance(getattr(l.self_attn.o_proj, "quant", None), UnquantizedMethod) for l in self.model.model.layers))
            self._rs_capable_cached = c
        return c

    def set_rs(self, num_tokens: int) -> None:
        """arm (or disarm, num_tokens = 0) the reduce-scatter path for the forward that follows"""
        nvl = self.tp.nvl
        if nvl is None:
            return
        nvl.rs_active = bool(num_tokens > 0 and self.aux_layer_ids is None and self.lora_manager is None and self._rs_capable()
                             and nvl.wants_rs_bytes(num_tokens * self.cfg.hidden_size * 2))

    def supports_tree_anc(self) -> bool:
        """can token-tree verification run on the tcgen05 attention with per-node ancestor bits (attention_tc.cu ``tree_anc``) instead of a
        dense custom mask on the reference path?"""
        c = getattr(self, "_tree_anc_ok", None)
        if c is None:
            layers = getattr(getattr(self.model, "model", None), "layers", None) or []
            attn = [getattr(getattr(l, "self_attn", None), "attn", None) for l in layers]
            c = bool(self.is_cuda and ops.load_native() and ops.attn_impl() == 0 and attn and all(
                a is not None and a.head_dim == 128 and 128 % (a.num_heads // a.num_kv_heads) == 0 and a.window <= 0 and not a.chunk
                and a.sinks is None for a in attn) and self.kv_pool.dtype == torch.bfloat16 and self.dcp is None)
            self._tree_anc_ok = c
        return c

    def greedy_head(self):
        """the vocab-sharded head when all-greedy decode batches may use the vocab-parallel argmax (TP > 1, plain LM head, no DP-attention
        row exchange); None otherwise"""
        from ..environ import envs
        from ..layers.dp_attention import get_dp_ctx
        from ..layers.linear import VocabParallelEmbedding
        from ..models.llama import LlamaForCausalLM

        if self.tp.world_size <= 1 or not envs.SGLB_TP_GREEDY_SHARD.get() or get_dp_ctx() is not None or not self.pp.is_last:
            return None
        m = self.model
        if getattr(type(m), "compute_logits", None) is not LlamaForCausalLM.compute_logits:
            return None
        head = m.lm_head if getattr(m, "lm_head", None) is not None else getattr(getattr(m, "model", None), "embed_tokens", None)
        return head if isinstance(head, VocabParallelEmbedding) else None

    def recapture_graphs(self) -> None:
        """components installed after construction that change what a decode step launches (the LoRA manager wraps every projection with
        its SGMV kernels) call this so the captured graphs contain them"""
        if not self.graphs:
            return
        self.graphs.clear()
        self.graphs_greedy.clear()
        self.prefill_graphs.clear()
        self.prefill_graph_tokens = []
        self.graph_bs = []
        with torch.cuda.stream(self.stream):
            self._capture_graphs()
        self.stream.synchronize()

    def _forward_graph(self, sb: ScheduleBatch) -> BatchResult:
        B = len(sb.reqs)
        import bisect

        bs = self.graph_bs[bisect.bisect_left(self.graph_bs, B)]
        s4 = sampling_rows(sb.reqs)                  # [4, B] float32: temperature, top_p, min_p, top_k (cached per request)
        g = self.graphs_greedy[bs] if (self.graphs_greedy and float(s4[3].max()) <= 1.0) else self.graphs[bs]
        # page-table deltas are applied eagerly (their count is unbounded), then one graph launch
        nd = 0 if sb.pt_rows is None else sb.pt_rows.size
        g.load(sb, self.step_counter, self.args.random_seed, s4)
        if self.lora_manager is not None:
            self.lora_manager.prepare_batch(sb)          # refresh the static tile table the captured SGMV kernels read
        if nd:
            d = g.delta_views(nd)
            ops.apply_pt_delta(self.page_table, d[0], d[1], d[2])
        g.graph.replay()
        self.graph_launches += g.num_kernels
        res = BatchResult()
        res.batch_size = B
        res.sample_rows = np.arange(B)
        self._finish_result(res, g.out_tokens[:B], None)
        return res

    # ------------------------------------------------------------------ control plane (weights, profiling, LoRA)
    def handle_control(self, msg: Any) -> Any:
        from . import control

        return control.handle(self, msg)


def sampling_rows(reqs) -> np.ndarray:
    """[4, B] float32 rows (temperature, top_p, min_p, top_k) from a per-request cached tuple: one list comprehension per step instead of four
    attribute-chasing ones (host time matters: at TP=8 a decode step is a few ms on the device)"""
    rows = []
    for r in reqs:
        t = r.samp4
        if t is None:
            p = r.sampling_params
            t = r.samp4 = (float(p.temperature), float(p.top_p), float(p.min_p), float(min(p.top_k, 1 << 30)))
        rows.append(t)
    return np.ascontiguousarray(np.asarray(rows, dtype=np.float32).reshape(-1, 4).T) if rows else np.zeros((4, 0), dtype=np.float32)


class _PrefillGraph:
    """One captured ragged-extend step for a token bucket: up to ``tcap`` new tokens spread over up to ``bcap`` requests."""

    def __init__(self, runner: ModelRunner, tcap: int):
        self.r = runner
        self.tcap = tcap
        self.bcap = bc = max(1, min(tcap // 4, 128))
        dev = runner.device
        # static int region: input_ids | positions | out_cache_loc (tcap each) | cu_q (bc + 1) | q_lens | req_slots | seq_lens | last_idx |
        # sample_slots | temperature | top_p | min_p | top_k (bc each) | step (2)
        self.n_int = 3 * tcap + (bc + 2) + 9 * bc + 2          # cu_q padded to bc + 2 ints so that the int64 step word is 8-byte aligned
        self.static = torch.zeros(self.n_int + 3 * 8192, dtype=torch.int32, device=dev)
        s = self.static
        self.input_ids, self.positions, self.out_cache_loc = s[0:tcap], s[tcap:2 * tcap], s[2 * tcap:3 * tcap]
        o = 3 * tcap
        self.cu_q = s[o:o + bc + 1]
        o += bc + 2
        self.q_lens, self.req_slots, self.seq_lens = s[o:o + bc], s[o + bc:o + 2 * bc], s[o + 2 * bc:o + 3 * bc]
        self.last_idx, self.sample_slots = s[o + 3 * bc:o + 4 * bc], s[o + 4 * bc:o + 5 * bc]
        self.temperature = s[o + 5 * bc:o + 6 * bc].view(torch.float32)
        self.top_p = s[o + 6 * bc:o + 7 * bc].view(torch.float32)
        self.min_p = s[o + 7 * bc:o + 8 * bc].view(torch.float32)
        self.top_k = s[o + 8 * bc:o + 9 * bc]
        self.o_req = o
        self.step = s[o + 9 * bc:o + 9 * bc + 2].view(torch.int64)
        self.out_tokens = torch.zeros(bc, dtype=torch.int32, device=dev)
        self.graph = torch.cuda.CUDAGraph()
        self.num_kernels = 0

    def _run(self) -> None:
        r = self.r
        samp = SamplingTensors(self.temperature, self.top_k, self.top_p, self.min_p)
        fb = ForwardBatch(mode=ForwardMode.EXTEND, batch_size=self.bcap, num_tokens=self.tcap, input_ids=self.input_ids,
                          positions=self.positions, out_cache_loc=self.out_cache_loc, req_slots=self.req_slots, seq_lens=self.seq_lens,
                          cu_q=self.cu_q, q_lens=self.q_lens, last_idx=self.last_idx, sample_slots=self.sample_slots,
                          max_q_len=self.tcap, max_kv_len=r.max_context_len, page_table=r.page_table, page_size=r.page_size,
                          kv_pool=r.kv_pool, sampling=samp, next_token_buf=r.next_token_buf)
        r.set_rs(self.tcap)
        try:
            hidden = r.model(fb)
        finally:
            r.set_rs(0)
        h_last = ops.gather_rows(hidden, self.last_idx)
        logits = r.model.compute_logits(h_last)
        torch.ops.sglb.sample(logits, self.temperature, self.top_k, self.top_p, self.min_p, None, self.sample_slots, r.next_token_buf,
                              self.out_tokens, None, int(r.args.random_seed), self.step, 0)

    def _neutral(self) -> np.ndarray:
        """host image of an empty step: one dummy request owning every token (so warm-up / capture exercise the real kernels)"""
        bc, tcap = self.bcap, self.tcap
        h = np.zeros(self.n_int, dtype=np.int32)
        h[tcap:2 * tcap] = np.arange(tcap)                    # positions
        o3 = 3 * tcap
        h[o3 + 1:o3 + bc + 1] = tcap                          # cu_q: request 0 owns all tokens, the rest are empty
        o = self.o_req
        h[o] = tcap                                           # q_lens[0]
        h[o + 2 * bc] = tcap                                  # seq_lens[0]
        h[o + 5 * bc:o + 7 * bc] = np.float32(1.0).view(np.int32)   # temperature, top_p
        h[o + 8 * bc:o + 9 * bc] = 1                          # top_k = 1
        return h

    def capture(self, pool) -> None:
        self.static[:self.n_int].copy_(torch.from_numpy(self._neutral()))
        stream = torch.cuda.current_stream()
        for _ in range(2):
            self._run()
        stream.synchronize()
        nvl = self.r.tp.nvl
        if nvl is not None:
            self.r.tp.barrier()
            nvl.begin_capture()
        c0 = ops.launch_count()
        try:
            with torch.cuda.graph(self.graph, pool=pool, stream=stream):
                self._run()
        finally:
            if nvl is not None:
                nvl.end_capture()
        self.num_kernels = ops.launch_count() - c0

    def load(self, sb: ScheduleBatch, step: int) -> int:
        r_ = self.r
        B, bc, tcap = len(sb.reqs), self.bcap, self.tcap
        T = int(sb.q_lens.sum())
        r_._ring_i += 1
        host = r_._stage_ring[r_._ring_i % len(r_._stage_ring)]
        if host.numel() < self.static.numel():
            raise RuntimeError("prefill graph staging buffer too small")
        h = host.numpy()
        h[0:T] = sb.input_ids; h[T:tcap] = 0
        h[tcap:tcap + T] = sb.positions; h[tcap + T:2 * tcap] = 0
        h[2 * tcap:2 * tcap + T] = sb.out_cache_loc; h[2 * tcap + T:3 * tcap] = 0
        o3 = 3 * tcap
        h[o3] = 0
        cu = np.cumsum(sb.q_lens)
        h[o3 + 1:o3 + 1 + B] = cu; h[o3 + 1 + B:o3 + bc + 1] = T
        o = self.o_req
        h[o:o + B] = sb.q_lens; h[o + B:o + bc] = 0
        h[o + bc:o + bc + B] = sb.req_slots; h[o + bc + B:o + 2 * bc] = 0
        h[o + 2 * bc:o + 2 * bc + B] = sb.seq_lens; h[o + 2 * bc + B:o + 3 * bc] = 0
        h[o + 3 * bc:o + 3 * bc + B] = cu - 1; h[o + 3 * bc + B:o + 4 * bc] = 0
        # requests whose prompt is not finished by this chunk sample into the dummy slot 0
        h[o + 4 * bc:o + 4 * bc + B] = np.where(sb.sample_mask, sb.req_slots, 0); h[o + 4 * bc + B:o + 5 * bc] = 0
        s4 = sampling_rows(sb.reqs)
        h[o + 5 * bc:o + 5 * bc + B] = s4[0].view(np.int32); h[o + 5 * bc + B:o + 6 * bc] = np.float32(1.0).view(np.int32)
        h[o + 6 * bc:o + 6 * bc + B] = s4[1].view(np.int32); h[o + 6 * bc + B:o + 7 * bc] = np.float32(1.0).view(np.int32)
        h[o + 7 * bc:o + 7 * bc + B] = s4[2].view(np.int32); h[o + 7 * bc + B:o + 8 * bc] = 0
        h[o + 8 * bc:o + 8 * bc + B] = s4[3].astype(np.int32); h[o + 8 * bc + B:o + 9 * bc] = 1
        h[o + 9 * bc:o + 9 * bc + 2] = np.asarray([step], dtype=np.int64).view(np.int32)
        n = self.n_int
        nd = 0 if sb.pt_rows is None else sb.pt_rows.size
        if nd:
            if n + 3 * nd > self.static.numel():
                raise RuntimeError("page-table delta overflow in graph staging buffer")
            h[n:n + nd] = sb.pt_rows; h[n + nd:n + 2 * nd] = sb.pt_cols; h[n + 2 * nd:n + 3 * nd] = sb.pt_vals
        tot = n + 3 * nd
        self.static[:tot].copy_(host[:tot], non_blocking=True)
        r_.h2d_bytes += tot * 4
        return nd

    def delta_views(self, nd: int):
        n = self.n_int
        s = self.static
        return s[n:n + nd], s[n + nd:n + 2 * nd], s[n + 2 * nd:n + 3 * nd]


class _DecodeGraph:
    """One captured decode step for a fixed (padded) batch size."""

    def __init__(self, runner: ModelRunner, bs: int, greedy: bool = False):
        self.r = runner
        self.bs = bs
        self.greedy = greedy
        dev = runner.device
        # static int region: positions | out_cache_loc | req_slots | seq_lens | step(2) | temperature | top_p | min_p | top_k [| out_cache_loc_swa]
        self.has_swa = getattr(runner, "swa", None) is not None
        self.n_int = 8 * bs + 2 + (bs if self.has_swa else 0)
        self.static = torch.zeros(self.n_int + 3 * 8192, dtype=torch.int32, device=dev)
        s = self.static
        self.positions, self.out_cache_loc, self.req_slots, self.seq_lens = s[0:bs], s[bs:2 * bs], s[2 * bs:3 * bs], s[3 * bs:4 * bs]
        self.step = s[4 * bs:4 * bs + 2].view(torch.int64)
        o = 4 * bs + 2
        self.temperature = s[o:o + bs].view(torch.float32)
        self.top_p = s[o + bs:o + 2 * bs].view(torch.float32)
        self.min_p = s[o + 2 * bs:o + 3 * bs].view(torch.float32)
        self.top_k = s[o + 3 * bs:o + 4 * bs]
        self.out_cache_loc_swa = s[8 * bs + 2:9 * bs + 2] if self.has_swa else None
        self.out_tokens = torch.zeros(bs, dtype=torch.int32, device=dev)
        self.greedy_buf = torch.zeros(runner.tp.world_size, (bs + 1) // 2 * 2, 4, dtype=torch.bfloat16, device=dev) if greedy else None
        self.graph = torch.cuda.CUDAGraph()
        self.num_kernels = 0
        self.num_splits = runner._pick_splits(bs, min(runner.max_context_len, 4096))
        self.workspace = None
        if self.num_splits > 1:
            a0 = runner.model.model.layers[0].self_attn.attn
            self.workspace = (torch.empty(self.num_splits, bs, a0.num_heads, a0.head_dim, dtype=torch.float32, device=dev),
                              torch.empty(self.num_splits, bs, a0.num_heads, dtype=torch.float32, device=dev))

    def _fb(self) -> ForwardBatch:
        r = self.r
        samp = SamplingTensors(self.temperature, self.top_k, self.top_p, self.min_p)
        return ForwardBatch(mode=ForwardMode.DECODE, batch_size=self.bs, num_tokens=self.bs, input_ids=None,
                            positions=self.positions, out_cache_loc=self.out_cache_loc, req_slots=self.req_slots,
                            seq_lens=self.seq_lens, cu_q=None, q_lens=None, last_idx=None, sample_slots=self.req_slots,
                            max_q_len=1, max_kv_len=r.max_context_len, page_table=r.page_table, page_size=r.page_size,
                            kv_pool=r.kv_pool, num_splits=self.num_splits, attn_workspace=self.workspace, sampling=samp,
                            next_token_buf=r.next_token_buf, out_cache_loc_swa=self.out_cache_loc_swa,
                            swa_pt_offset=getattr(r, "swa_pt_offset", 0))

    def _run(self) -> None:
        r = self.r
        fb = self._fb()
        r.set_rs(self.bs)
        try:
            hidden = r.model(fb)
        finally:
            r.set_rs(0)
        if self.greedy:
            from ..layers.linear import vocab_parallel_greedy

            vocab_parallel_greedy(r.greedy_head(), hidden, self.greedy_buf, self.req_slots, r.next_token_buf, self.out_tokens)
            return
        logits = r.model.compute_logits(hidden)
        torch.ops.sglb.sample(logits, self.temperature, self.top_k, self.top_p, self.min_p, None, self.req_slots,
                              r.next_token_buf, self.out_tokens, None, int(r.args.random_seed), self.step, 0)

    def capture(self, pool) -> None:
        # neutral padding values
        bs = self.bs
        host = torch.zeros(self.static.numel(), dtype=torch.int32)
        hn = host.numpy()
        hn[3 * bs:4 * bs] = 1                       # seq_lens = 1
        o = 4 * bs + 2
        hn[o:o + 2 * bs] = np.float32(1.0).view(np.int32)   # temperature, top_p
        hn[o + 3 * bs:o + 4 * bs] = 1               # top_k = 1 (greedy)
        self.static.copy_(host)
        stream = torch.cuda.current_stream()
        for _ in range(2):   # warm-up outside capture (cuBLAS workspaces, lazy inits)
            self._run()
        stream.synchronize()
        nvl = self.r.tp.nvl
        if nvl is not None:
            self.r.tp.barrier()       # all ranks finished their eager warm-up collectives before anyone records
            nvl.begin_capture()       # collectives recorded into the graph use the graph arena region, parity from 0
        c0 = ops.launch_count()
        try:
            with torch.cuda.graph(self.graph, pool=pool, stream=stream):
                self._run()
        finally:
            if nvl is not None:
                nvl.end_capture()
        self.num_kernels = ops.launch_count() - c0

    def load(self, sb: ScheduleBatch, step: int, seed: int, s4: Optional[np.ndarray] = None) -> None:
        B, bs = len(sb.reqs), self.bs
        r_ = self.r
        r_._ring_i += 1
        host = r_._stage_ring[r_._ring_i % len(r_._stage_ring)]
        h = host.numpy()
        h[0:B] = sb.positions; h[B:bs] = 0
        h[bs:bs + B] = sb.out_cache_loc; h[bs + B:2 * bs] = 0
        h[2 * bs:2 * bs + B] = sb.req_slots; h[2 * bs + B:3 * bs] = 0
        h[3 * bs:3 * bs + B] = sb.seq_lens; h[3 * bs + B:4 * bs] = 1
        h[4 * bs:4 * bs + 2] = np.asarray([step], dtype=np.int64).view(np.int32)
        o = 4 * bs + 2
        if s4 is None:
            s4 = sampling_rows(sb.reqs)
        h[o:o + B] = s4[0].view(np.int32)
        h[o + B:o + bs] = np.float32(1.0).view(np.int32)
        h[o + bs:o + bs + B] = s4[1].view(np.int32)
        h[o + bs + B:o + 2 * bs] = np.float32(1.0).view(np.int32)
        h[o + 2 * bs:o + 2 * bs + B] = s4[2].view(np.int32)
        h[o + 2 * bs + B:o + 3 * bs] = 0
        h[o + 3 * bs:o + 3 * bs + B] = s4[3].astype(np.int32)
        h[o + 3 * bs + B:o + 4 * bs] = 1
        if self.has_swa:
            h[8 * bs + 2:8 * bs + 2 + B] = sb.out_cache_loc_swa; h[8 * bs + 2 + B:9 * bs + 2] = 0
        n = self.n_int
        nd = 0 if sb.pt_rows is None else sb.pt_rows.size
        if nd:
            if n + 3 * nd > h.size:
                raise RuntimeError("page-table delta overflow in graph staging buffer")
            h[n:n + nd] = sb.pt_rows; h[n + nd:n + 2 * nd] = sb.pt_cols; h[n + 2 * nd:n + 3 * nd] = sb.pt_vals
        tot = n + 3 * nd
        self.static[:tot].copy_(host[:tot], non_blocking=True)
        self.r.h2d_bytes += tot * 4

    def delta_views(self, nd: int):
        n = self.n_int
        s = self.static
        return s[n:n + nd], s[n + nd:n + 2 * nd], s[n + 2 * nd:n + 3 * nd]
