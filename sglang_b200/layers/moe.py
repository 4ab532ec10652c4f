"""Mixture-of-experts layer: top-k routing + grouped expert MLP, tensor- or expert-parallel.

Parity: reference ``python/sglang/srt/layers/moe/`` (``topk.py`` softmax top-k + renormalise,
``fused_moe_triton/layer.py`` FusedMoE, ``token_dispatcher/standard.py:98`` replicated-token dispatch + final
all-reduce, ``ep_moe/layer.py`` expert sharding, ``moe_align_kernel.cu`` align/permute).

Sharding: with ``ep_size == 1`` every rank holds all experts with the intermediate dim split over TP (moe-TP);
with ``ep_size == tp_size`` each rank holds ``E/ep`` whole experts (moe-EP).  In both cases tokens are replicated
across the TP group after attention (they already are, post all-reduce), each rank computes the contribution of
its weight shard and the partial outputs are summed by ONE all-reduce — the same collective the dense MLP needs,
so EP costs no extra dispatch/combine traffic inside a node (NVSwitch all-reduce of [T,H] vs an all-to-all of
[T*topk,H]: for top-2 the bytes are equal and the all-reduce has no routing imbalance).
An all-to-all token dispatcher (``A2ADispatcher``) is provided for the DP-attention configuration where tokens are
NOT replicated.
"""
from __future__ import annotations

from typing import Tuple

import torch
from torch import nn

from .. import ops
from ..ops import torch_ref
from ..parallel.parallel_state import get_ep_group, get_tp_group


def select_experts(router_logits: torch.Tensor, top_k: int, renormalize: bool = True) -> Tuple[torch.Tensor, torch.Tensor]:
    from ..ops import moe_ops

    return moe_ops.topk_softmax(router_logits, top_k, renormalize)


class FusedMoE(nn.Module):
    def __init__(self, num_experts: int, top_k: int, hidden_size: int, intermediate_size: int, dtype=torch.bfloat16,
                 renormalize: bool = True, ep_size: int = 1, quant=None, layer_id: int = 0):
        super().__init__()
        self.tp = get_tp_group()
        self.num_experts, self.top_k, self.renormalize, self.layer_id = num_experts, top_k, renormalize, layer_id
        tp = self.tp.world_size
        self.ep_size = ep_size if ep_size > 1 else 1
        if self.ep_size > 1:
            assert tp % self.ep_size == 0 and num_experts % self.ep_size == 0
            self.moe_tp = tp // self.ep_size                   # TP inside each expert shard
            self.ep_rank = self.tp.rank // self.moe_tp
            self.moe_tp_rank = self.tp.rank % self.moe_tp
        else:
            self.moe_tp, self.ep_rank, self.moe_tp_rank = tp, 0, self.tp.rank
        self.local_experts = num_experts // self.ep_size
        self.expert_offset = self.ep_rank * self.local_experts
        assert intermediate_size % self.moe_tp == 0
        self.inter_local = intermediate_size // self.moe_tp
        self.w13 = nn.Parameter(torch.zeros(self.local_experts, 2 * self.inter_local, hidden_size, dtype=dtype), requires_grad=False)
        self.w2 = nn.Parameter(torch.zeros(self.local_experts, hidden_size, self.inter_local, dtype=dtype), requires_grad=False)
        self.expert_load = None  # EPLB recorder hook
        self.placement = None    # eplb/placement.py ExpertPlacement once experts are replicated / moved; None = contiguous default
        self.rep_table = self.rep_count = None

    def weight_loader(self, param: torch.Tensor, loaded: torch.Tensor, expert_id: int, shard: str) -> None:
        """shard in {'w1' (gate), 'w3' (up), 'w2' (down)}; ``loaded`` is the full (unsharded) expert matrix"""
        le = expert_id - self.expert_offset
        if le < 0 or le >= self.local_experts:
            return
        I, r = self.inter_local, self.moe_tp_rank
        if shard == "w1":
            param.data[le, :I].copy_(loaded.narrow(0, r * I, I))
        elif shard == "w3":
            param.data[le, I:].copy_(loaded.narrow(0, r * I, I))
        elif shard == "w2":
            param.data[le].copy_(loaded.narrow(1, r * I, I))
        else:
            raise ValueError(shard)

    def set_placement(self, pl, w13: torch.Tensor, w2: torch.Tensor) -> None:
        """Adopt a physical placement (eplb/placement.py): ``w13``/``w2`` are this rank's new slot contents."""
        assert self.moe_tp == 1 or pl.ep_size == self.ep_size
        self.placement = pl
        self.w13 = nn.Parameter(w13, requires_grad=False)
        self.w2 = nn.Parameter(w2, requires_grad=False)
        self.local_experts = pl.slots
        self.expert_offset = self.ep_rank * pl.slots
        self.rep_table, self.rep_count = pl.tables(self.num_experts, w13.device)

    def to_physical(self, ids: torch.Tensor) -> torch.Tensor:
        """logical expert ids [T, K] -> physical slot ids; replicas of a hot expert take its tokens round-robin by token index"""
        lid = ids.long()
        T, K = lid.shape
        pick = (torch.arange(T, device=ids.device)[:, None] + torch.arange(K, device=ids.device)[None]) % self.rep_count[lid]
        return self.rep_table[lid, pick].to(ids.dtype)

    def forward(self, x: torch.Tensor, router_logits: torch.Tensor) -> torch.Tensor:
        w, ids = select_experts(router_logits, self.top_k, self.renormalize)
        return self.forward_routed(x, w, ids)

    def forward_a2a(self, x: torch.Tensor, w: torch.Tensor, ids: torch.Tensor) -> torch.Tensor:
        """All-to-all expert dispatch for DP-attention (tokens are NOT replicated across ranks): every (token, k) pair travels to the
        rank that owns its expert, the owners run their experts on what they received, results travel back and are combined with the
        routing weights on the home rank.  Reference: ``token_dispatcher/{deepep,flashinfer}.py`` dispatch/combine; here over
        ``all_to_all_single`` (NCCL on NVLink / Gloo on the CPU tier).  Requires whole experts per rank (ep_size == tp_size)."""
        from ..ops import moe_ops

        tp = self.tp
        W = tp.world_size
        assert self.moe_tp == 1 and self.ep_size == W, "all-to-all dispatch needs ep_size == tp_size"
        self._a2a_calls = getattr(self, "_a2a_calls", 0) + 1
        T, H = x.shape
        K = ids.shape[1]
        flat_e = ids.reshape(-1).long()
        dest = flat_e // self.local_experts
        order = torch.argsort(dest, stable=True)
        send_counts = torch.bincount(dest, minlength=W)
        recv_counts = torch.empty_like(send_counts)
        tp.all_to_all_single(recv_counts, send_counts)
        sc, rc = send_counts.tolist(), recv_counts.tolist()
        tok = order // K
        send_x = x.index_select(0, tok).contiguous()
        send_e = flat_e.index_select(0, order).to(torch.int32).contiguous()
        R = int(sum(rc))
        recv_x = torch.empty(R, H, dtype=x.dtype, device=x.device)
        recv_e = torch.empty(R, dtype=torch.int32, device=x.device)
        tp.all_to_all_single(recv_x, send_x, rc, sc)
        tp.all_to_all_single(recv_e, send_e, rc, sc)
        if R:
            ones = torch.ones(R, 1, dtype=torch.float32, device=x.device)
            y = moe_ops.fused_moe(recv_x, self.w13, self.w2, ones, recv_e.view(R, 1), self.expert_offset)
        else:
            y = recv_x
        back = torch.empty(send_x.shape[0], H, dtype=x.dtype, device=x.device)
        tp.all_to_all_single(back, y.contiguous(), sc, rc)
        out = torch.zeros(T, H, dtype=torch.float32, device=x.device)
        out.index_add_(0, tok, back.float() * w.reshape(-1).index_select(0, order).float()[:, None])
        return out.to(x.dtype)

    def forward_routed(self, x: torch.Tensor, w: torch.Tensor, ids: torch.Tensor) -> torch.Tensor:
        """expert MLPs for an externally computed routing (``w`` fp32 [T,K], ``ids`` int32 [T,K] global expert ids) — grouped /
        biased / sigmoid routers (DeepSeek) live in the model file"""
        from ..ops import moe_ops
        from .dp_attention import get_dp_ctx

        ctx = get_dp_ctx()
        if self.expert_load is not None:
            self.expert_load.record(self.layer_id, ids)          # logical ids: what the rebalancer plans on
        if self.rep_table is not None:
            ids = self.to_physical(ids)
        if ctx is not None and getattr(ctx, "a2a_active", False):
            from .dwdp import get_dwdp_manager

            mgr = get_dwdp_manager()
            if mgr is not None and mgr.use(ctx, x):          # big prefill: move the weights, not the tokens (layers/dwdp.py)
                return mgr.forward(self, x, w, ids)
            return self.forward_a2a(x, w, ids)
        tp = self.tp
        if (tp.world_size > 1 and tp.nvl is not None and x.is_cuda and x.dtype == torch.bfloat16 and x.dim() == 2
                and moe_ops._has("moe_grouped_mlp_push") and tp.world_size * x.numel() * 2 <= tp.nvl.slot_bytes):
            from ..environ import envs
            from ..parallel.parallel_state import get_pp_group
            from .linear import PendingAllReduce

            if envs.SGLB_FUSED_AR.get() == "push" and get_pp_group().world_size == 1:
                # combine kernel pushes the partial sums to every peer; the next RMSNorm reduces them locally
                off = tp.nvl.moe_push(x, self.w13, self.w2, w, ids.to(torch.int32), self.expert_offset)
                return PendingAllReduce(tp.nvl, off, x.shape[0], x.shape[1], "push")
        y = moe_ops.fused_moe(x, self.w13, self.w2, w, ids, self.expert_offset)
        if tp.world_size > 1:
            y = tp.all_reduce(y)
        return y
