"""Process groups: one process per GPU, ``torch.distributed`` (NCCL on CUDA, Gloo on CPU) for the plumbing.

Parity: reference ``python/sglang/srt/distributed/parallel_state.py`` (``GroupCoordinator:237``,
``init_distributed_environment:2236``, ``initialize_model_parallel:2328``: world / TP / PP / EP groups, a CPU
(Gloo) twin of every group for control-plane objects, ``all_reduce:648`` dispatch).

B200-first: a GroupCoordinator may carry an ``NvlComm`` (``parallel/nvl_comm.py``: symmetric-memory arena with
peer + multicast pointers over NVSwitch).  ``all_reduce`` uses our one-shot / two-shot kernels when the message
fits the arena, and the fused GEMM+all-reduce path bypasses this API entirely (see ``layers/linear.py``).
NCCL stays for bootstrap, large prefill messages, PP send/recv and as the measured baseline.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Any, List, Optional

import torch
import torch.distributed as dist


class GroupCoordinator:
    def __init__(self, ranks: List[int], rank: int, device: torch.device, name: str):
        self.ranks = ranks
        self.world_size = len(ranks)
        self.rank_in_group = ranks.index(rank) if rank in ranks else -1
        self.global_rank = rank
        self.device = device
        self.name = name
        self.device_group = None
        self.cpu_group = None
        self.nvl = None  # NvlComm
        self.comm_bytes = 0
        self.comm_calls = 0

    # ---- properties
    @property
    def rank(self) -> int:
        return self.rank_in_group

    @property
    def is_first(self) -> bool:
        return self.rank_in_group == 0

    @property
    def is_last(self) -> bool:
        return self.rank_in_group == self.world_size - 1

    # ---- collectives
    def all_reduce(self, x: torch.Tensor) -> torch.Tensor:
        if self.world_size == 1:
            return x
        self.comm_calls += 1
        self.comm_bytes += x.numel() * x.element_size()
        if self.nvl is not None and self.nvl.can_all_reduce(x):
            return self.nvl.all_reduce(x)
        dist.all_reduce(x, group=self.device_group)
        return x

    def all_gather(self, x: torch.Tensor, dim: int = -1) -> torch.Tensor:
        if self.world_size == 1:
            return x
        if dim < 0:
            dim += x.dim()
        x = x.contiguous()
        if x.dim() == 0:
            x = x.view(1)
        self.comm_calls += 1
        # flat [world * n0, ...] output: the layout both NCCL and Gloo accept for all_gather_into_tensor
        flat = torch.empty((self.world_size * x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
        dist.all_gather_into_tensor(flat, x, group=self.device_group)
        out = flat.view((self.world_size,) + tuple(x.shape)).movedim(0, dim)
        shp = list(x.shape)
        shp[dim] *= self.world_size
        return out.reshape(shp)

    def reduce_scatter(self, x: torch.Tensor) -> torch.Tensor:
        """x [world*n, ...] -> [n, ...] (sum over ranks of each rank's slice)"""
        if self.world_size == 1:
            return x
        n = x.shape[0] // self.world_size
        if not x.is_cuda:       # Gloo has no reduce-scatter: all-reduce and keep this rank's slice (CPU tier only)
            x = x.contiguous().clone()
            dist.all_reduce(x, group=self.device_group)
            return x[self.rank_in_group * n:(self.rank_in_group + 1) * n].contiguous()
        out = torch.empty((n,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
        dist.reduce_scatter_tensor(out, x.contiguous(), group=self.device_group)
        return out

    def all_to_all_single(self, out: torch.Tensor, inp: torch.Tensor, out_splits=None, in_splits=None) -> torch.Tensor:
        if self.world_size == 1:
            out.copy_(inp)
            return out
        dist.all_to_all_single(out, inp, out_splits, in_splits, group=self.device_group)
        return out

    def broadcast(self, x: torch.Tensor, src: int = 0) -> torch.Tensor:
        if self.world_size == 1:
            return x
        dist.broadcast(x, src=self.ranks[src], group=self.device_group)
        return x

    def send(self, x: torch.Tensor, dst: int) -> None:
        dist.send(x, dst=self.ranks[dst], group=self.device_group)

    def recv(self, x: torch.Tensor, src: int) -> torch.Tensor:
        dist.recv(x, src=self.ranks[src], group=self.device_group)
        return x

    def barrier(self) -> None:
        if self.world_size > 1:
            dist.barrier(group=self.cpu_group)

    def broadcast_pyobj(self, obj: Any, src: int = 0) -> Any:
        """control-plane object broadcast over the Gloo twin (reference ``broadcast_pyobj``)"""
        if self.world_size == 1:
            return obj
        lst = [obj if self.rank_in_group == src else None]
        dist.broadcast_object_list(lst, src=self.ranks[src], group=self.cpu_group)
        return lst[0]

    def all_gather_pyobj(self, obj: Any) -> List[Any]:
        if self.world_size == 1:
            return [obj]
        out = [None] * self.world_size
        dist.all_gather_object(out, obj, group=self.cpu_group)
        return out

    def send_pyobj(self, obj: Any, dst: int) -> None:
        dist.send_object_list([obj], dst=self.ranks[dst], group=self.cpu_group)

    def recv_pyobj(self, src: int) -> Any:
        lst = [None]
        dist.recv_object_list(lst, src=self.ranks[src], group=self.cpu_group)
        return lst[0]


@dataclass
class ParallelState:
    world: GroupCoordinator
    tp: GroupCoordinator
    pp: GroupCoordinator
    ep: GroupCoordinator
    rank: int
    world_size: int
    local_rank: int
    # DP-attention: attention is tensor-parallel only inside ``attn_tp`` (tp_size / dp_size consecutive ranks); each of the
    # ``attn_dp_size`` groups serves its own requests, the MLP / MoE / vocab layers stay parallel over the full ``tp`` group
    attn_tp: Any = None
    attn_dp_rank: int = 0
    attn_dp_size: int = 1
    # decode context parallelism: ``dcp`` = consecutive ranks of the attention-TP group that hold the SAME (replicated) KV head and
    # split the tokens of every sequence between them (layers/dcp.py)
    dcp: Any = None
    # prefill context parallelism: the TP group is split into ``cp_size`` attention groups exactly like DP-attention, but the groups share
    # every request: in a prefill each takes a zig-zag slice of the query rows and all-gathers K/V over ``cp`` (layers/cp.py)
    cp: Any = None
    cp_mode: bool = False


_STATE: Optional[ParallelState] = None


def _single(device: torch.device) -> ParallelState:
    g = lambda n: GroupCoordinator([0], 0, device, n)
    tp = g("tp")
    return ParallelState(g("world"), tp, g("pp"), g("ep"), 0, 1, 0, tp, 0, 1)


def compute_group_ranks(world_size: int, tp_size: int, pp_size: int, ep_size: int = 1):
    """Rank layout (reference ``initialize_model_parallel`` docstring :2363-2390): TP is the fastest-varying
    dimension, PP the slowest; EP groups partition each TP group into contiguous blocks of ``ep_size``."""
    assert world_size == tp_size * pp_size, (world_size, tp_size, pp_size)
    tp_groups = [list(range(p * tp_size, (p + 1) * tp_size)) for p in range(pp_size)]
    pp_groups = [list(range(t, world_size, tp_size)) for t in range(tp_size)]
    ep_groups = []
    for g in tp_groups:
        for i in range(0, tp_size, max(1, ep_size)):
            ep_groups.append(g[i:i + max(1, ep_size)])
    return tp_groups, pp_groups, ep_groups


def init_distributed(tp_size: int = 1, pp_size: int = 1, ep_size: int = 1, device: Optional[str] = None,
                     init_method: Optional[str] = None, rank: Optional[int] = None, world_size: Optional[int] = None,
                     enable_nvl: bool = True, timeout_s: int = 600, attn_dp_size: int = 1, dcp_size: int = 1, cp_size: int = 1) -> ParallelState:
    global _STATE
    if cp_size > 1:
        assert attn_dp_size == 1, "prefill CP and DP-attention both partition the TP group: enable one of them"
        attn_dp_size = cp_size
    use_cuda = (device or ("cuda" if torch.cuda.is_available() else "cpu")).startswith("cuda")
    ws = tp_size * pp_size
    if (_STATE is not None and ws > 1 and _STATE.world_size == ws and _STATE.tp.world_size == tp_size
            and _STATE.pp.world_size == pp_size and _STATE.ep.world_size == max(1, ep_size) and _STATE.attn_dp_size == max(1, attn_dp_size)
            and (_STATE.dcp.world_size if _STATE.dcp is not None else 1) == max(1, dcp_size) and _STATE.cp_mode == (cp_size > 1)):
        return _STATE       # idempotent: a second Engine in the same SPMD job reuses the groups and the NVL arena
    if ws == 1:
        dev = torch.device("cuda", torch.cuda.current_device()) if use_cuda else torch.device("cpu")
        _STATE = _single(dev)
        return _STATE
    rank = int(os.environ.get("RANK", 0)) if rank is None else rank
    world_size = int(os.environ.get("WORLD_SIZE", ws)) if world_size is None else world_size
    local_rank = int(os.environ.get("LOCAL_RANK", rank))
    assert world_size == ws, f"world size {world_size} != tp*pp {ws}"
    if use_cuda:
        torch.cuda.set_device(local_rank)
        dev = torch.device("cuda", local_rank)
    else:
        dev = torch.device("cpu")
    if not dist.is_initialized():
        import datetime

        if init_method is None:
            addr = os.environ.get("MASTER_ADDR", "127.0.0.1")
            port = os.environ.get("MASTER_PORT", "29500")
            init_method = f"tcp://{addr}:{port}"
        kw = {}
        if use_cuda:
            kw["device_id"] = dev
        dist.init_process_group(backend="nccl" if use_cuda else "gloo", init_method=init_method, rank=rank,
                                world_size=world_size, timeout=datetime.timedelta(seconds=timeout_s), **kw)

    def make(groups, name) -> GroupCoordinator:
        mine = None
        for ranks in groups:
            dg = dist.new_group(ranks, backend="nccl" if use_cuda else "gloo") if len(ranks) > 1 else None
            cg = dist.new_group(ranks, backend="gloo") if len(ranks) > 1 else None
            if rank in ranks:
                mine = GroupCoordinator(ranks, rank, dev, name)
                mine.device_group, mine.cpu_group = dg, cg
        return mine

    tp_groups, pp_groups, ep_groups = compute_group_ranks(world_size, tp_size, pp_size, ep_size)
    world = make([list(range(world_size))], "world")
    tp = make(tp_groups, "tp")
    pp = make(pp_groups, "pp")
    ep = make(ep_groups, "ep")
    attn_tp, dp_rank = tp, 0
    if attn_dp_size > 1:
        assert tp_size % attn_dp_size == 0, "dp attention: dp_size must divide tp_size"
        a = tp_size // attn_dp_size
        groups = [g[i:i + a] for g in tp_groups for i in range(0, tp_size, a)]
        attn_tp = make(groups, "attn_tp")
        dp_rank = tp.rank // a
    _STATE = ParallelState(world, tp, pp, ep, rank, world_size, local_rank, attn_tp, dp_rank, max(1, attn_dp_size))
    if cp_size > 1:
        a = tp_size // cp_size
        _STATE.cp = make([[g[i + j * a] for j in range(cp_size)] for g in tp_groups for i in range(a)], "cp")
        _STATE.cp_mode = True
    if dcp_size > 1:
        a = tp_size // max(1, attn_dp_size)
        assert a % dcp_size == 0, "dcp_size must divide the attention-TP size"
        attn_groups = [g[i:i + a] for g in tp_groups for i in range(0, tp_size, a)]
        _STATE.dcp = make([g[i:i + dcp_size] for g in attn_groups for i in range(0, a, dcp_size)], "dcp")
    from ..environ import envs as _envs

    if use_cuda and enable_nvl and tp.world_size > 1 and _envs.SGLB_COMM_IMPL.get() != "nccl":      # "nccl": library collectives only (A/B runs)
        try:
            from .nvl_comm import NvlComm

            tp.nvl = NvlComm.create(tp)
        except Exception as e:  # keep NCCL as the fallback; say so loudly
            import logging

            logging.getLogger(__name__).warning("NVL symmetric-memory comm unavailable, using NCCL only: %s", e)
    return _STATE


def get_state() -> ParallelState:
    global _STATE
    if _STATE is None:
        _STATE = _single(torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu"))
    return _STATE


def set_state(s: Optional[ParallelState]) -> None:
    global _STATE
    _STATE = s


def get_tp_group() -> GroupCoordinator:
    return get_state().tp


def get_pp_group() -> GroupCoordinator:
    return get_state().pp


def get_ep_group() -> GroupCoordinator:
    return get_state().ep


def get_attn_tp_group() -> GroupCoordinator:
    st = get_state()
    return st.attn_tp if st.attn_tp is not None else st.tp


def get_cp_group() -> Optional[GroupCoordinator]:
    st = get_state()
    return st.cp if st.cp_mode and st.cp is not None and st.cp.world_size > 1 else None


def get_dcp_group() -> Optional[GroupCoordinator]:
    d = get_state().dcp
    return d if d is not None and d.world_size > 1 else None


def get_tp_rank() -> int:
    return get_state().tp.rank


def get_tp_size() -> int:
    return get_state().tp.world_size


def destroy_distributed() -> None:
    global _STATE
    _STATE = None
    if dist.is_initialized():
        dist.destroy_process_group()
