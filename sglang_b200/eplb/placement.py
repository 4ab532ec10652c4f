"""Physical expert placement with redundancy, live rebalancing, and rank-fault recovery (elastic EP).

Reference behaviour: ``python/sglang/srt/eplb/`` (``eplb_algorithms/deepseek.py`` replicate-then-pack, ``expert_location.py``
logical->physical dispatch tables, ``expert_location_updater.py`` live weight moves) and ``python/sglang/srt/elastic_ep/``
(``elastic_ep.py`` rank-fault masks + re-planning on the surviving ranks, ``expert_backup_manager.py`` host-memory expert backups).

Model: every EP rank owns ``slots`` PHYSICAL expert slots.  A placement maps each physical slot to a logical expert (or -1);
hot experts may own several slots (replicas), dead ranks own none.  The MoE layer translates the router's logical ids to physical
ids (replicas are picked round-robin by token index) and everything downstream — the grouped GEMM, all-to-all dispatch, DWDP —
works on physical ids unchanged, because physical id // slots is still "the rank that owns it".
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist


@dataclass
class ExpertPlacement:
    phys_to_logical: np.ndarray                       # [ep_size, slots] int64, -1 = empty slot

    @property
    def ep_size(self) -> int:
        return self.phys_to_logical.shape[0]

    @property
    def slots(self) -> int:
        return self.phys_to_logical.shape[1]

    @staticmethod
    def default(num_experts: int, ep_size: int) -> "ExpertPlacement":
        return ExpertPlacement(np.arange(num_experts, dtype=np.int64).reshape(ep_size, num_experts // ep_size))

    def replicas(self, num_experts: int) -> List[List[int]]:
        out: List[List[int]] = [[] for _ in range(num_experts)]
        flat = self.phys_to_logical.reshape(-1)
        for p, e in enumerate(flat):
            if e >= 0:
                out[int(e)].append(p)
        return out

    def tables(self, num_experts: int, device) -> Tuple[torch.Tensor, torch.Tensor]:
        reps = self.replicas(num_experts)
        assert all(reps), f"experts without a live replica: {[e for e, r in enumerate(reps) if not r]}"
        R = max(len(r) for r in reps)
        tab = np.asarray([r + [r[0]] * (R - len(r)) for r in reps], dtype=np.int64)
        return torch.from_numpy(tab).to(device), torch.tensor([len(r) for r in reps], dtype=torch.long, device=device)


def plan(load: np.ndarray, ep_size: int, slots: int, live: Optional[Sequence[int]] = None) -> ExpertPlacement:
    """Replicate-then-pack.  ``load[e]`` is the observed token count of logical expert e.  All ``len(live) * slots`` physical
    slots are used: extra slots go, one at a time, to the expert with the highest load PER REPLICA; replicas are then packed onto
    the live ranks heaviest-first onto the least-loaded rank with a free slot, avoiding two replicas of one expert on a rank."""
    E = int(load.shape[0])
    live = list(range(ep_size)) if live is None else sorted(int(r) for r in live)
    cap = len(live) * slots
    if cap < E:
        raise ValueError(f"{len(live)} live ranks x {slots} slots cannot hold {E} experts")
    nrep = np.ones(E, dtype=np.int64)
    w = load.astype(np.float64) + 1e-6
    for _ in range(cap - E):
        score = np.where(nrep < len(live), w / nrep, -1.0)          # a second replica on the same rank buys nothing
        if score.max() < 0:
            break
        nrep[int(np.argmax(score))] += 1
    items = sorted(((w[e] / nrep[e], e) for e in range(E) for _ in range(nrep[e])), key=lambda t: (-t[0], t[1]))
    fill: Dict[int, List[int]] = {r: [] for r in live}
    rload = {r: 0.0 for r in live}
    for wt, e in items:
        cands = [r for r in live if len(fill[r]) < slots]
        fresh = [r for r in cands if e not in fill[r]] or cands
        r = min(fresh, key=lambda r: (rload[r], r))
        fill[r].append(e)
        rload[r] += wt
    p2l = np.full((ep_size, slots), -1, dtype=np.int64)
    for r in live:
        p2l[r, :len(fill[r])] = fill[r]
    return ExpertPlacement(p2l)


def imbalance(load: np.ndarray, pl: ExpertPlacement) -> float:
    """max rank load / mean live-rank load when every expert's tokens split evenly over its replicas"""
    reps = pl.replicas(load.shape[0])
    rl = np.zeros(pl.ep_size)
    for e, ps in enumerate(reps):
        for p in ps:
            rl[p // pl.slots] += load[e] / len(ps)
    livem = (pl.phys_to_logical >= 0).any(axis=1)
    return float(rl.max() / max(rl[livem].mean(), 1e-9))


class ExpertBackupStore:
    """Host-memory copy of the ring neighbour's experts (rank r keeps rank r+1's), so the loss of any single rank — or any set of
    non-adjacent ranks — is recoverable without touching the checkpoint."""

    def __init__(self) -> None:
        self.data: Dict[Tuple[int, int], Tuple[torch.Tensor, torch.Tensor]] = {}     # (layer index, logical expert) -> (w13, w2) on host
        self.bytes = 0

    @staticmethod
    def build(layers: List[Any], group: Any) -> "ExpertBackupStore":
        st = ExpertBackupStore()
        W, me = group.world_size, group.rank
        for li, m in enumerate(layers):
            g13, g2 = _gather_layer(m, group)
            nb = (me + 1) % W
            s = m.local_experts
            for slot in range(s):
                e = int(m.placement.phys_to_logical[nb, slot])
                if e < 0 or (li, e) in st.data:
                    continue
                a, b = g13[nb * s + slot].cpu().clone(), g2[nb * s + slot].cpu().clone()
                st.data[(li, e)] = (a, b)
                st.bytes += a.numel() * a.element_size() + b.numel() * b.element_size()
        return st


def _gather_layer(m: Any, group: Any) -> Tuple[torch.Tensor, torch.Tensor]:
    W = group.world_size
    g13 = torch.empty((W * m.w13.shape[0],) + tuple(m.w13.shape[1:]), dtype=m.w13.dtype, device=m.w13.device)
    g2 = torch.empty((W * m.w2.shape[0],) + tuple(m.w2.shape[1:]), dtype=m.w2.dtype, device=m.w2.device)
    if W == 1:
        return m.w13.data, m.w2.data
    dist.all_gather_into_tensor(g13, m.w13.data.contiguous(), group=group.device_group)
    dist.all_gather_into_tensor(g2, m.w2.data.contiguous(), group=group.device_group)
    return g13, g2


def apply_placement(layers: List[Any], new: List[ExpertPlacement], group: Any, dead: Sequence[int] = (),
                    backup: Optional[ExpertBackupStore] = None) -> Dict[str, int]:
    """Move expert weights so every layer matches its new placement.  Sources, in order: this rank's own slots, a LIVE peer's slot
    (one all-gather per layer), the backup holder (broadcast).  ``dead`` ranks' memory is never read."""
    W, me = group.world_size, group.rank
    dead = set(int(d) for d in dead)
    stats = {"local": 0, "peer": 0, "backup": 0}
    for li, (m, pl) in enumerate(zip(layers, new)):
        old = m.placement
        g13, g2 = _gather_layer(m, group)
        s_old, s_new = old.slots, pl.slots
        src: Dict[int, int] = {}                                # logical expert -> old physical id on a live rank (own rank preferred)
        for p, e in enumerate(old.phys_to_logical.reshape(-1)):
            r = p // s_old
            if e >= 0 and r not in dead and (int(e) not in src or r == me):
                src[int(e)] = p
        need = sorted({int(e) for e in pl.phys_to_logical.reshape(-1) if e >= 0} - set(src))
        restored: Dict[int, Tuple[torch.Tensor, torch.Tensor]] = {}
        for e in need:                                          # only a backup holds it: the lowest live holder broadcasts
            have = torch.tensor([1 if (backup is not None and (li, e) in backup.data and me not in dead) else 0])
            allh = [torch.zeros_like(have) for _ in range(W)]
            if W > 1:
                dist.all_gather(allh, have, group=group.cpu_group)
            else:
                allh = [have]
            holders = [r for r in range(W) if int(allh[r]) and r not in dead]
            if not holders:
                raise RuntimeError(f"layer {li} expert {e}: no live replica and no backup — reload it from the checkpoint")
            h = holders[0]
            a = backup.data[(li, e)][0].to(m.w13.device) if me == h else torch.empty(m.w13.shape[1:], dtype=m.w13.dtype, device=m.w13.device)
            b = backup.data[(li, e)][1].to(m.w2.device) if me == h else torch.empty(m.w2.shape[1:], dtype=m.w2.dtype, device=m.w2.device)
            if W > 1:
                dist.broadcast(a, group.ranks[h], group=group.device_group)
                dist.broadcast(b, group.ranks[h], group=group.device_group)
            restored[e] = (a, b)
        n13 = torch.zeros((s_new,) + tuple(m.w13.shape[1:]), dtype=m.w13.dtype, device=m.w13.device)
        n2 = torch.zeros((s_new,) + tuple(m.w2.shape[1:]), dtype=m.w2.dtype, device=m.w2.device)
        for slot, e in enumerate(pl.phys_to_logical[me]):
            e = int(e)
            if e < 0:
                continue
            if e in restored:
                n13[slot], n2[slot] = restored[e]
                stats["backup"] += 1
            else:
                p = src[e]
                n13[slot], n2[slot] = g13[p], g2[p]
                stats["local" if p // s_old == me else "peer"] += 1
        m.set_placement(pl, n13, n2)
    return stats


# ---- runner-level entry points (reached through Engine.collective_rpc / the /eplb_rebalance and /elastic_ep routes) ---------------
def _moe_layers(runner: Any) -> List[Any]:
    layers = [m for m in runner.model.modules() if m.__class__.__name__ == "FusedMoE"]
    if not layers:
        raise ValueError("model has no MoE layers")
    if any(m.moe_tp != 1 for m in layers):
        raise ValueError("expert placement moves whole experts: needs ep_size == tp_size")
    for m in layers:
        if m.placement is None:
            m.placement = ExpertPlacement.default(m.num_experts, m.ep_size)
    return layers


def _after_move(runner: Any) -> None:
    if getattr(runner, "dwdp", None) is not None:               # DWDP's ping-pong buffers are sized by the slot count
        from ..layers import dwdp

        dwdp.attach(runner, runner.dwdp.min_tokens)
    if hasattr(runner, "invalidate_graphs"):
        runner.invalidate_graphs()


def build_backup(runner: Any) -> Dict[str, int]:
    layers = _moe_layers(runner)
    runner.expert_backup = ExpertBackupStore.build(layers, layers[0].tp)
    return {"experts": len(runner.expert_backup.data), "bytes": runner.expert_backup.bytes}


def rebalance(runner: Any, redundant_slots_per_rank: int = 0, load: Optional[List[List[float]]] = None) -> Dict[str, Any]:
    """Re-plan every MoE layer from the recorded (or given) expert load and move the weights.  Returns before/after imbalance."""
    layers = _moe_layers(runner)
    rec = getattr(runner, "expert_recorder", None)
    counts = np.asarray(load, dtype=np.float64) if load is not None else (rec.counts.astype(np.float64) if rec is not None else None)
    if counts is None:
        raise ValueError("no expert load recorded: start the expert-distribution recorder first, or pass load=")
    grp = layers[0].tp
    if grp.world_size > 1:                                        # every rank must plan from the same numbers
        t = torch.from_numpy(counts.copy())
        dist.all_reduce(t, group=grp.cpu_group)
        counts = t.numpy()
    dead = sorted(getattr(runner, "dead_ep_ranks", ()))
    live = [r for r in range(layers[0].ep_size) if r not in dead]
    before, after, new = [], [], []
    for li, m in enumerate(layers):
        slots = max(-(-m.num_experts // len(live)), m.num_experts // m.ep_size) + int(redundant_slots_per_rank)
        pl = plan(counts[li], m.ep_size, slots, live)
        before.append(imbalance(counts[li], m.placement))
        after.append(imbalance(counts[li], pl))
        new.append(pl)
    stats = apply_placement(layers, new, grp, dead, getattr(runner, "expert_backup", None))
    _after_move(runner)
    return {"imbalance_before": before, "imbalance_after": after, "moves": stats, "slots": new[0].slots}


def recover(runner: Any, dead_ranks: Sequence[int]) -> Dict[str, Any]:
    """Rank-fault recovery: re-plan on the surviving ranks (uniform load unless a recording exists) and restore the experts whose
    only copies died from the host backups."""
    layers = _moe_layers(runner)
    runner.dead_ep_ranks = sorted(set(getattr(runner, "dead_ep_ranks", ())) | {int(d) for d in dead_ranks})
    live = [r for r in range(layers[0].ep_size) if r not in runner.dead_ep_ranks]
    if not live:
        raise ValueError("no surviving EP rank")
    rec = getattr(runner, "expert_recorder", None)
    new = []
    for li, m in enumerate(layers):
        load = rec.counts[li].astype(np.float64) if rec is not None and rec.counts[li].sum() > 0 else np.ones(m.num_experts)
        new.append(plan(load, m.ep_size, -(-m.num_experts // len(live)), live))
    stats = apply_placement(layers, new, layers[0].tp, runner.dead_ep_ranks, getattr(runner, "expert_backup", None))
    _after_move(runner)
    return {"live_ranks": live, "moves": stats, "slots": new[0].slots}
