"""Control-plane handlers executed by the model runner between steps: live weight updates (RL), weight
introspection, memory release/resume, profiling, LoRA adapters, expert-distribution recording.

Parity: reference ``model_executor/model_runner_components/weight_updater.py``, ``entrypoints/engine.py:1396-1648``
(update_weights_from_{disk,tensor,distributed}, get_weights_by_name, release/resume_memory_occupation),
``managers/scheduler_components/profiler_manager.py:53`` (torch profiler start/stop, per-rank traces).
"""
from __future__ import annotations

import logging
import os
import time
from typing import Any, List, Tuple

import torch

from ..managers import io_struct as ios

logger = logging.getLogger(__name__)


def handle(runner: Any, msg: Any) -> Any:
    if isinstance(msg, ios.UpdateWeightsFromDiskReq):
        return _update_from_disk(runner, msg)
    if isinstance(msg, ios.UpdateWeightsFromTensorReq):
        return _update_from_tensor(runner, msg)
    if isinstance(msg, ios.GetWeightsByNameReq):
        return _get_weights(runner, msg)
    if isinstance(msg, ios.SendWeightsToRemoteInstanceReq):
        return _send_weights_to_remote(runner, msg)
    if isinstance(msg, ios.WeightsCheckerReq):
        return _weights_checker(runner, msg)
    if isinstance(msg, ios.InitWeightsUpdateGroupReq):
        return _init_update_group(runner, msg)
    if isinstance(msg, ios.DestroyWeightsUpdateGroupReq):
        getattr(runner, "_update_groups", {}).pop(msg.group_name, None)
        return ios.GenericResp(True, f"destroyed {msg.group_name}")
    if isinstance(msg, ios.UpdateWeightsFromDistributedReq):
        return _update_from_distributed(runner, msg)
    if isinstance(msg, ios.ReleaseMemoryOccupationReq):
        return _release(runner, msg)
    if isinstance(msg, ios.ResumeMemoryOccupationReq):
        return _resume(runner, msg)
    if isinstance(msg, ios.ProfileReq):
        return _profile(runner, msg)
    if isinstance(msg, ios.LoadLoRAAdapterReq):
        if runner.lora_manager is None:
            return ios.GenericResp(False, "LoRA is not enabled (--enable-lora)")
        runner.lora_manager.load_adapter(msg.lora_name, msg.lora_path)
        return ios.GenericResp(True, f"loaded {msg.lora_name}")
    if isinstance(msg, ios.UnloadLoRAAdapterReq):
        if runner.lora_manager is None:
            return ios.GenericResp(False, "LoRA is not enabled (--enable-lora)")
        runner.lora_manager.unload_adapter(msg.lora_name)
        return ios.GenericResp(True, f"unloaded {msg.lora_name}")
    if isinstance(msg, ios.ExpertDistributionReq):
        from ..eplb.expert_distribution import handle_request

        return handle_request(runner, msg)
    raise TypeError(type(msg))


def _p2p_group(address: str, port: int, name: str, rank: int, cuda: bool, backend: str, is_master: bool):
    import datetime

    import torch.distributed as dist

    store = dist.PrefixStore(name, dist.TCPStore(address, int(port), 2, is_master=is_master, timeout=datetime.timedelta(seconds=120)))
    if cuda and backend == "nccl":
        return dist.ProcessGroupNCCL(store, rank, 2, dist.ProcessGroupNCCL.Options())
    return dist.ProcessGroupGloo(store, rank, 2, datetime.timedelta(seconds=120))


def _send_weights_to_remote(runner: Any, msg: Any) -> Any:
    """remote-instance loader, seed side (reference loader.py:3268 RemoteInstanceModelLoader + /send_weights_to_remote_instance): this rank's
    shards go to the same TP rank of the new instance, parameter by parameter, in ``named_parameters`` order (identical layouts on both sides)"""
    import torch.distributed as dist

    _sync(runner)
    try:
        pg = _p2p_group(msg.master_address, msg.master_port + runner.tp.rank, msg.group_name, 0, runner.is_cuda, msg.backend, False)
        opts = dist.BroadcastOptions()
        opts.rootRank = 0
        n = 0
        for _, p in runner.model.named_parameters():
            t = p.data if p.dtype != torch.float8_e4m3fn else p.data.view(torch.uint8)
            pg.broadcast([t.contiguous()], opts).wait()
            n += 1
        for _, b in runner.model.named_buffers():
            if b.numel() and b.dtype.is_floating_point:
                pg.broadcast([b.data.contiguous()], opts).wait()
        _sync(runner)
        return ios.GenericResp(True, f"sent {n} tensors")
    except Exception as e:
        logger.exception("send_weights_to_remote_instance failed")
        return ios.GenericResp(False, f"{type(e).__name__}: {e}")


def receive_weights_from_seed(model: torch.nn.Module, seed_url: str, tp_rank: int, cuda: bool, port: int, address: str = "127.0.0.1",
                              backend: str = "nccl", group_name: str = "remote_instance_load") -> int:
    """remote-instance loader, new-instance side: host the rendezvous store, ask the seed to start sending (rank 0 posts the request), receive"""
    import json
    import threading
    import urllib.request

    import torch.distributed as dist

    err: list = []

    def ask():
        try:
            body = json.dumps({"master_address": address, "master_port": port, "group_name": group_name, "backend": backend}).encode()
            req = urllib.request.Request(seed_url.rstrip("/") + "/send_weights_to_remote_instance", data=body,
                                         headers={"Content-Type": "application/json"}, method="POST")
            with urllib.request.urlopen(req, timeout=600) as r:
                j = json.loads(r.read())
                if not j.get("success"):
                    err.append(j.get("message"))
        except Exception as e:      # noqa: BLE001
            err.append(repr(e))

    th = None
    if tp_rank == 0:
        th = threading.Thread(target=ask, daemon=True)
        th.start()
    pg = _p2p_group(address, port + tp_rank, group_name, 1, cuda, backend, True)
    opts = dist.BroadcastOptions()
    opts.rootRank = 0
    n = 0
    for _, p in model.named_parameters():
        t = p.data if p.dtype != torch.float8_e4m3fn else p.data.view(torch.uint8)
        buf = torch.empty_like(t)
        pg.broadcast([buf], opts).wait()
        t.copy_(buf)
        n += 1
    for _, b in model.named_buffers():
        if b.numel() and b.dtype.is_floating_point:
            buf = torch.empty_like(b.data)
            pg.broadcast([buf], opts).wait()
            b.data.copy_(buf)
    if th is not None:
        th.join(timeout=600)
    if err:
        raise RuntimeError(f"seed instance failed to send weights: {err[0]}")
    return n


def _weights_checker(runner: Any, msg: Any) -> Any:
    """per-parameter checksums of THIS rank's shards: sum and a position-weighted sum in fp64 (any flipped / permuted element changes it)"""
    sums = {}
    for n, p in runner.model.named_parameters():
        f = p.detach().double().reshape(-1)
        if p.dtype == torch.float8_e4m3fn:
            f = p.detach().view(torch.uint8).double().reshape(-1)
        w = torch.arange(1, f.numel() + 1, dtype=torch.float64, device=f.device) % 8191
        sums[n] = (float(f.sum()), float((f * w).sum()))
    total = float(sum(abs(a) + abs(b) for a, b in sums.values()))
    if msg.action == "snapshot":
        runner._weights_snapshot = sums
        return ios.GenericResp(True, f"snapshot of {len(sums)} tensors (fingerprint {total:.6e})")
    if msg.action == "compare":
        snap = getattr(runner, "_weights_snapshot", None)
        if snap is None:
            return ios.GenericResp(False, "no snapshot to compare against")
        diff = sorted(n for n in sums if snap.get(n) != sums[n])
        return ios.GenericResp(not diff, "identical" if not diff else f"{len(diff)} tensors changed: {diff[:8]}")
    return ios.GenericResp(True, __import__("json").dumps({"fingerprint": total, "num_tensors": len(sums), "checksums": sums}))


def _sync(runner) -> None:
    if runner.is_cuda:
        torch.cuda.synchronize(runner.device)


def _update_from_disk(runner, msg) -> ios.UpdateWeightsResp:
    from ..model_loader.loader import iter_pt, iter_safetensors, process_weights_after_loading

    _sync(runner)
    try:
        it = iter_pt(msg.model_path) if msg.load_format == "pt" else iter_safetensors(msg.model_path)
        n = len(runner.model.load_weights(it))
        process_weights_after_loading(runner.model)
    except Exception as e:
        logger.exception("update_weights_from_disk failed")
        return ios.UpdateWeightsResp(False, f"{type(e).__name__}: {e}")
    _recapture(runner)
    return ios.UpdateWeightsResp(True, f"loaded {n} tensors from {msg.model_path}")


def _update_from_tensor(runner, msg) -> ios.UpdateWeightsResp:
    """named_tensors: list of (name, tensor | nested list).  Full (unsharded) tensors; each rank takes its shard."""
    _sync(runner)
    try:
        items: List[Tuple[str, torch.Tensor]] = []
        for name, t in msg.named_tensors:
            items.append((name, t if isinstance(t, torch.Tensor) else torch.tensor(t)))
        dtype = next(runner.model.parameters()).dtype
        n = len(runner.model.load_weights((k, v.to(dtype)) for k, v in items))
    except Exception as e:
        logger.exception("update_weights_from_tensor failed")
        return ios.UpdateWeightsResp(False, f"{type(e).__name__}: {e}")
    return ios.UpdateWeightsResp(True, f"updated {n} tensors")


def _init_update_group(runner, msg) -> ios.GenericResp:
    """Standalone communicator with an external trainer process (its rank 0 broadcasts fresh weights after every RL step).
    Reference: ``model_runner_components/weight_updater.py`` ``init_weights_update_group`` (custom process group over a TCP store;
    our rank = rank_offset + tp_rank).  Built from a private TCPStore + ProcessGroup{NCCL,Gloo} so it does not touch the serving
    job's default group."""
    import datetime

    import torch.distributed as dist

    rank = msg.rank_offset + runner.tp.rank
    backend = msg.backend if runner.is_cuda else "gloo"
    try:
        store = dist.TCPStore(msg.master_address, int(msg.master_port), msg.world_size, is_master=False,
                              timeout=datetime.timedelta(seconds=120))
        store = dist.PrefixStore(msg.group_name, store)
        if backend == "nccl":
            pg = dist.ProcessGroupNCCL(store, rank, msg.world_size, dist.ProcessGroupNCCL.Options())
        else:
            pg = dist.ProcessGroupGloo(store, rank, msg.world_size, datetime.timedelta(seconds=120))
    except Exception as e:
        logger.exception("init_weights_update_group failed")
        return ios.GenericResp(False, f"{type(e).__name__}: {e}")
    if not hasattr(runner, "_update_groups"):
        runner._update_groups = {}
    runner._update_groups[msg.group_name] = pg
    return ios.GenericResp(True, f"joined {msg.group_name} as rank {rank}/{msg.world_size} ({backend})")


def _update_from_distributed(runner, msg) -> ios.UpdateWeightsResp:
    """receive full (unsharded) tensors by broadcast from the trainer (rank 0) and load each rank's shard"""
    import torch.distributed as dist

    pg = getattr(runner, "_update_groups", {}).get(msg.group_name)
    if pg is None:
        return ios.UpdateWeightsResp(False, f"weight update group {msg.group_name!r} is not initialised")
    _sync(runner)
    dt = {"bfloat16": torch.bfloat16, "float16": torch.float16, "float32": torch.float32}
    try:
        pdtype = next(runner.model.parameters()).dtype
        n = 0
        for name, dtype, shape in zip(msg.names, msg.dtypes, msg.shapes):
            t = torch.empty(tuple(shape), dtype=dt.get(str(dtype).replace("torch.", ""), torch.float32), device=runner.device)
            opts = dist.BroadcastOptions()
            opts.rootRank = 0
            pg.broadcast([t], opts).wait()
            n += len(runner.model.load_weights([(name, t.to(pdtype))]))
    except Exception as e:
        logger.exception("update_weights_from_distributed failed")
        return ios.UpdateWeightsResp(False, f"{type(e).__name__}: {e}")
    return ios.UpdateWeightsResp(True, f"updated {n} tensors over {msg.group_name}")


def _get_weights(runner, msg) -> ios.GetWeightsByNameResp:
    params = dict(runner.model.named_parameters())
    if msg.name not in params:
        return ios.GetWeightsByNameResp(None)
    return ios.GetWeightsByNameResp(params[msg.name].detach().float().flatten()[: msg.truncate_size].cpu().tolist())


def _release(runner, msg) -> ios.GenericResp:
    """free the KV pool (and optionally park the weights on the host) so a co-located trainer can use the HBM"""
    tags = msg.tags or ["kv_cache", "weights"]
    _sync(runner)
    if "kv_cache" in tags and runner.kv_pool.buf is not None:
        runner._released_kv_shape = (tuple(runner.kv_pool.buf.shape), runner.kv_pool.buf.dtype)
        runner.kv_pool.buf = None
        runner.graphs.clear()
    if "weights" in tags:
        runner._parked = {k: p.data.to("cpu", non_blocking=False) for k, p in runner.model.named_parameters()}
        for p in runner.model.parameters():
            p.data = torch.empty(0, dtype=p.dtype, device=p.device)
    if runner.is_cuda:
        torch.cuda.empty_cache()
    return ios.GenericResp(True, f"released {tags}")


def _resume(runner, msg) -> ios.GenericResp:
    tags = msg.tags or ["kv_cache", "weights"]
    if "weights" in tags and getattr(runner, "_parked", None):
        for k, p in runner.model.named_parameters():
            p.data = runner._parked[k].to(runner.device)
        runner._parked = None
    if "kv_cache" in tags and runner.kv_pool.buf is None:
        shape, dt = runner._released_kv_shape
        runner.kv_pool.buf = torch.zeros(shape, dtype=dt, device=runner.device)
        _recapture(runner)
    return ios.GenericResp(True, f"resumed {tags}")


def _recapture(runner) -> None:
    # CUDA graphs hold raw pointers to the KV pool; weights are updated in place so they stay valid
    if runner.is_cuda and not runner.graphs and not runner.args.disable_cuda_graph and runner.pp.world_size == 1:
        with torch.cuda.stream(runner.stream):
            runner._capture_graphs()
        runner.stream.synchronize()


def _profile(runner, msg) -> ios.ProfileResp:
    from torch.profiler import ProfilerActivity, profile

    if msg.action == "start":
        if runner.profiler is not None:
            return ios.ProfileResp(False, "profiler already running")
        acts = [ProfilerActivity.CPU] + ([ProfilerActivity.CUDA] if runner.is_cuda else [])
        out = msg.output_dir or os.environ.get("SGLB_TORCH_PROFILER_DIR", "/tmp/sglb_profiles")
        os.makedirs(out, exist_ok=True)
        runner.profiler = profile(activities=acts, record_shapes=True, with_stack=False)
        runner.profiler.__enter__()
        runner._profile_dir = out
        runner._profile_stop_at = runner.forward_ct + msg.num_steps if msg.num_steps else None
        return ios.ProfileResp(True, f"profiling to {out}")
    if runner.profiler is None:
        return ios.ProfileResp(False, "profiler not running")
    runner.profiler.__exit__(None, None, None)
    from ..parallel.parallel_state import get_state

    path = os.path.join(runner._profile_dir, f"trace-rank{get_state().rank}-{int(time.time())}.json")
    runner.profiler.export_chrome_trace(path)
    runner.profiler = None
    return ios.ProfileResp(True, path)
