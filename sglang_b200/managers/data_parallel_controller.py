"""Data-parallel controller: N independent engine replicas (each a TP group) behind one dispatch front.

Parity: reference ``python/sglang/srt/managers/data_parallel_controller.py:132`` (``launch_dp_schedulers:363``, dispatch
policies ``round_robin:753`` / total_requests / total_tokens, worker liveness tracking ``:234-300``).

Each replica is a process that owns ``tp_size`` GPUs starting at ``base_gpu_id + rank * tp_size`` and runs an ``Engine``
with its scheduler loop on a thread; requests / results travel over ``multiprocessing`` queues as the msgpack-encoded
IPC structs.  Dispatch policies: ``round_robin``, ``total_requests`` (fewest in flight), ``total_tokens`` (fewest
outstanding prompt + budgeted output tokens).  A dead replica is detected (process liveness) and taken out of rotation;
its in-flight requests are failed with an abort finish reason.
"""
from __future__ import annotations

import dataclasses
import logging
import multiprocessing as mp
import os
import queue
import threading
import time
import uuid
from typing import Any, Dict, List, Optional

from ..server_args import ServerArgs

logger = logging.getLogger(__name__)


def _replica_main(args_dict: Dict[str, Any], dp_rank: int, inq: Any, outq: Any) -> None:
    from ..utils.process_tree import die_with_parent

    die_with_parent()
    args = ServerArgs(**args_dict)
    if args.device == "cuda":
        os.environ["CUDA_VISIBLE_DEVICES"] = ",".join(str(args.base_gpu_id + dp_rank * args.tp_size + i) for i in range(args.tp_size))
        args.base_gpu_id = 0
    from ..entrypoints.engine import Engine

    eng = Engine(server_args=args)
    eng.start_background_loop()
    outq.put(("ready", dp_rank, None))

    def pump(rid: str, kw: Dict[str, Any]) -> None:
        try:
            out = eng.generate(**kw)
            outq.put(("done", rid, out))
        except Exception as e:
            outq.put(("error", rid, f"{type(e).__name__}: {e}"))

    while True:
        msg = inq.get()
        if msg is None:
            break
        kind, rid, kw = msg
        if kind == "generate":
            threading.Thread(target=pump, args=(rid, kw), daemon=True).start()
        elif kind == "control":
            try:
                outq.put(("done", rid, getattr(eng, kw["method"])(*kw.get("args", ()))))
            except Exception as e:
                outq.put(("error", rid, str(e)))
    eng.shutdown()


class DataParallelController:
    def __init__(self, server_args: ServerArgs, policy: Optional[str] = None):
        self.args = server_args
        self.policy = policy or server_args.load_balance_method
        self.dp = server_args.dp_size
        ctx = mp.get_context("spawn")
        d = dataclasses.asdict(server_args)
        d = {k: v for k, v in d.items() if k in {f.name for f in dataclasses.fields(ServerArgs)}}
        d["dp_size"] = 1
        self._ctx, self._replica_args = ctx, d
        self.inqs = [ctx.Queue() for _ in range(self.dp)]
        self.outq = ctx.Queue()
        self.procs = [ctx.Process(target=_replica_main, args=(d, r, self.inqs[r], self.outq), daemon=True) for r in range(self.dp)]
        for p in self.procs:
            p.start()
        self.alive = [True] * self.dp
        self.inflight_reqs = [0] * self.dp
        self.inflight_tokens = [0] * self.dp
        self._rr = 0
        self._waiters: Dict[str, "queue.Queue[Any]"] = {}
        self._owner: Dict[str, int] = {}
        self._cost: Dict[str, int] = {}
        self._lock = threading.Lock()
        ready = 0
        t0 = time.time()
        while ready < self.dp:
            try:
                kind, r, _ = self.outq.get(timeout=600)
            except queue.Empty:
                raise RuntimeError("data-parallel replicas failed to start")
            if kind == "ready":
                ready += 1
        self._stop = threading.Event()
        threading.Thread(target=self._collect, daemon=True).start()
        threading.Thread(target=self._liveness, daemon=True).start()
        logger.info("%d DP replicas ready in %.1fs", self.dp, time.time() - t0)

    # ------------------------------------------------------------------ dispatch
    def _pick(self, bootstrap_room: Optional[int] = None) -> int:
        live = [i for i in range(self.dp) if self.alive[i]]
        if not live:
            raise RuntimeError("no live data-parallel replica")
        if self.policy == "follow_bootstrap_room" and bootstrap_room is not None:
            # PD disaggregation: the prefill and the decode fleet must pick the SAME dp rank for a request without talking to each other
            # (reference data_parallel_controller follow_bootstrap_room): the room id decides
            return live[int(bootstrap_room) % len(live)]
        if self.policy == "total_requests":
            return min(live, key=lambda i: self.inflight_reqs[i])
        if self.policy == "total_tokens":
            return min(live, key=lambda i: self.inflight_tokens[i])
        self._rr += 1
        return live[self._rr % len(live)]

    def generate(self, prompt: Any = None, sampling_params: Optional[Dict] = None, input_ids: Any = None, routed_dp_rank: Optional[int] = None,
                 **kw: Any) -> Any:
        rid = uuid.uuid4().hex
        cost = (len(input_ids) if input_ids is not None and input_ids and isinstance(input_ids[0], int) else len(prompt or "")) + \
            int((sampling_params or {}).get("max_new_tokens", 128))
        box: "queue.Queue[Any]" = queue.Queue()
        with self._lock:
            r = routed_dp_rank if routed_dp_rank is not None and 0 <= routed_dp_rank < self.dp and self.alive[routed_dp_rank] \
                else self._pick(kw.get("bootstrap_room"))
            self._waiters[rid], self._owner[rid], self._cost[rid] = box, r, cost
            self.inflight_reqs[r] += 1
            self.inflight_tokens[r] += cost
        self.inqs[r].put(("generate", rid, dict(prompt=prompt, sampling_params=sampling_params, input_ids=input_ids, **kw)))
        kind, out = box.get()
        if kind == "error":
            raise RuntimeError(out)
        return out

    def control(self, method: str, *args: Any) -> List[Any]:
        outs = []
        for r in range(self.dp):
            if not self.alive[r]:
                continue
            rid = uuid.uuid4().hex
            box: "queue.Queue[Any]" = queue.Queue()
            with self._lock:
                self._waiters[rid], self._owner[rid], self._cost[rid] = box, r, 0
            self.inqs[r].put(("control", rid, {"method": method, "args": args}))
            outs.append(box.get()[1])
        return outs

    def _collect(self) -> None:
        while not self._stop.is_set():
            try:
                kind, rid, out = self.outq.get(timeout=0.2)
            except queue.Empty:
                continue
            if kind == "ready":             # a replica added / restarted at run time
                self._pending_ready = getattr(self, "_pending_ready", set())
                self._pending_ready.add(rid)
                continue
            with self._lock:
                box = self._waiters.pop(rid, None)
                r = self._owner.pop(rid, None)
                c = self._cost.pop(rid, 0)
                if r is not None and c:
                    self.inflight_reqs[r] -= 1
                    self.inflight_tokens[r] -= c
            if box is not None:
                box.put((kind, out))

    def _liveness(self) -> None:
        while not self._stop.wait(0.5):
            for r, p in enumerate(self.procs):
                if self.alive[r] and not p.is_alive():
                    self.alive[r] = False
                    logger.error("DP replica %d died; failing its in-flight requests", r)
                    with self._lock:
                        dead = [rid for rid, o in self._owner.items() if o == r]
                        for rid in dead:
                            box = self._waiters.pop(rid, None)
                            self._owner.pop(rid, None)
                            self._cost.pop(rid, None)
                            if box is not None:
                                box.put(("error", f"replica {r} died"))

    # ------------------------------------------------------------------ elastic replicas (reference elastic workers :234-300)
    def add_replica(self, timeout: float = 600.0) -> int:
        """start one more replica at run time; returns its dp rank"""
        with self._lock:
            r = self.dp
            self.inqs.append(self._ctx.Queue())
            p = self._ctx.Process(target=_replica_main, args=(self._replica_args, r, self.inqs[r], self.outq), daemon=True)
            self.procs.append(p)
            self.alive.append(False)
            self.inflight_reqs.append(0)
            self.inflight_tokens.append(0)
            self.dp += 1
        self._pending_ready = getattr(self, "_pending_ready", set())
        p.start()
        t0 = time.time()
        while r not in self._pending_ready:          # the collector thread sees the "ready" message
            if time.time() - t0 > timeout or not p.is_alive():
                raise RuntimeError(f"replica {r} failed to start")
            time.sleep(0.05)
        self.alive[r] = True
        return r

    def remove_replica(self, r: int, drain_timeout: float = 60.0) -> None:
        """stop routing to replica ``r``, let its in-flight requests finish, then stop the process"""
        self.alive[r] = False
        t0 = time.time()
        while self.inflight_reqs[r] > 0 and time.time() - t0 < drain_timeout:
            time.sleep(0.05)
        self.inqs[r].put(None)
        self.procs[r].join(timeout=10)
        if self.procs[r].is_alive():
            self.procs[r].terminate()

    def restart_replica(self, r: int, timeout: float = 600.0) -> None:
        """bring a dead replica back under the same rank"""
        assert not self.procs[r].is_alive(), "replica still running"
        self.inqs[r] = self._ctx.Queue()
        p = self._ctx.Process(target=_replica_main, args=(self._replica_args, r, self.inqs[r], self.outq), daemon=True)
        self.procs[r] = p
        self._pending_ready = getattr(self, "_pending_ready", set())
        self._pending_ready.discard(r)
        p.start()
        t0 = time.time()
        while r not in self._pending_ready:
            if time.time() - t0 > timeout or not p.is_alive():
                raise RuntimeError(f"replica {r} failed to restart")
            time.sleep(0.05)
        self.inflight_reqs[r] = self.inflight_tokens[r] = 0
        self.alive[r] = True

    def shutdown(self) -> None:
        self._stop.set()
        for q_ in self.inqs:
            q_.put(None)
        for p in self.procs:
            p.join(timeout=10)
            if p.is_alive():
                p.terminate()
