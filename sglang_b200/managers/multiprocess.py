"""Multi-process serving topology: front end (HTTP + tokenizer + request states) | scheduler process(es) | detokenizer process.

Parity: reference ``python/sglang/srt/entrypoints/engine.py:905-1003`` (``_launch_subprocesses``: scheduler procs + detokenizer proc +
TokenizerManager in the main process, wired with ZMQ), ``managers/tokenizer_manager.py:386,765`` and ``managers/detokenizer_manager.py``.

Why it exists here: the default topology keeps one process per GPU (tokenizer, detokenizer, request states and the scheduler loop share it —
zero IPC for the offline ``Engine``).  Behind an HTTP server that means FastAPI handlers, tokenisation, detokenisation and SSE serialisation
all contend for the scheduler thread's GIL.  ``--enable-multi-process-frontend`` moves everything that is not scheduling out of that process:

    front end (this process)  --PUSH tokenised requests-->  scheduler process (rank 0; spawns / drives its TP x PP peers as before)
          ^                                                      |
          |                                                      +--PUSH BatchTokenIDOutput-->  detokenizer process
          +----------------PULL BatchStrOutput--------------------------------------------------------+
    control RPCs: front end DEALER <-> scheduler ROUTER (request id correlated, any number in flight)

Messages are the tagged msgpack frames of ``managers/io_struct.py`` (no pickles on the wire); sockets are ``ipc://`` files in a private
temp dir.  ``FrontendEngine`` subclasses ``Engine`` and overrides exactly the three seams that touch the scheduler (``_submit``,
``_control`` and the drive loop), so every public method, the OpenAI layer and the adapters work unchanged on top of it.
"""
from __future__ import annotations

import itertools
import logging
import multiprocessing as mp
import os
import queue
import tempfile
import threading
import time
from typing import Any, Dict, List, Optional

from . import io_struct as ios
from ..server_args import ServerArgs

logger = logging.getLogger(__name__)


def _ipc_names(num_workers: int = 1) -> Dict[str, str]:
    d = tempfile.mkdtemp(prefix="sglb-ipc-")
    names = {k: f"ipc://{d}/{k}" for k in ("req", "tok", "str", "ctl")}
    for i in range(num_workers):          # one result socket per tokenizer worker (reference: MultiTokenizerRouter)
        names[f"str{i}"] = f"ipc://{d}/str{i}"
    return names


WORKER_SEP = "~"        # request ids travel as "<worker>~<rid>" between a tokenizer worker and the scheduler; the detokenizer routes on the prefix


def split_by_worker(out: "ios.BatchStrOutput", num_workers: int) -> Dict[int, "ios.BatchStrOutput"]:
    """one BatchStrOutput per tokenizer worker: every per-request list is sliced by the worker prefix of its request id (prefix removed)"""
    import dataclasses

    groups: Dict[int, List[int]] = {}
    clean: List[str] = []
    for i, rid in enumerate(out.rids):
        w, sep, rest = rid.partition(WORKER_SEP)
        if sep and w.isdigit() and int(w) < num_workers:
            groups.setdefault(int(w), []).append(i)
            clean.append(rest)
        else:
            groups.setdefault(0, []).append(i)
            clean.append(rid)
    n = len(out.rids)
    res = {}
    for w, idx in groups.items():
        kw = {}
        for f in dataclasses.fields(out):
            v = getattr(out, f.name)
            if f.name == "rids":
                kw[f.name] = [clean[i] for i in idx]
            elif isinstance(v, list) and len(v) == n:
                kw[f.name] = [v[i] for i in idx]
            else:
                kw[f.name] = v
        res[w] = type(out)(**kw)
    return res


# ------------------------------------------------------------------------------------------------------------------ scheduler process
def scheduler_process_main(server_args: ServerArgs, ipc: Dict[str, str], ready: Any) -> None:
    """rank-0 scheduler process: a regular ``Engine`` (it spawns / joins its TP x PP peers exactly as the single-process server does)
    whose input comes from ZMQ and whose token-id outputs go to the detokenizer process"""
    import zmq

    from ..entrypoints.engine import Engine

    from ..utils.process_tree import die_with_parent

    die_with_parent()
    try:
        os.environ["SGLB_IN_SCHEDULER_PROCESS"] = "1"
        from ..entrypoints.http_server import spawn_peers

        peers = spawn_peers(server_args)      # TP x PP ranks > 0 (sets RANK / MASTER_* for this process when world > 1)
        eng = Engine(server_args=server_args)
        ctx = zmq.Context(2)
        s_req = ctx.socket(zmq.PULL); s_req.bind(ipc["req"])
        s_tok = ctx.socket(zmq.PUSH); s_tok.connect(ipc["tok"])
        s_ctl = ctx.socket(zmq.ROUTER); s_ctl.bind(ipc["ctl"])
        send_lock = threading.Lock()

        def forward(msg: ios.BatchTokenIDOutput, relayed: bool = False) -> None:
            s_tok.send(ios.encode(msg))      # called from the scheduler thread only

        eng.scheduler.send_to_detokenizer = forward      # raw token ids leave the process; text is made elsewhere
        eng.start_background_loop()
        ready.put({"ok": True, "memory": eng.runner.get_memory_info(), "max_context_len": eng.scheduler.max_context_len,
                   "weight_bytes": int(eng.runner.weight_bytes)})
    except Exception as e:      # noqa: BLE001
        logger.exception("scheduler process failed to start")
        ready.put({"ok": False, "error": f"{type(e).__name__}: {e}"})
        return

    def run_ctl(ident: bytes, cid: bytes, msg: Any) -> None:
        try:
            rep = eng._control(msg)
        except Exception as e:      # noqa: BLE001
            rep = ios.GenericResp(False, f"{type(e).__name__}: {e}")
        with send_lock:
            s_ctl.send_multipart([ident, cid, ios.encode(rep)])

    poller = zmq.Poller()
    poller.register(s_req, zmq.POLLIN)
    poller.register(s_ctl, zmq.POLLIN)
    while True:
        for sock, _ in poller.poll(1000):
            if sock is s_req:
                while True:
                    try:
                        buf = s_req.recv(zmq.NOBLOCK)
                    except zmq.Again:
                        break
                    eng._inbox.put(ios.decode(buf))
                eng._wake.set()
            else:
                ident, cid, buf = s_ctl.recv_multipart()
                msg = ios.decode(buf)
                if isinstance(msg, ios.ShutdownReq):
                    with send_lock:
                        s_ctl.send_multipart([ident, cid, ios.encode(ios.GenericResp(True, "bye"))])
                    s_tok.send(ios.encode(ios.ShutdownReq()))      # the detokenizer leaves too
                    eng.shutdown()
                    for p in peers:
                        p.terminate()
                    return
                # control messages may block (drain, weight load): never stall the request socket behind them
                threading.Thread(target=run_ctl, args=(ident, cid, msg), daemon=True).start()
        if eng._bg_thread is not None and not eng._bg_thread.is_alive():
            logger.error("scheduler loop died; leaving the scheduler process")
            return


# ------------------------------------------------------------------------------------------------------------------ detokenizer process
def detokenizer_process_main(server_args: ServerArgs, ipc: Dict[str, str], ready: Any) -> None:
    import zmq

    from ..configs.model_config import load_model_config
    from ..utils.tokenizer import get_tokenizer
    from .detokenizer import Detokenizer
    from ..utils.process_tree import die_with_parent

    die_with_parent()
    try:
        cfg = load_model_config(server_args.model_path, None, server_args.context_length)
        tok = get_tokenizer(server_args.tokenizer_path, cfg.vocab_size, cfg.eos_token_id, cfg.bos_token_id, server_args.trust_remote_code)
        detok = Detokenizer(tok)
        ctx = zmq.Context(1)
        s_tok = ctx.socket(zmq.PULL); s_tok.bind(ipc["tok"])
        n_workers = max(1, int(getattr(server_args, "tokenizer_worker_num", 1) or 1))
        if n_workers == 1:
            outs = {0: ctx.socket(zmq.PUSH)}
            outs[0].connect(ipc["str"])
        else:
            outs = {}
            for i in range(n_workers):
                outs[i] = ctx.socket(zmq.PUSH)
                outs[i].connect(ipc[f"str{i}"])
        ready.put({"ok": True})
    except Exception as e:      # noqa: BLE001
        ready.put({"ok": False, "error": f"{type(e).__name__}: {e}"})
        return
    while True:
        msg = ios.decode(s_tok.recv())
        if isinstance(msg, ios.ShutdownReq):
            return
        out = detok.process(msg)
        if n_workers == 1:
            outs[0].send(ios.encode(out))
        else:
            for w, part in split_by_worker(out, n_workers).items():
                outs[w].send(ios.encode(part))


# ------------------------------------------------------------------------------------------------------------------ front end
class _Sentinel:
    """stands in for the scheduler thread object: the base class only asks ``is not None``"""

    def join(self, timeout: Optional[float] = None) -> None:
        pass

    def is_alive(self) -> bool:
        return True


class FrontendEngine:
    """Factory: builds an ``Engine`` subclass instance whose scheduler lives in another process"""

    def __new__(cls, server_args: Optional[ServerArgs] = None, attach: Optional[Dict[str, Any]] = None, **kwargs: Any):
        """``attach`` = {"ipc": names, "worker_id": i, "sched_info": {...}}: connect to a scheduler / detokenizer that a supervisor already started
        (tokenizer worker i of ``--tokenizer-worker-num``) instead of spawning them"""
        from ..entrypoints.engine import Engine

        class _Frontend(Engine):
            def __init__(self, a: ServerArgs):
                import zmq

                from ..configs.model_config import load_model_config
                from ..utils.tokenizer import get_tokenizer

                self.server_args = a
                logging.basicConfig(level=getattr(logging, a.log_level.upper(), logging.INFO), format="[%(asctime)s %(name)s] %(message)s")
                self.model_config = load_model_config(a.model_path, __import__("json").loads(a.json_model_override_args or "{}") or None,
                                                      a.context_length)
                self.tokenizer = None if a.skip_tokenizer_init else get_tokenizer(
                    a.tokenizer_path, self.model_config.vocab_size, self.model_config.eos_token_id, self.model_config.bos_token_id,
                    a.trust_remote_code)
                self.detok = None
                self.is_driver, self.tp_rank, self.pp_rank = True, 0, 0
                self._dp, self._dp_rank, self._dp_ctr = 1, 0, 0
                self._relay_src, self._relay_buf = None, []
                self._world = _SoloWorld()
                self.states: Dict[str, Any] = {}
                self._lock = threading.RLock()
                self._inbox = queue.Queue()
                self._bg_stop, self._wake = threading.Event(), threading.Event()
                self._bg_thread = _Sentinel()
                self.tracer = None
                if a.enable_trace:
                    from ..observability.trace import get_tracer, init_tracing

                    ep = a.otlp_traces_endpoint if str(a.otlp_traces_endpoint or "").startswith("http") else None
                    self.tracer = get_tracer() or init_tracing(level=int(a.trace_level or 2), file=a.trace_file, otlp_endpoint=ep)
                self.metrics = None
                if a.enable_metrics:
                    from ..observability.metrics import MetricsCollector

                    self.metrics = MetricsCollector(a.served_model_name)
                self.weight_version = "default"
                self.start_time = time.time()
                self.runner = _RunnerView(self)
                self.scheduler = _SchedulerView(self)
                # ---- processes and sockets
                self._worker_id = None
                self._ctx = zmq.Context(2)
                if attach is not None:          # tokenizer worker: the supervisor owns the scheduler and the detokenizer
                    self._ipc, self._worker_id, self._sched_info = attach["ipc"], int(attach["worker_id"]), attach["sched_info"]
                    self._s_str = self._ctx.socket(zmq.PULL); self._s_str.bind(self._ipc[f"str{self._worker_id}"])
                    self._p_detok = self._p_sched = None
                else:
                    self._ipc = _ipc_names()
                    spawn = mp.get_context("spawn")
                    rq_s, rq_d = spawn.Queue(), spawn.Queue()
                    self._s_str = self._ctx.socket(zmq.PULL); self._s_str.bind(self._ipc["str"])
                    self._p_detok = spawn.Process(target=detokenizer_process_main, args=(a, self._ipc, rq_d), daemon=True, name="sglb-detokenizer")
                    self._p_detok.start()
                    r = rq_d.get(timeout=600)
                    if not r.get("ok"):
                        raise RuntimeError(f"detokenizer process failed: {r.get('error')}")
                    self._p_sched = spawn.Process(target=scheduler_process_main, args=(a, self._ipc, rq_s), daemon=False, name="sglb-scheduler")
                    self._p_sched.start()
                    r = rq_s.get(timeout=3600)
                    if not r.get("ok"):
                        raise RuntimeError(f"scheduler process failed: {r.get('error')}")
                    self._sched_info = r
                self._s_req = self._ctx.socket(zmq.PUSH); self._s_req.connect(self._ipc["req"])
                self._s_ctl = self._ctx.socket(zmq.DEALER); self._s_ctl.connect(self._ipc["ctl"])
                self._req_lock, self._ctl_lock = threading.Lock(), threading.Lock()
                self._ctl_ids = itertools.count(1)
                self._ctl_wait: Dict[bytes, "queue.Queue[Any]"] = {}
                self._closed = False
                threading.Thread(target=self._recv_outputs, daemon=True, name="sglb-front-out").start()
                threading.Thread(target=self._recv_ctl, daemon=True, name="sglb-front-ctl").start()

            # ---- seams
            def _tag(self, rid: str) -> str:
                return rid if self._worker_id is None or not rid else f"{self._worker_id}{WORKER_SEP}{rid}"

            def _dispatch(self, t: ios.TokenizedGenerateReqInput) -> None:
                if self._worker_id is not None:          # the scheduler sees worker-qualified ids; results come back already stripped
                    import dataclasses

                    t = dataclasses.replace(t, rid=self._tag(t.rid))
                with self._req_lock:
                    self._s_req.send(ios.encode(t))

            def _control(self, msg: Any) -> Any:
                if self._worker_id is not None and isinstance(msg, ios.AbortReq) and msg.rid:
                    import dataclasses

                    msg = dataclasses.replace(msg, rid=self._tag(msg.rid))
                if self._worker_id is not None and isinstance(msg, ios.ShutdownReq):
                    return ios.GenericResp(True, "worker detached")      # only the supervisor stops the shared scheduler
                cid = str(next(self._ctl_ids)).encode()
                box: "queue.Queue[Any]" = queue.Queue()
                self._ctl_wait[cid] = box
                with self._ctl_lock:
                    self._s_ctl.send_multipart([cid, ios.encode(msg)])
                rep = box.get()
                self._ctl_wait.pop(cid, None)
                return rep

            def _recv_ctl(self) -> None:
                import zmq

                while not self._closed:
                    try:
                        if not self._s_ctl.poll(500):
                            continue
                        cid, buf = self._s_ctl.recv_multipart()
                    except zmq.ZMQError:
                        return
                    box = self._ctl_wait.get(cid)
                    if box is not None:
                        box.put(ios.decode(buf))

            def _recv_outputs(self) -> None:
                import zmq

                while not self._closed:
                    try:
                        if not self._s_str.poll(500):
                            continue
                        out = ios.decode(self._s_str.recv())
                    except zmq.ZMQError:
                        return
                    try:
                        self._apply_output(out)
                    except Exception:      # noqa: BLE001
                        logger.exception("applying a detokenizer batch failed")

            def start_background_loop(self) -> None:
                pass

            def _drive(self, states: List[Any]) -> None:
                for st in states:
                    st.done.wait()

            def dump_requests_before_crash(self) -> Optional[str]:
                return None

            def shutdown(self) -> None:
                if self._closed:
                    return
                try:
                    self._control(ios.ShutdownReq())
                except Exception:      # noqa: BLE001
                    pass
                self._closed = True
                for p in (self._p_sched, self._p_detok):
                    if p is None:
                        continue
                    p.join(timeout=10)
                    if p.is_alive():
                        p.terminate()
                        p.join(timeout=5)
                self._ctx.destroy(linger=0)

        return _Frontend(server_args or ServerArgs(**kwargs))


class _SoloWorld:
    """the front end is not a rank of the model-parallel world"""

    world_size = 1


class _RunnerView:
    """the runner facts the API layer reports (sizes), taken from the scheduler process's start-up hand-shake"""

    def __init__(self, front: Any):
        self._f = front
        self.model = None
        self.lora_manager = None

    @property
    def weight_bytes(self) -> int:
        return int(self._f._sched_info["weight_bytes"])

    @property
    def max_context_len(self) -> int:
        return int(self._f._sched_info["max_context_len"])

    def get_memory_info(self) -> Dict[str, int]:
        return dict(self._f._sched_info["memory"])


class _SchedulerView:
    """the few scheduler attributes the HTTP layer reads, answered over the control channel"""

    def __init__(self, front: Any):
        self._f = front

    @property
    def max_context_len(self) -> int:
        return int(self._f._sched_info["max_context_len"])

    def get_internal_state(self) -> Dict[str, Any]:
        return self._f._control(ios.GetInternalStateReq()).state

    def has_work(self) -> bool:
        return bool(self._f.states)

    @property
    def tree(self) -> Any:
        raise AttributeError("the radix tree lives in the scheduler process (use /get_server_info or the KV-event stream)")


# ------------------------------------------------------------------------------------------------------------------ multi-tokenizer serving
def _http_worker_main(server_args: ServerArgs, attach: Dict[str, Any]) -> None:
    """one HTTP + tokenizer worker: its own front end (request states, tokenizer, OpenAI layer) attached to the shared scheduler; all workers
    listen on the SAME port with SO_REUSEPORT, the kernel spreads incoming connections"""
    import socket

    import uvicorn

    from ..entrypoints.http_server import create_app
    from ..utils.process_tree import die_with_parent, graceful_sigterm

    die_with_parent()
    graceful_sigterm()
    engine = FrontendEngine(server_args, attach=attach)
    app = create_app(engine)
    wid = attach["worker_id"]

    @app.middleware("http")
    async def _tag(request, call_next):          # lets a client (and the tests) see which worker answered
        resp = await call_next(request)
        resp.headers["x-sglb-worker"] = str(wid)
        return resp

    if getattr(server_args, "http_frontend", "uvicorn") == "native":
        # every worker runs its own C++ epoll front end on the shared port (SO_REUSEPORT: the kernel spreads new connections)
        from ..entrypoints.native_http import NativeAsgiServer

        try:
            NativeAsgiServer(app, server_args.host, server_args.port, reuse_port=True,
                             graceful_timeout=max(1.0, float(server_args.graceful_shutdown_timeout))).run()
        finally:
            engine.shutdown()
        return
    sock = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
    sock.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
    sock.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEPORT, 1)
    sock.bind((server_args.host, server_args.port))
    sock.listen(2048)
    cfg = uvicorn.Config(app, log_level=server_args.log_level, timeout_keep_alive=5,
                         timeout_graceful_shutdown=max(1, int(server_args.graceful_shutdown_timeout)))
    try:
        uvicorn.Server(cfg).run(sockets=[sock])
    finally:
        engine.shutdown()


def serve_multi_tokenizer(server_args: ServerArgs) -> None:
    """supervisor of ``--tokenizer-worker-num N`` (reference tokenizer_manager multi-worker mode + MultiTokenizerRouter): start the detokenizer
    (which routes results by worker), the scheduler process, then N HTTP workers; watch all of them, tear everything down on SIGTERM or when
    one dies"""
    import signal
    import zmq

    from ..utils.process_tree import SubprocessWatchdog, graceful_sigterm, kill_process_tree

    n = int(server_args.tokenizer_worker_num)
    ipc = _ipc_names(n)
    spawn = mp.get_context("spawn")
    rq_s, rq_d = spawn.Queue(), spawn.Queue()
    p_detok = spawn.Process(target=detokenizer_process_main, args=(server_args, ipc, rq_d), daemon=True, name="sglb-detokenizer")
    p_detok.start()
    r = rq_d.get(timeout=600)
    if not r.get("ok"):
        raise RuntimeError(f"detokenizer process failed: {r.get('error')}")
    p_sched = spawn.Process(target=scheduler_process_main, args=(server_args, ipc, rq_s), daemon=False, name="sglb-scheduler")
    p_sched.start()
    info = rq_s.get(timeout=3600)
    if not info.get("ok"):
        kill_process_tree([p_detok])
        raise RuntimeError(f"scheduler process failed: {info.get('error')}")
    workers = []
    for i in range(n):
        p = spawn.Process(target=_http_worker_main, args=(server_args, {"ipc": ipc, "worker_id": i, "sched_info": info}), daemon=False,
                          name=f"sglb-http-{i}")
        p.start()
        workers.append(p)
    logger.info("multi-tokenizer serving: %d HTTP workers on %s:%d, one scheduler, one detokenizer", n, server_args.host, server_args.port)
    graceful_sigterm()
    stop = threading.Event()
    dog = SubprocessWatchdog([p_sched, p_detok] + workers, on_death=lambda name, code: (logger.critical("%s died (exit code %s)", name, code), stop.set()))
    dog.start()
    try:
        while not stop.wait(0.5):
            pass
    except (KeyboardInterrupt, SystemExit):
        pass
    finally:
        dog.stop()
        for p in workers:                 # workers first: they drain their in-flight HTTP requests on SIGTERM
            if p.is_alive():
                p.terminate()
        for p in workers:
            p.join(timeout=max(2.0, server_args.graceful_shutdown_timeout + 2))
        try:                              # then the shared back end, through its own control socket
            ctx = zmq.Context(1)
            s = ctx.socket(zmq.DEALER); s.connect(ipc["ctl"])
            s.send_multipart([b"0", ios.encode(ios.ShutdownReq())])
            s.poll(5000)
            ctx.destroy(linger=0)
        except Exception:      # noqa: BLE001
            pass
        kill_process_tree([p_sched, p_detok] + workers, grace_s=5.0)
