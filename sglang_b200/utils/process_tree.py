"""Process-tree hygiene for the multi-process topologies (frontend <-> scheduler <-> detokenizer, TP x PP peers, DP replicas).

Parity: reference ``python/sglang/srt/utils/watchdog.py:166-223`` (``SubprocessWatchdog``: poll child liveness, bring the parent down when a
child dies) and ``kill_process_tree`` / ``kill_itself_when_parent_died`` (``utils/common.py``).  Three pieces:

* ``die_with_parent()``   called first thing in every child: the kernel sends the child a signal when its parent exits, however it exits
  (``prctl(PR_SET_PDEATHSIG)``), so a killed server can not leave GPU-holding orphans behind;
* ``SubprocessWatchdog``  the parent notices a crashed child within a second, logs which one and with what exit code, and terminates itself
  (through the normal SIGTERM path, so the remaining children are shut down in order);
* ``graceful_sigterm()``  SIGTERM becomes ``SystemExit`` in the main thread, so ``finally:`` blocks that shut engines down actually run (uvicorn
  re-raises the signal it captured once its loop has stopped; with the default handler that kills the process before any clean-up).
"""
from __future__ import annotations

import ctypes
import logging
import os
import signal
import sys
import threading
import time
from typing import Any, Callable, List, Optional, Sequence

logger = logging.getLogger(__name__)

PR_SET_PDEATHSIG = 1


def die_with_parent(sig: int = signal.SIGKILL) -> bool:
    """ask the kernel to deliver ``sig`` to this process when its parent dies (Linux); returns False where unsupported"""
    parent = os.getppid()
    try:
        libc = ctypes.CDLL("libc.so.6", use_errno=True)
        if libc.prctl(PR_SET_PDEATHSIG, int(sig), 0, 0, 0) != 0:
            return False
    except OSError:
        return False
    if os.getppid() != parent:           # the parent went away between fork and prctl
        os.kill(os.getpid(), sig)
    return True


def graceful_sigterm() -> None:
    """SIGTERM -> SystemExit(0) in the main thread (no-op outside the main thread, where handlers can not be installed)"""
    if threading.current_thread() is not threading.main_thread():
        return

    def _h(signum, frame):
        raise SystemExit(0)

    signal.signal(signal.SIGTERM, _h)


class SubprocessWatchdog(threading.Thread):
    """watches ``multiprocessing.Process`` / ``subprocess.Popen`` children; the first unexpected death triggers ``on_death`` (default: log and
    SIGTERM ourselves).  ``stop()`` before an orderly shutdown."""

    def __init__(self, procs: Sequence[Any], names: Optional[Sequence[str]] = None, interval_s: float = 0.5,
                 on_death: Optional[Callable[[str, Optional[int]], None]] = None):
        super().__init__(daemon=True, name="sglb-subprocess-watchdog")
        self.procs = list(procs)
        self.names = list(names) if names else [getattr(p, "name", f"child{i}") for i, p in enumerate(self.procs)]
        self.interval_s = interval_s
        self.on_death = on_death or self._default
        self._stop = threading.Event()
        self.dead: List[str] = []

    @staticmethod
    def _alive(p: Any) -> bool:
        if hasattr(p, "is_alive"):
            return p.is_alive()
        return p.poll() is None

    @staticmethod
    def _code(p: Any) -> Optional[int]:
        return getattr(p, "exitcode", None) if hasattr(p, "is_alive") else p.returncode

    def _default(self, name: str, code: Optional[int]) -> None:
        logger.critical("child process %s died unexpectedly (exit code %s): shutting the server down", name, code)
        os.kill(os.getpid(), signal.SIGTERM)

    def stop(self) -> None:
        self._stop.set()

    def run(self) -> None:
        while not self._stop.wait(self.interval_s):
            for p, n in zip(self.procs, self.names):
                if not self._alive(p):
                    if self._stop.is_set():
                        return
                    self.dead.append(n)
                    self.on_death(n, self._code(p))
                    return


def kill_process_tree(procs: Sequence[Any], grace_s: float = 5.0) -> None:
    """terminate exactly these children (never by pattern), escalate to kill after ``grace_s``"""
    for p in procs:
        try:
            p.terminate()
        except Exception:      # noqa: BLE001
            pass
    t_end = time.time() + grace_s
    for p in procs:
        left = max(0.0, t_end - time.time())
        try:
            if hasattr(p, "join"):
                p.join(timeout=left)
                if p.is_alive():
                    p.kill()
            else:
                try:
                    p.wait(timeout=left)
                except Exception:      # noqa: BLE001
                    p.kill()
        except Exception:      # noqa: BLE001
            pass


# ------------------------------------------------------------------------------------------------ pid registry (``cli killall``)
def _pid_dir() -> str:
    d = os.environ.get("SGLB_PID_DIR") or os.path.join(os.environ.get("XDG_RUNTIME_DIR") or "/tmp", "sglb-servers")
    os.makedirs(d, exist_ok=True)
    return d


def register_server(port: int) -> str:
    """top-level server processes record themselves (pid, port, start time) so that ``python -m sglang_b200.cli killall`` can stop exactly the
    servers of this installation — by pid, never by matching command lines (reference ``sglang.cli.killall`` / ``kill_process_tree``)"""
    import atexit
    import json

    path = os.path.join(_pid_dir(), f"{os.getpid()}.json")
    with open(path, "w") as f:
        json.dump({"pid": os.getpid(), "port": int(port), "started": time.time(), "cmd": " ".join(sys.argv[:6])}, f)

    def _rm():
        try:
            os.remove(path)
        except OSError:
            pass
    atexit.register(_rm)
    return path


def list_servers():
    import json

    out = []
    for name in sorted(os.listdir(_pid_dir())):
        if not name.endswith(".json"):
            continue
        path = os.path.join(_pid_dir(), name)
        try:
            rec = json.load(open(path))
            os.kill(int(rec["pid"]), 0)          # still alive?
            out.append(rec)
        except (OSError, ValueError, KeyError):
            try:
                os.remove(path)                   # stale entry
            except OSError:
                pass
    return out


def killall_servers(grace_s: float = 10.0, port: Optional[int] = None) -> int:
    """SIGTERM (graceful drain) every registered server (or the one on ``port``), SIGKILL what is still there after ``grace_s``"""
    victims = [r for r in list_servers() if port is None or r["port"] == port]
    for r in victims:
        try:
            os.kill(int(r["pid"]), signal.SIGTERM)
        except OSError:
            pass
    deadline = time.time() + grace_s
    while time.time() < deadline and any(_alive(int(r["pid"])) for r in victims):
        time.sleep(0.1)
    for r in victims:
        if _alive(int(r["pid"])):
            try:
                os.kill(int(r["pid"]), signal.SIGKILL)
            except OSError:
                pass
    return len(victims)


def _alive(pid: int) -> bool:
    try:
        os.kill(pid, 0)
    except OSError:
        return False
    try:          # a zombie child of ours still answers kill(0)
        with open(f"/proc/{pid}/stat") as f:
            return f.read().rsplit(")", 1)[1].split()[0] != "Z"
    except OSError:
        return True
