"""Watchdogs (reference ``python/sglang/srt/utils/watchdog.py:20-223``): a thread checks that the scheduler's
``forward_ct`` advances while work is pending; on a hard timeout it dumps scheduler state + all thread stacks and kills
the process tree; a soft timeout only logs.  ``SubprocessWatchdog`` polls child liveness."""
from __future__ import annotations

import faulthandler
import logging
import os
import signal
import sys
import threading
import time
from typing import Any, Callable, List, Optional

logger = logging.getLogger(__name__)


class SchedulerWatchdog:
    def __init__(self, scheduler: Any, hard_timeout: float, soft_timeout: Optional[float] = None, interval: float = 1.0,
                 on_hard: Optional[Callable[[], None]] = None):
        self.s, self.hard, self.soft, self.interval = scheduler, hard_timeout, soft_timeout, interval
        self.on_hard = on_hard or self._kill
        self._stop = threading.Event()
        self._t = threading.Thread(target=self._run, name="sglb-watchdog", daemon=True)
        self.fired_soft = 0
        self.fired_hard = 0

    def start(self) -> "SchedulerWatchdog":
        self._t.start()
        return self

    def stop(self) -> None:
        self._stop.set()

    def _run(self) -> None:
        last_ct, last_change = self.s.forward_ct, time.perf_counter()
        warned = False
        while not self._stop.wait(self.interval):
            now = time.perf_counter()
            if self.s.forward_ct != last_ct or not self.s.has_work():
                last_ct, last_change, warned = self.s.forward_ct, now, False
                continue
            stuck = now - last_change
            if self.soft and stuck > self.soft and not warned:
                warned = True
                self.fired_soft += 1
                logger.warning("scheduler made no progress for %.1fs: %s", stuck, self.s.get_internal_state())
            if stuck > self.hard:
                self.fired_hard += 1
                logger.error("watchdog timeout (%.1fs without a forward pass). state=%s", stuck, self.s.get_internal_state())
                faulthandler.dump_traceback(file=sys.stderr, all_threads=True)
                self.on_hard()
                return

    @staticmethod
    def _kill() -> None:
        os.killpg(os.getpgid(0), signal.SIGKILL) if hasattr(os, "killpg") else os._exit(1)


class SubprocessWatchdog:
    def __init__(self, procs: List[Any], interval: float = 1.0, on_dead: Optional[Callable[[Any], None]] = None):
        self.procs, self.interval = procs, interval
        self.on_dead = on_dead or (lambda p: os.kill(os.getpid(), signal.SIGQUIT))
        self._stop = threading.Event()
        threading.Thread(target=self._run, daemon=True).start()

    def stop(self) -> None:
        self._stop.set()

    def _run(self) -> None:
        while not self._stop.wait(self.interval):
            for p in self.procs:
                alive = p.is_alive() if hasattr(p, "is_alive") else p.poll() is None
                if not alive:
                    logger.error("child process %s died", getattr(p, "pid", "?"))
                    self.on_dead(p)
                    return
