"""Multi-turn sessions: a request may continue from an earlier request of the same session without the client re-sending (or the
engine re-computing) the history.

Parity: reference ``python/sglang/srt/session/session_controller.py`` (``Session.create_req``: ``rid`` selects the parent turn,
``offset`` truncates it, ``replace`` branches from the parent's parent, ``drop_previous_output`` keeps only the parent's input) and
``io_struct.SessionParams``.  Differences by design: the history lives as token ids here and as KV pages in the radix tree; while a
session is open the tree node of its newest turn is pinned (``inc_lock_ref``) so eviction cannot drop the conversation's prefix —
closing the session (or superseding the turn) unpins it.
"""
from __future__ import annotations

from collections import OrderedDict
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Tuple


@dataclass
class SessionTurn:
    rid: str
    parent: Optional[str]
    input_ids: List[int]                 # full prompt of this turn (history + new input)
    output_ids: List[int] = field(default_factory=list)
    finished: bool = False
    pinned_node: Any = None


@dataclass
class Session:
    session_id: str
    capacity: int = 0
    turns: Dict[str, SessionTurn] = field(default_factory=dict)
    last_rid: Optional[str] = None


class SessionController:
    """``max_sessions`` bounds open sessions (the route may be reachable by any API-key holder); ``max_pinned`` bounds how many sessions keep
    their newest turn PINNED in the radix tree — beyond it the least-recently-finished pin is dropped (the history stays as token ids and is
    simply recomputed / re-matched), and ``unpin_lru`` lets the scheduler reclaim pins under KV pressure before it retracts or aborts."""

    def __init__(self, tree: Any, page_size: int, max_sessions: int = 4096, max_pinned: int = 256):
        self.tree, self.ps = tree, page_size
        self.sessions: Dict[str, Session] = {}
        self.max_sessions, self.max_pinned = max_sessions, max_pinned
        self._pin_order: "OrderedDict[str, SessionTurn]" = OrderedDict()     # session id -> pinned turn, oldest first
        self._n_opened = 0

    # ---- lifecycle
    def open(self, session_id: Optional[str], capacity: int = 0) -> Tuple[Optional[str], bool]:
        if session_id is None:
            self._n_opened += 1
        sid = session_id or f"sess-{self._n_opened}"
        if sid in self.sessions:
            return sid, False
        if len(self.sessions) >= self.max_sessions:
            return None, False
        self.sessions[sid] = Session(sid, capacity)
        return sid, True

    def close(self, session_id: str) -> None:
        s = self.sessions.pop(session_id, None)
        self._pin_order.pop(session_id, None)
        if s is not None:
            for t in s.turns.values():
                self._unpin(t)

    def __contains__(self, sid: str) -> bool:
        return sid in self.sessions

    # ---- request construction
    def build_input(self, params: Dict[str, Any], new_ids: List[int], rid: str) -> Tuple[Optional[List[int]], Optional[str]]:
        """returns (full input ids, error).  ``params``: {id, rid?, offset?, replace?, drop_previous_output?}"""
        s = self.sessions.get(params.get("id"))
        if s is None:
            return None, f"session {params.get('id')!r} does not exist"
        parent_rid = params.get("rid")
        if parent_rid is None:
            parent_rid = s.last_rid
        elif parent_rid not in s.turns:
            return None, f"request {parent_rid!r} is not part of session {s.session_id!r}"
        parent = s.turns.get(parent_rid) if parent_rid is not None else None
        if parent is not None and not parent.finished:
            return None, f"request {parent_rid!r} of session {s.session_id!r} is still running"
        if parent is not None and params.get("replace"):
            parent = s.turns.get(parent.parent) if parent.parent else None      # branch: sibling of the replaced turn
        hist: List[int] = []
        if parent is not None:
            hist = list(parent.input_ids)
            if not params.get("drop_previous_output"):
                out = parent.output_ids
                off = params.get("offset")
                if off is not None and off != -1:
                    out = out[:off] if off >= 0 else out[:off + 1]
                hist += list(out)
        full = hist + list(new_ids)
        if s.capacity and len(full) > s.capacity:
            return None, f"session {s.session_id!r} exceeded its capacity ({len(full)} > {s.capacity} tokens)"
        s.turns[rid] = SessionTurn(rid, parent.rid if parent is not None else None, full)
        return full, None

    # ---- completion hooks
    def on_finished(self, session_id: str, rid: str, output_ids: List[int], extra_key: Any = None) -> None:
        s = self.sessions.get(session_id)
        if s is None or rid not in s.turns:
            return
        t = s.turns[rid]
        t.output_ids, t.finished = list(output_ids), True
        # pin the newest turn's prefix in the radix tree; the previous newest turn no longer needs its own pin
        if s.last_rid is not None and s.last_rid in s.turns and s.last_rid != rid:
            self._unpin(s.turns[s.last_rid])
        s.last_rid = rid
        ids = t.input_ids + t.output_ids
        n = (len(ids) // self.ps) * self.ps
        if n and not getattr(self.tree, "disable", False):
            try:
                pages, node = self.tree.match_prefix(ids[:n], extra_key)
                if len(pages) and node is not None:
                    self.tree.inc_lock_ref(node)
                    t.pinned_node = node
                    self._pin_order.pop(session_id, None)
                    self._pin_order[session_id] = t
                    while len(self._pin_order) > self.max_pinned:
                        _, old = self._pin_order.popitem(last=False)
                        self._unpin(old)
            except Exception:      # pinning is an optimisation; the history itself is kept as token ids
                t.pinned_node = None

    def on_aborted(self, session_id: str, rid: str) -> None:
        s = self.sessions.get(session_id)
        if s is not None:
            t = s.turns.pop(rid, None)
            if t is not None:
                self._unpin(t)

    def _unpin(self, t: SessionTurn) -> None:
        if t.pinned_node is not None:
            try:
                self.tree.dec_lock_ref(t.pinned_node)
            finally:
                t.pinned_node = None

    def unpin_lru(self, n: int = 1) -> int:
        """KV pressure: drop the ``n`` oldest session pins (their pages become evictable); returns how many were dropped"""
        k = 0
        while self._pin_order and k < n:
            _, old = self._pin_order.popitem(last=False)
            self._unpin(old)
            k += 1
        return k

    def unpin_all(self) -> None:
        """cache flush: the tree is about to be reset, drop every pin (token histories stay)"""
        self._pin_order.clear()
        for s in self.sessions.values():
            for t in s.turns.values():
                t.pinned_node = None
