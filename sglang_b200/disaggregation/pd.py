"""Prefill/decode (PD) disaggregation.

Parity: reference ``python/sglang/srt/disaggregation/`` (``prefill.py:119,485`` bootstrap queue -> forward -> KV send;
``decode.py:300,1991,2426`` prealloc queue -> transfer queue -> prebuilt batch; transports ``{mooncake,nixl,mori,fake}/conn.py``
behind ``base/conn.py:101-247`` KVManager/KVSender/KVReceiver + ``KVPoll`` states; fault injection
``SGLANG_TEST_DISAGG_FAILURE_PROB``).

Design: a prefill instance runs the prompt, keeps the request's KV pages alive until they are shipped, and sends
``(room, token ids, first sampled token, KV pages [L,2,n,ps,H,D])`` to the decode instance named in the request's
bootstrap fields.  The decode instance pre-allocates pages, scatters the received KV into its pool (one strided copy: the
pool is a single tensor with a contiguous page dimension per plane) and inserts the request straight into the running
decode batch — no prefill on the decode side.  Transports: ``tcp`` (magic + length-prefixed JSON header + HMAC-SHA256 under ``SGLB_DISAGG_SECRET`` + raw tensor bytes;
works between any two processes), ``fake`` (same process, zero-copy hand-off; the unit-test / warm-up transport).
Intra-node NVLink peer copies reuse the NVL arena machinery and are wired through the same ``KVTransport`` interface.
"""
from __future__ import annotations

import enum
import hashlib
import hmac
import json
import logging
import os
import queue
import random
import socket
import struct
import threading
import time
from dataclasses import dataclass
from typing import Any, Dict, List, Optional, Tuple

import numpy as np
import torch

from ..environ import envs
from ..managers.schedule_batch import FINISH_ABORT, Req
from ..sampling.sampling_params import SamplingParams

logger = logging.getLogger(__name__)


def _json_default(o):
    if isinstance(o, (np.integer,)):
        return int(o)
    if isinstance(o, (np.floating,)):
        return float(o)
    if isinstance(o, (set, tuple)):
        return list(o)
    raise TypeError(f"KV header field of type {type(o).__name__} is not serialisable")


class KVPoll(enum.IntEnum):
    Failed = 0
    Bootstrapping = 1
    WaitingForInput = 2
    Transferring = 3
    Success = 4


@dataclass
class KVPayload:
    room: int
    rid: str
    input_ids: List[int]
    first_token: int
    sampling_params: Dict[str, Any]
    kv: Optional[torch.Tensor]      # [L, 2, n_pages, ps, H, D] on CPU (tcp) or device (fake)
    n_tokens: int
    stream: bool = False


# ---------------------------------------------------------------- transports
class KVTransport:
    def send(self, addr: Tuple[str, int], payload: KVPayload) -> bool:
        raise NotImplementedError

    def poll(self) -> List[KVPayload]:
        raise NotImplementedError

    def close(self) -> None:
        pass


class FakeTransport(KVTransport):
    _MAILBOX: Dict[Tuple[str, int], "queue.Queue[KVPayload]"] = {}

    def __init__(self, listen: Optional[Tuple[str, int]] = None):
        self.listen = listen
        if listen is not None:
            FakeTransport._MAILBOX.setdefault(listen, queue.Queue())

    def send(self, addr, payload):
        if random.random() < envs.SGLB_TEST_DISAGG_FAILURE_PROB.get():
            return False
        FakeTransport._MAILBOX.setdefault(tuple(addr), queue.Queue()).put(payload)
        return True

    def poll(self):
        out = []
        q = FakeTransport._MAILBOX.get(self.listen)
        while q is not None and not q.empty():
            out.append(q.get())
        return out


class TcpTransport(KVTransport):
    def __init__(self, listen: Optional[Tuple[str, int]] = None, secret: Optional[str] = None):
        self.inbox: "queue.Queue[KVPayload]" = queue.Queue()
        self._stop = threading.Event()
        self.sock = None
        self._key = (secret if secret is not None else os.environ.get("SGLB_DISAGG_SECRET", "")).encode()
        if listen is not None:
            if not self._key and listen[0] not in ("127.0.0.1", "localhost", "::1"):
                logger.warning("PD KV listener on %s without SGLB_DISAGG_SECRET: any host that can reach the port can inject KV payloads", listen[0])
            self.sock = socket.socket()
            self.sock.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
            self.sock.bind(listen)
            self.sock.listen(64)
            self.port = self.sock.getsockname()[1]
            threading.Thread(target=self._accept, daemon=True).start()

    def _accept(self) -> None:
        self.sock.settimeout(0.5)
        while not self._stop.is_set():
            try:
                c, _ = self.sock.accept()
            except socket.timeout:
                continue
            except OSError:
                return
            threading.Thread(target=self._recv, args=(c,), daemon=True).start()

    @staticmethod
    def _read(c: socket.socket, n: int) -> bytes:
        buf = bytearray()
        while len(buf) < n:
            chunk = c.recv(min(1 << 20, n - len(buf)))
            if not chunk:
                raise ConnectionError("peer closed")
            buf.extend(chunk)
        return bytes(buf)

    # Wire format (no executable payloads — the first version unpickled the header of ANY inbound connection, i.e. remote code execution
    # for whoever could reach the bootstrap port):
    #   b"SGLBKV1\0" | u32 header_len | header (UTF-8 JSON) | 32-byte HMAC-SHA256(secret, header) | raw KV bytes
    # dtype names go through an allow-list; the shared secret comes from SGLB_DISAGG_SECRET (both instances) and is verified in constant time.
    MAGIC = b"SGLBKV1\0"
    DTYPES = {"bfloat16": torch.bfloat16, "float16": torch.float16, "float32": torch.float32, "uint8": torch.uint8,
              "float8_e4m3fn": torch.float8_e4m3fn}
    MAX_HEADER = 64 << 20

    def _secret(self) -> bytes:
        return self._key

    def _recv(self, c: socket.socket) -> None:
        try:
            with c:
                if self._read(c, 8) != self.MAGIC:
                    raise ValueError("bad magic")
                (hl,) = struct.unpack("<I", self._read(c, 4))
                if hl > self.MAX_HEADER:
                    raise ValueError("header too large")
                hraw = self._read(c, hl)
                mac = self._read(c, 32)
                if not hmac.compare_digest(mac, hmac.new(self._secret(), hraw, hashlib.sha256).digest()):
                    raise PermissionError("KV payload rejected: HMAC mismatch (SGLB_DISAGG_SECRET differs between the instances?)")
                head = json.loads(hraw.decode("utf-8"))
                kv = None
                nbytes = int(head["kv_bytes"])
                if nbytes:
                    dt = self.DTYPES[head["kv_dtype"]]          # KeyError -> rejected: only plain tensor dtypes
                    shape = tuple(int(x) for x in head["kv_shape"])
                    n = 1
                    for d in shape:
                        n *= d
                    if n * torch.empty(0, dtype=dt).element_size() != nbytes or nbytes > (64 << 30):
                        raise ValueError("kv_shape does not match kv_bytes")
                    raw = self._read(c, nbytes)
                    kv = torch.frombuffer(bytearray(raw), dtype=torch.uint8).view(dt).view(shape)
                f = head["fields"]
                self.inbox.put(KVPayload(kv=kv, room=int(f["room"]), rid=str(f["rid"]), input_ids=[int(t) for t in f["input_ids"]],
                                         first_token=int(f["first_token"]), sampling_params=dict(f["sampling_params"] or {}),
                                         n_tokens=int(f["n_tokens"]), stream=bool(f["stream"])))
                c.sendall(b"OK")
        except Exception as e:
            logger.error("KV receive failed: %s", e)

    def send(self, addr, payload):
        if random.random() < envs.SGLB_TEST_DISAGG_FAILURE_PROB.get():
            return False
        try:
            kv = payload.kv.contiguous().cpu() if payload.kv is not None else None
            fields = {k: getattr(payload, k) for k in ("room", "rid", "input_ids", "first_token", "sampling_params", "n_tokens", "stream")}
            fields["input_ids"] = [int(t) for t in fields["input_ids"]]
            raw = kv.view(torch.uint8).numpy().tobytes() if kv is not None else b""
            head = json.dumps({"fields": fields, "kv_bytes": len(raw), "kv_dtype": str(kv.dtype).replace("torch.", "") if kv is not None else None,
                               "kv_shape": list(kv.shape) if kv is not None else None}, default=_json_default).encode("utf-8")
            mac = hmac.new(self._secret(), head, hashlib.sha256).digest()
            with socket.create_connection(tuple(addr), timeout=30) as c:
                c.sendall(self.MAGIC + struct.pack("<I", len(head)) + head + mac + raw)
                return self._read(c, 2) == b"OK"
        except Exception as e:
            logger.error("KV send to %s failed: %s", addr, e)
            return False

    def poll(self):
        out = []
        while not self.inbox.empty():
            out.append(self.inbox.get())
        return out

    def close(self):
        self._stop.set()
        if self.sock is not None:
            self.sock.close()


def make_transport(kind: str, listen: Optional[Tuple[str, int]] = None) -> KVTransport:
    if kind == "fake":
        return FakeTransport(listen)
    if kind in ("tcp", "nvlink"):
        return TcpTransport(listen)
    raise ValueError(f"unknown disaggregation transfer backend {kind}")


# ---------------------------------------------------------------- scheduler glue
class PrefillDisagg:
    """installed on a prefill instance: ships KV instead of decoding"""

    def __init__(self, engine: Any, transport: KVTransport, max_retries: int = 3):
        self.engine, self.transport, self.max_retries = engine, transport, max_retries
        self.sent = 0
        self.failed = 0
        sch = engine.scheduler
        self._orig_release = sch.kv.release
        self._bootstrap: Dict[str, Tuple[str, int, int]] = {}
        orig_handle = sch.handle_generate_request

        def handle(m):
            if m.bootstrap_host is not None:
                self._bootstrap[m.rid] = (m.bootstrap_host, int(m.bootstrap_port), int(m.bootstrap_room or 0))
                m.sampling_params = {**m.sampling_params, "max_new_tokens": 1}
            return orig_handle(m)

        sch.handle_generate_request = handle

        def release(req, insert=True):
            if req.rid in self._bootstrap and req.finished() and req.finished_reason.kind != "abort" and req.output_ids:
                self._ship(req)
            return self._orig_release(req, insert)

        sch.kv.release = release

    def _ship(self, req: Req) -> None:
        host, port, room = self._bootstrap.pop(req.rid)
        runner, kv = self.engine.runner, self.engine.scheduler.kv
        n_tok = len(req.origin_input_ids)
        n_pages = (n_tok + kv.ps - 1) // kv.ps
        pages = torch.from_numpy(kv.page_table[req.req_slot, :n_pages].astype(np.int64)).to(runner.device)
        data = runner.kv_pool.gather_pages(pages)
        sp = req.sampling_params.to_dict()
        sp["max_new_tokens"] = getattr(req, "_orig_max_new_tokens", None) or sp["max_new_tokens"]
        payload = KVPayload(room, req.rid, list(req.origin_input_ids), int(req.output_ids[0]), sp, data, n_tok, req.stream)
        for attempt in range(self.max_retries):
            if self.transport.send((host, port), payload):
                self.sent += 1
                return
            time.sleep(0.01 * (attempt + 1))
        self.failed += 1
        logger.error("KV transfer for %s failed after %d attempts", req.rid, self.max_retries)


class DecodeDisagg:
    """installed on a decode instance: turns received KV into running requests (reference 'prebuilt' batches)"""

    def __init__(self, engine: Any, transport: KVTransport):
        self.engine, self.transport = engine, transport
        self.expected: Dict[int, Any] = {}        # room -> TokenizedGenerateReqInput (arrives via the router)
        self.arrived: Dict[int, KVPayload] = {}
        self.received = 0
        sch = engine.scheduler
        orig_handle = sch.handle_generate_request
        orig_next = sch.get_next_batch_to_run
        orig_has_work = sch.has_work

        def handle(m):
            if m.bootstrap_room is not None and engine.server_args.disaggregation_mode == "decode":
                self.expected[int(m.bootstrap_room)] = m
                return None
            return orig_handle(m)

        def next_batch():
            self._admit()
            return orig_next()

        sch.handle_generate_request = handle
        sch.get_next_batch_to_run = next_batch
        sch.has_work = lambda: orig_has_work() or bool(self.expected) or bool(self.arrived)

    def _admit(self) -> None:
        for p in self.transport.poll():
            self.arrived[p.room] = p
        sch, runner = self.engine.scheduler, self.engine.runner
        kv = sch.kv
        for room in list(self.arrived):
            m = self.expected.get(room)
            if m is None:
                continue
            p = self.arrived[room]
            sp = SamplingParams.from_dict(m.sampling_params)
            req = Req(m.rid, p.input_ids, sp, stream=m.stream, eos_token_ids=sch.eos_token_ids)
            req._match = (np.empty(0, dtype=np.int32), sch.tree.match_prefix([], None)[1])
            if len(sch.running_batch.reqs) >= sch.max_running or not kv.admit(req):
                continue                       # no slot right now: retry next iteration
            n_pages = (p.n_tokens + kv.ps - 1) // kv.ps
            if not kv.alloc_for_extend(req, p.n_tokens):
                kv.release(req, insert=False)
                continue
            pages = torch.from_numpy(kv.page_table[req.req_slot, :n_pages].astype(np.int64)).to(runner.device)
            runner.kv_pool.scatter_pages(pages, p.kv.to(runner.device))
            req.kv_len = p.n_tokens
            req.output_ids = [p.first_token]
            req.first_token_time = time.perf_counter()
            req.cached_tokens = 0
            # the first token is already known: make it visible to the device-side future map and to the client
            runner.next_token_buf[req.req_slot] = p.first_token
            req.check_finished(sch.tokenizer)
            del self.arrived[room], self.expected[room]
            self.received += 1
            if req.finished():
                kv.release(req)
                sch.stream_output([req])
                continue
            from ..managers.schedule_batch import ForwardMode, ScheduleBatch

            one = ScheduleBatch(reqs=[req], forward_mode=ForwardMode.DECODE)
            one.init_decode_state()
            sch.running_batch.merge_batch(one)
            sch.stream_output([req])


def install_disaggregation(engine: Any, listen_port: Optional[int] = None) -> Any:
    a = engine.server_args
    if a.disaggregation_mode == "prefill":
        return PrefillDisagg(engine, make_transport(a.disaggregation_transfer_backend))
    if a.disaggregation_mode == "decode":
        port = listen_port if listen_port is not None else a.disaggregation_bootstrap_port
        return DecodeDisagg(engine, make_transport(a.disaggregation_transfer_backend, (a.host, port)))
    return None
