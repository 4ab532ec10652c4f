"""L3 storage backends of the hierarchical KV cache: pages keyed by their chained content hash.

Parity: reference ``python/sglang/srt/mem_cache/hicache_storage.py:150`` (``HiCacheStorage`` ABC, ``HiCacheFile``) and the backend zoo under
``mem_cache/storage/`` (mooncake / hf3fs / nixl / aibrix / eic / lmcache / shm / sim).  The external engines are not available offline; what is
implemented here covers the same *deployment shapes*:

* ``file``    one file per page under a directory (a shared file system gives cross-instance reuse);
* ``shm``     raw pages in ``/dev/shm`` through ``numpy.memmap`` — node-local, shared by every engine process on the host, no serialisation;
* ``memory``  in-process dictionary with an LRU capacity (tests / simulation, the reference's ``sim`` backend);
* ``remote``  a TCP page server (``python -m sglang_b200.mem_cache.hicache_storage --port P``): any number of engines on any hosts share one
  pool — the stand-in for the RDMA object stores; batched ``exists / get / put`` keep round trips per request constant.

Out-of-tree backends register with ``register_storage_backend(name, factory)`` (e.g. from a plugin).
"""
from __future__ import annotations

import json
import os
import socket
import socketserver
import struct
import threading
from collections import OrderedDict
from typing import Any, Callable, Dict, List, Optional, Sequence

import numpy as np
import torch

_DT = {"torch.bfloat16": torch.bfloat16, "torch.float16": torch.float16, "torch.float32": torch.float32, "torch.float8_e4m3fn": torch.float8_e4m3fn,
       "torch.uint8": torch.uint8}


def _to_bytes(t: torch.Tensor) -> bytes:
    t = t.contiguous().cpu()
    hdr = json.dumps({"dtype": str(t.dtype), "shape": list(t.shape)}).encode()
    return struct.pack("<I", len(hdr)) + hdr + t.view(torch.uint8).numpy().tobytes()


def _from_bytes(b: bytes) -> torch.Tensor:
    n = struct.unpack("<I", b[:4])[0]
    h = json.loads(b[4:4 + n])
    raw = torch.frombuffer(bytearray(b[4 + n:]), dtype=torch.uint8)
    return raw.view(_DT[h["dtype"]]).reshape(h["shape"])


class HiCacheStorage:
    """page store interface; the batch methods have loop defaults so a backend only needs the three scalar ones"""

    def exists(self, h: int) -> bool:
        raise NotImplementedError

    def put(self, h: int, page: torch.Tensor) -> None:
        raise NotImplementedError

    def get(self, h: int) -> Optional[torch.Tensor]:
        raise NotImplementedError

    def batch_exists(self, hs: Sequence[int]) -> List[bool]:
        return [self.exists(h) for h in hs]

    def batch_get(self, hs: Sequence[int]) -> List[Optional[torch.Tensor]]:
        return [self.get(h) for h in hs]

    def batch_put(self, hs: Sequence[int], pages: Sequence[torch.Tensor]) -> None:
        for h, p in zip(hs, pages):
            self.put(h, p)

    def stats(self) -> Dict[str, Any]:
        return {}

    def close(self) -> None:
        pass


class FileStorage(HiCacheStorage):
    """one file per page hash (reference HiCacheFile)"""

    def __init__(self, root: str):
        self.root = root
        os.makedirs(root, exist_ok=True)

    def _p(self, h: int) -> str:
        return os.path.join(self.root, f"{h:016x}.kv")

    def exists(self, h: int) -> bool:
        return os.path.exists(self._p(h))

    def put(self, h: int, page: torch.Tensor) -> None:
        tmp = self._p(h) + f".tmp{os.getpid()}"
        with open(tmp, "wb") as f:
            f.write(_to_bytes(page))
        os.replace(tmp, self._p(h))

    def get(self, h: int) -> Optional[torch.Tensor]:
        try:
            with open(self._p(h), "rb") as f:
                return _from_bytes(f.read())
        except FileNotFoundError:
            return None

    def stats(self):
        return {"backend": "file", "pages": len([f for f in os.listdir(self.root) if f.endswith(".kv")])}


class ShmStorage(FileStorage):
    """the same layout under /dev/shm: page bytes never touch a disk and are visible to every process of the host"""

    def __init__(self, name: str = "sglb_hicache"):
        super().__init__(os.path.join("/dev/shm" if os.path.isdir("/dev/shm") else "/tmp", name))

    def get(self, h: int) -> Optional[torch.Tensor]:
        p = self._p(h)
        if not os.path.exists(p):
            return None
        mm = np.memmap(p, dtype=np.uint8, mode="r")
        n = int(np.frombuffer(mm[:4].tobytes(), dtype="<u4")[0])
        hdr = json.loads(bytes(mm[4:4 + n]))
        return torch.from_numpy(np.array(mm[4 + n:])).view(_DT[hdr["dtype"]]).reshape(hdr["shape"])

    def stats(self):
        return {**super().stats(), "backend": "shm"}


class MemoryStorage(HiCacheStorage):
    def __init__(self, capacity_pages: int = 1 << 16):
        self.cap = capacity_pages
        self.d: "OrderedDict[int, bytes]" = OrderedDict()
        self.lock = threading.Lock()
        self.hits = self.misses = 0

    def exists(self, h):
        with self.lock:
            return h in self.d

    def put(self, h, page):
        b = _to_bytes(page)
        with self.lock:
            self.d[h] = b
            self.d.move_to_end(h)
            while len(self.d) > self.cap:
                self.d.popitem(last=False)

    def get(self, h):
        with self.lock:
            b = self.d.get(h)
            if b is None:
                self.misses += 1
                return None
            self.hits += 1
            self.d.move_to_end(h)
        return _from_bytes(b)

    def stats(self):
        return {"backend": "memory", "pages": len(self.d), "hits": self.hits, "misses": self.misses}


# ------------------------------------------------------------------------------------------------ remote page server
# wire: request = u8 op ('E' exists / 'G' get / 'P' put / 'S' stats), u32 n, n x u64 hashes, for 'P' n x (u32 len, bytes);
#       reply   = 'E': n x u8;  'G': n x (u32 len or 0, bytes);  'P': u8 ok;  'S': u32 len, json
def _recv(sock: socket.socket, n: int) -> bytes:
    buf = bytearray()
    while len(buf) < n:
        chunk = sock.recv(min(1 << 20, n - len(buf)))
        if not chunk:
            raise ConnectionError("page server connection closed")
        buf += chunk
    return bytes(buf)


class _Handler(socketserver.BaseRequestHandler):
    def handle(self):
        store: MemoryStorage = self.server.store
        s = self.request
        try:
            while True:
                head = s.recv(1)
                if not head:
                    return
                op = head.decode()
                n = struct.unpack("<I", _recv(s, 4))[0]
                hs = list(struct.unpack(f"<{n}Q", _recv(s, 8 * n))) if n else []
                if op == "E":
                    s.sendall(bytes(int(store.exists(h)) for h in hs))
                elif op == "G":
                    out = bytearray()
                    for h in hs:
                        with store.lock:
                            b = store.d.get(h)
                            if b is not None:
                                store.d.move_to_end(h)
                        out += struct.pack("<I", len(b) if b else 0) + (b or b"")
                    s.sendall(bytes(out))
                elif op == "P":
                    for h in hs:
                        ln = struct.unpack("<I", _recv(s, 4))[0]
                        b = _recv(s, ln)
                        with store.lock:
                            store.d[h] = b
                            store.d.move_to_end(h)
                            while len(store.d) > store.cap:
                                store.d.popitem(last=False)
                    s.sendall(b"\x01")
                elif op == "S":
                    j = json.dumps(store.stats()).encode()
                    s.sendall(struct.pack("<I", len(j)) + j)
                else:
                    return
        except (ConnectionError, OSError):
            return


class PageServer(socketserver.ThreadingTCPServer):
    allow_reuse_address = True
    daemon_threads = True

    def __init__(self, host: str = "127.0.0.1", port: int = 0, capacity_pages: int = 1 << 18):
        super().__init__((host, port), _Handler)
        self.store = MemoryStorage(capacity_pages)

    @property
    def port(self) -> int:
        return self.server_address[1]

    def start(self) -> "PageServer":
        threading.Thread(target=self.serve_forever, daemon=True).start()
        return self


class RemoteStorage(HiCacheStorage):
    def __init__(self, host: str, port: int, timeout: float = 10.0):
        self.addr, self.timeout = (host, int(port)), timeout
        self.sock: Optional[socket.socket] = None
        self.lock = threading.Lock()

    def _s(self) -> socket.socket:
        if self.sock is None:
            self.sock = socket.create_connection(self.addr, timeout=self.timeout)
            self.sock.setsockopt(socket.IPPROTO_TCP, socket.TCP_NODELAY, 1)
        return self.sock

    def _req(self, op: str, hs: Sequence[int], payload: bytes = b"") -> socket.socket:
        s = self._s()
        s.sendall(op.encode() + struct.pack("<I", len(hs)) + struct.pack(f"<{len(hs)}Q", *[h & (2 ** 64 - 1) for h in hs]) + payload)
        return s

    def batch_exists(self, hs):
        with self.lock:
            s = self._req("E", hs)
            return [bool(b) for b in _recv(s, len(hs))]

    def batch_get(self, hs):
        with self.lock:
            s = self._req("G", hs)
            out = []
            for _ in hs:
                ln = struct.unpack("<I", _recv(s, 4))[0]
                out.append(_from_bytes(_recv(s, ln)) if ln else None)
            return out

    def batch_put(self, hs, pages):
        blobs = [_to_bytes(p) for p in pages]
        with self.lock:
            s = self._req("P", hs, b"".join(struct.pack("<I", len(b)) + b for b in blobs))
            _recv(s, 1)

    def exists(self, h):
        return self.batch_exists([h])[0]

    def get(self, h):
        return self.batch_get([h])[0]

    def put(self, h, page):
        self.batch_put([h], [page])

    def stats(self):
        with self.lock:
            s = self._req("S", [])
            return {**json.loads(_recv(s, struct.unpack("<I", _recv(s, 4))[0])), "backend": "remote"}

    def close(self):
        if self.sock is not None:
            self.sock.close()
            self.sock = None


# ------------------------------------------------------------------------------------------------ registry
_BACKENDS: Dict[str, Callable[[Dict[str, Any]], HiCacheStorage]] = {
    "file": lambda c: FileStorage(c.get("dir") or "/tmp/sglb_hicache"),
    "shm": lambda c: ShmStorage(c.get("name", "sglb_hicache")),
    "memory": lambda c: MemoryStorage(int(c.get("capacity_pages", 1 << 16))),
    "sim": lambda c: MemoryStorage(int(c.get("capacity_pages", 1 << 16))),
    "remote": lambda c: RemoteStorage(c.get("host", "127.0.0.1"), int(c["port"])),
}


def register_storage_backend(name: str, factory: Callable[[Dict[str, Any]], HiCacheStorage]) -> None:
    _BACKENDS[name] = factory


def create_storage(name: Optional[str], config: Optional[Dict[str, Any]] = None) -> Optional[HiCacheStorage]:
    if not name:
        return None
    if name.startswith("remote://"):
        host, port = name[len("remote://"):].rsplit(":", 1)
        return RemoteStorage(host, int(port))
    if name not in _BACKENDS:
        raise ValueError(f"unknown hicache storage backend {name!r}; available: {sorted(_BACKENDS)}")
    return _BACKENDS[name](config or {})


if __name__ == "__main__":
    import argparse
    import time

    ap = argparse.ArgumentParser("sglang_b200 hicache page server")
    ap.add_argument("--host", default="0.0.0.0")
    ap.add_argument("--port", type=int, default=31000)
    ap.add_argument("--capacity-pages", type=int, default=1 << 18)
    ns = ap.parse_args()
    srv = PageServer(ns.host, ns.port, ns.capacity_pages).start()
    print(f"hicache page server on {ns.host}:{srv.port}", flush=True)
    while True:
        time.sleep(3600)
