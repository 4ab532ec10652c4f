"""Runtime (JIT) build + cache for small native ops that are not part of the ahead-of-time ``_C.so``.

Parity: reference ``python/sglang/kernels/jit/utils/compile/{cache,loader,toolchain,spec}.py`` (per-GPU JIT: nvcc + ninja ->
cached ``.so`` keyed on the source hash, ``arch.py:73-95`` picks ``sm_XXa``).  One target here (sm_100a), so: content-addressed
cache dir ``<SGLB_JIT_CACHE or ~/.cache/sglang_b200/jit>/<name>-<sha1(sources+flags+torch ABI)>/``, an ``fcntl`` lock so that
concurrent ranks build once, nvcc for ``.cu`` / g++ for ``.cpp``, loaded through ``torch.ops.load_library``.

    from sglang_b200 import jit
    jit.load("sglb_jit_weak_ref", [jit.CSRC_JIT / "weak_ref.cpp"])
    alias = torch.ops.sglb_jit_weak_ref.weak_ref_tensor(t)

User kernels: ``jit.load_inline(name, cuda_source, functions=...)`` wraps a CUDA string the same way (custom logit processors,
experiments) without touching the in-tree build.
"""
from __future__ import annotations

import fcntl
import hashlib
import os
import subprocess
from pathlib import Path
from typing import Dict, Sequence

ROOT = Path(__file__).resolve().parent
CSRC_JIT = ROOT / "csrc" / "jit"
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]
_LOADED: Dict[str, Path] = {}


def cache_dir() -> Path:
    d = Path(os.environ.get("SGLB_JIT_CACHE", Path.home() / ".cache" / "sglang_b200" / "jit"))
    d.mkdir(parents=True, exist_ok=True)
    return d


def _key(name: str, sources: Sequence[Path], flags: Sequence[str]) -> str:
    import torch

    h = hashlib.sha1()
    h.update(name.encode())
    for s in sources:
        h.update(Path(s).read_bytes())
    h.update(" ".join(flags).encode())
    h.update(torch.__version__.encode())
    h.update(str(torch._C._GLIBCXX_USE_CXX11_ABI).encode())
    return h.hexdigest()[:16]


def build(name: str, sources: Sequence[Path], extra_cflags: Sequence[str] = (), extra_cuda_cflags: Sequence[str] = (),
          extra_ldflags: Sequence[str] = ()) -> Path:
    """compile (if not cached) and return the path of the shared object"""
    import torch
    from torch.utils import cpp_extension as ce

    sources = [Path(s) for s in sources]
    has_cu = any(s.suffix == ".cu" for s in sources)
    inc = [f"-I{p}" for p in ce.include_paths(device_type="cuda" if has_cu else "cpu")] + [f"-I{ROOT / 'csrc'}"]
    common = ["-O2", "-std=c++17", "-fPIC", f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}", "-DTORCH_API_INCLUDE_EXTENSION_H"]
    cflags = common + list(extra_cflags)
    cuflags = ARCH_FLAGS + ["-O3", "-std=c++17", "-lineinfo", "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC",
                            f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}"] + list(extra_cuda_cflags)
    key = _key(name, sources, cflags + cuflags + list(extra_ldflags))
    out_dir = cache_dir() / f"{name}-{key}"
    so = out_dir / f"{name}.so"
    if so.exists():
        return so
    out_dir.mkdir(parents=True, exist_ok=True)
    with open(out_dir / ".lock", "w") as lk:
        fcntl.flock(lk, fcntl.LOCK_EX)       # one builder per cache entry; the others wait and find the .so
        try:
            if so.exists():
                return so
            objs = []
            for s in sources:
                obj = out_dir / (s.stem + ".o")
                cmd = ([NVCC] + cuflags + inc + ["-c", str(s), "-o", str(obj)]) if s.suffix == ".cu" else \
                      (["g++"] + cflags + inc + ["-c", str(s), "-o", str(obj)])
                p = subprocess.run(cmd, capture_output=True, text=True)
                (out_dir / (s.stem + ".log")).write_text("$ " + " ".join(cmd) + "\n" + p.stdout + p.stderr)
                if p.returncode != 0:
                    raise RuntimeError(f"jit build of {name} failed:\n{p.stderr[-4000:]}")
                objs.append(str(obj))
            tlib = os.path.join(os.path.dirname(torch.__file__), "lib")
            link = ["g++", "-shared", "-o", str(so) + ".tmp"] + objs + [f"-L{tlib}", "-lc10", "-ltorch", "-ltorch_cpu", f"-Wl,-rpath,{tlib}"]
            if has_cu:
                link += ["-lc10_cuda", "-ltorch_cuda", "-L/usr/local/cuda/lib64", "-lcudart"]
            link += list(extra_ldflags)
            p = subprocess.run(link, capture_output=True, text=True)
            if p.returncode != 0:
                raise RuntimeError(f"jit link of {name} failed:\n{p.stderr[-4000:]}")
            os.replace(str(so) + ".tmp", so)      # atomic publish
        finally:
            fcntl.flock(lk, fcntl.LOCK_UN)
    return so


def load(name: str, sources: Sequence[Path], **kw) -> Path:
    """build (cached) + ``torch.ops.load_library``; idempotent per process"""
    import torch

    if name in _LOADED:
        return _LOADED[name]
    so = build(name, sources, **kw)
    torch.ops.load_library(str(so))
    _LOADED[name] = so
    return so


def load_inline(name: str, source: str, is_cuda: bool = True, **kw) -> Path:
    """same as ``load`` for a source string (must contain its own TORCH_LIBRARY block registering ops under ``name``)"""
    d = cache_dir() / "inline-src"
    d.mkdir(parents=True, exist_ok=True)
    f = d / f"{name}-{hashlib.sha1(source.encode()).hexdigest()[:12]}{'.cu' if is_cuda else '.cpp'}"
    if not f.exists():
        f.write_text(source)
    return load(name, [f], **kw)


def weak_ref_tensor(t):
    """non-owning alias (see csrc/jit/weak_ref.cpp)"""
    import torch

    load("sglb_jit_weak_ref", [CSRC_JIT / "weak_ref.cpp"])
    return torch.ops.sglb_jit_weak_ref.weak_ref_tensor(t)
