"""In-tree native build: nvcc (sm_100a) for the kernels, g++ for bindings and the C++ runtime.

Produces, next to this file:
  _C.so        torch.ops.sglb.*   (CUDA kernels + torch bindings)
  sglb_rt*.so  pybind11 module    (C++ host runtime: radix tree, page allocator, n-gram corpus, hashing, router)

Replaces the reference's two build systems (CMake AOT wheel ``python/sglang/kernels/aot/CMakeLists.txt`` and the
tvm-ffi/ninja JIT cache ``kernels/jit/utils/compile/*``): there is exactly one target arch, so an explicit,
incremental (mtime + flags hash) ahead-of-time build is enough, and the .so files travel with the source tree.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
import sysconfig
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path
from typing import List

ROOT = Path(__file__).resolve().parent
CSRC = ROOT / "csrc"
BUILD = ROOT.parent / "build" / "sglb"
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
CUDA_HOME = os.environ.get("CUDA_HOME", "/usr/local/cuda")
ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ARCH_FLAGS + ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
                           "--expt-relaxed-constexpr", "-Xptxas", "-v"]
CXX_FLAGS = ["-O2", "-std=c++17", "-fPIC", "-Wall", "-Wno-unused-function"]


def _run(cmd: List[str], log: Path) -> None:
    p = subprocess.run(cmd, capture_output=True, text=True)
    log.write_text("$ " + " ".join(cmd) + "\n" + p.stdout + p.stderr)
    if p.returncode != 0:
        sys.stderr.write(p.stdout[-4000:] + p.stderr[-8000:])
        raise RuntimeError(f"build failed: {' '.join(cmd[:3])} ... (see {log})")


def _stale(out: Path, deps: List[Path], flags: List[str]) -> bool:
    stamp = out.with_suffix(out.suffix + ".flags")
    h = hashlib.sha1(" ".join(flags).encode()).hexdigest()
    if not out.exists() or not stamp.exists() or stamp.read_text() != h:
        return True
    t = out.stat().st_mtime
    return any(d.stat().st_mtime > t for d in deps)


def _mark(out: Path, flags: List[str]) -> None:
    out.with_suffix(out.suffix + ".flags").write_text(hashlib.sha1(" ".join(flags).encode()).hexdigest())


def build_kernels(verbose: bool = False) -> Path:
    import torch
    from torch.utils import cpp_extension as ce

    BUILD.mkdir(parents=True, exist_ok=True)
    headers = list(CSRC.glob("kernels/*.h")) + list(CSRC.glob("kernels/*.cuh")) + list(CSRC.glob("comm/*.cuh")) + \
        list(CSRC.glob("comm/*.h"))
    cu_srcs = sorted(list(CSRC.glob("kernels/*.cu")) + list(CSRC.glob("comm/*.cu")))
    objs, jobs = [], []
    for src in cu_srcs:
        obj = BUILD / (src.stem + ".o")
        objs.append(obj)
        if _stale(obj, [src] + headers, NVCC_FLAGS):
            jobs.append((src, obj, [NVCC] + NVCC_FLAGS + ["-I", str(CSRC), "-c", str(src), "-o", str(obj)], NVCC_FLAGS))
    # torch binding (host only, but needs CUDA + torch headers)
    tinc = [f"-I{p}" for p in ce.include_paths(device_type="cuda")] + [f"-I{CUDA_HOME}/include"]
    bflags = CXX_FLAGS + tinc + [f"-I{CSRC}", "-DTORCH_API_INCLUDE_EXTENSION_H", f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}"]
    for src in sorted(CSRC.glob("*.cpp")):
        obj = BUILD / (src.stem + ".o")
        objs.append(obj)
        if _stale(obj, [src] + headers, bflags):
            jobs.append((src, obj, ["g++"] + bflags + ["-c", str(src), "-o", str(obj)], bflags))

    def do(job):
        src, obj, cmd, flags = job
        if verbose:
            print("[sglb build]", src.name, flush=True)
        _run(cmd, BUILD / (obj.stem + ".log"))
        _mark(obj, flags)

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 4)) as ex:
        list(ex.map(do, jobs))

    so = ROOT / "_C.so"
    tlib = Path(torch.__file__).parent / "lib"
    lflags = ["-shared", f"-L{tlib}", f"-L{CUDA_HOME}/lib64", "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch",
              "-lcudart", f"-Wl,-rpath,{tlib}", f"-Wl,-rpath,{CUDA_HOME}/lib64"]
    if jobs or _stale(so, objs, lflags):
        _run(["g++", "-o", str(so)] + [str(o) for o in objs] + lflags, BUILD / "link_C.log")
        _mark(so, lflags)
    return so


def build_runtime(verbose: bool = False) -> Path:
    import pybind11

    BUILD.mkdir(parents=True, exist_ok=True)
    srcs = sorted((CSRC / "runtime").glob("*.cpp"))
    headers = list((CSRC / "runtime").glob("*.h"))
    ext = sysconfig.get_config_var("EXT_SUFFIX")
    so = ROOT / f"sglb_rt{ext}"
    flags = CXX_FLAGS + ["-O3", f"-I{pybind11.get_include()}", f"-I{sysconfig.get_paths()['include']}", "-fvisibility=hidden", "-pthread"]
    objs, jobs = [], []
    for src in srcs:
        obj = BUILD / ("rt_" + src.stem + ".o")
        objs.append(obj)
        if _stale(obj, [src] + headers, flags):
            jobs.append((src, obj, ["g++"] + flags + ["-c", str(src), "-o", str(obj)], flags))

    def do(job):
        src, obj, cmd, fl = job
        if verbose:
            print("[sglb build]", src.name, flush=True)
        _run(cmd, BUILD / (obj.stem + ".log"))
        _mark(obj, fl)

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 4)) as ex:
        list(ex.map(do, jobs))
    if jobs or _stale(so, objs, ["link"]):
        _run(["g++", "-shared", "-pthread", "-o", str(so)] + [str(o) for o in objs], BUILD / "link_rt.log")
        _mark(so, ["link"])
    return so


def build_all(verbose: bool = False) -> None:
    build_runtime(verbose)
    build_kernels(verbose)


if __name__ == "__main__":
    build_all(verbose=True)
    print("built:", ROOT / "_C.so")
