"""In-tree build of the dpb200 CUDA library (sm_100a only).

``build()`` compiles every ``csrc/*.cu`` with nvcc into ``lib/libdpb200.so`` (the C ABI declared in
``include/dpb200.h``) and ``csrc/deepmd_gpu_shim.cc`` into ``lib/libdeepmd_op_cuda.so`` (the C++
``deepmd::*_gpu`` symbols the reference's op layers link against).  Nothing is JIT-cached outside
the tree, so the built files travel with a snapshot of the repository.
"""
from __future__ import annotations

import os
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(HERE, "build")
LIB = os.path.join(LIBDIR, "libdpb200.so")
SHIM = os.path.join(LIBDIR, "libdeepmd_op_cuda.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the dpb200 CUDA library cannot be built")


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(OBJDIR, exist_ok=True)
    nvcc = _nvcc()
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(os.path.dirname(HERE), "include", "dpb200.h"))
    jobs = []
    objs = []
    for src in _sources():
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJDIR, src[:-3] + ".o")
        objs.append(o)
        if force or _stale(o, [s] + headers):
            jobs.append([nvcc] + NVCC_FLAGS + ["-c", s, "-o", o])

    def run(cmd):
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)

    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        list(ex.map(run, jobs))
    if force or jobs or _stale(LIB, objs):
        run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs + ["-cudart", "static"])
    build_shim(force, run)
    return LIB


def build_shim(force=False, run=subprocess.check_call) -> str | None:
    """lib/libdeepmd_op_cuda.so: the reference's C++ `deepmd::*_gpu` symbols on top of the C ABI.  It is
    compiled against the REFERENCE'S OWN headers (never copied), so it is only (re)built where a
    DeePMD-kit source tree is available ($DEEPMD_SOURCE_DIR or /root/reference)."""
    ref = os.environ.get("DEEPMD_SOURCE_DIR", "/root/reference")
    inc = os.path.join(ref, "source", "lib", "include")
    shim_src = os.path.join(CSRC, "deepmd_gpu_shim.cc")
    if not os.path.isdir(inc) or not os.path.exists(shim_src):
        return SHIM if os.path.exists(SHIM) else None
    if force or _stale(SHIM, [shim_src, LIB, os.path.join(os.path.dirname(HERE), "include", "dpb200.h")]):
        cxx = shutil.which("g++") or "g++"
        cuda_inc = os.path.join(os.path.dirname(os.path.dirname(_nvcc())), "include")
        run([cxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-DGOOGLE_CUDA=1", "-I", inc,
             "-I", os.path.join(os.path.dirname(HERE), "include"), "-I", cuda_inc, shim_src, "-o", SHIM,
             "-L", LIBDIR, "-ldpb200", "-L", os.path.join(os.path.dirname(cuda_inc), "lib64"), "-lcudart",
             "-Wl,-rpath,$ORIGIN"])
    return SHIM


if __name__ == "__main__":
    import sys

    print(build(force="--force" in sys.argv, verbose=True))
