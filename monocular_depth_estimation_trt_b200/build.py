"""Compile libmde_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

The library links cudart statically and resolves the one driver symbol it needs
(cuTensorMapEncodeTiled) at run time, so it loads on a box without a GPU; every compute
entry point then fails with MDE_ERR_CUDA -- there is no CPU fallback.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.path.join(HERE, "libmde_b200.so")
SOURCES = ["kernels.cu", "engine.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17", *os.environ.get("MDE_NVCC_EXTRA", "").split(),
    "-Xcompiler", "-fPIC",
    "--cudart", "static",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("[MDET] nvcc not found; cannot build libmde_b200.so")


def _stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    built = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h"))]
    deps.append(os.path.join(os.path.dirname(HERE), "include", "mde_b200.h"))
    return any(os.path.getmtime(p) > built for p in deps if os.path.isfile(p))


def build(force: bool = False, verbose: bool = False) -> str:
    """Build (if stale) and return the path of the shared library.  Safe to call from several processes at once (one rank
    per GPU under torchrun): an exclusive file lock serialises the builders, staleness is re-checked under the lock, objects
    and the library are written under private names and moved into place atomically."""
    if not force and not _stale():
        return LIB_PATH
    import fcntl
    with open(os.path.join(HERE, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not _stale():
                return LIB_PATH          # another process built it while this one waited
            return _build_locked(verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(verbose: bool) -> str:
    nvcc = _nvcc()
    objs = []
    procs = []
    tag = f".{os.getpid()}"
    for src in SOURCES:
        obj = os.path.join(CSRC, src.replace(".cu", tag + ".o"))
        cmd = [nvcc, *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, p in procs:
        out, _ = p.communicate()
        if verbose and out:
            print(out)
        if p.returncode != 0:
            for o in objs:
                if os.path.exists(o):
                    os.remove(o)
            raise RuntimeError(f"[MDET] nvcc failed on {src}:\n{out}")
    tmp_lib = LIB_PATH + tag
    link = [nvcc, "-shared", "--cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a",
            "-o", tmp_lib, *objs, "-ldl", "-lpthread", "-lrt"]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    for o in objs:
        if os.path.exists(o):
            os.remove(o)
    if r.returncode != 0:
        raise RuntimeError(f"[MDET] link failed:\n{r.stdout}")
    os.replace(tmp_lib, LIB_PATH)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
