"""Compile the CUDA extension in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libvrg_b200.so")
SOURCES = ["vrg_b200.cu", "vrg_phantom.cu", "vrg_edt.cu", "vrg_mask.cu", "vrg_strict.cu"]
DEPS = SOURCES + ["vrg_kernels.cuh", "vrg_tail.cuh", "vrg_p2p.cuh", "vrg_parzen.cuh", "vrg_scratch.cuh", os.path.join("..", "..", "include", "vrg_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC"]


def find_nvcc():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    return None


FLAGS_STAMP = os.path.join(CSRC, ".build.flags")


def stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    if any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in DEPS):
        return True
    try:  # a library built with other flags is stale too (the stamp travels with the .so)
        with open(FLAGS_STAMP) as f:
            return f.read() != " ".join(NVCC_FLAGS)
    except OSError:
        return False  # prebuilt library without a stamp (e.g. built by hand): trust the time stamps


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile libvrg_b200.so.  Several ranks (torchrun, mp.spawn) may get here at once: one of them compiles, under a file
    lock, into a temporary file that is renamed into place, so nobody ever loads a half-written library."""
    if not force and not stale():
        return LIB
    nvcc = find_nvcc()
    if nvcc is None:
        raise RuntimeError("nvcc not found: cannot build %s" % LIB)
    import fcntl
    with open(os.path.join(CSRC, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not stale():  # another process built it while this one waited
                return LIB
            tmp = LIB + ".tmp.%d" % os.getpid()
            cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", tmp] + SOURCES
            try:
                subprocess.check_call(cmd, cwd=CSRC)
                os.replace(tmp, LIB)
            finally:
                if os.path.exists(tmp):
                    os.remove(tmp)
            with open(FLAGS_STAMP, "w") as f:
                f.write(" ".join(NVCC_FLAGS))
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
