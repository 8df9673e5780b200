"""TEST INFRASTRUCTURE ONLY -- ctypes loader for ``oracle/vrg_oracle.c``.

Only ``tests/``, ``__graft_entry__`` (build + smoke check) and ``bench.py``'s
CPU-baseline / ``--impl reference`` legs may import this.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "vrg_oracle.c")
LIB = os.path.join(HERE, "_build", "libvrg_oracle.so")

ERRORS = {-1: "initial valueMap may only hold labels 0, 3 and 4", -2: "empty seed set",
          -3: "seed has no boundary", -4: "more than 65536 distinct intensity levels", -5: "out of memory"}


def build(force: bool = False) -> str:
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        os.makedirs(os.path.dirname(LIB), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-fopenmp", "-shared", "-fPIC", "-o", LIB, SRC, "-lm"])
    return LIB


_lib = None


def _load():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        i64, p = ctypes.c_int64, ctypes.c_void_p
        _lib.vrg_oracle_run.restype = ctypes.c_int
        _lib.vrg_oracle_run.argtypes = [p, p, i64, i64, i64, ctypes.c_double, i64, i64, p, p, p, i64, p, p,
                                        p, i64, p, i64, p, ctypes.c_int]
        _lib.vrg_oracle_hash_labels.restype = ctypes.c_uint64
        _lib.vrg_oracle_hash_labels.argtypes = [p, i64, i64, ctypes.c_int]
    return _lib


def hash_labels(labels, base=0, nthreads=0) -> int:
    """Position-sensitive 64-bit hash of a uint8 label volume whose first voxel has global linear index ``base``: the
    number ``vrg_labels_hash`` (include/vrg_b200.h) computes on the device.  Slab hashes add up modulo 2^64."""
    lab = np.ascontiguousarray(labels, dtype=np.uint8)
    return int(_load().vrg_oracle_hash_labels(lab.ctypes.data, lab.size, int(base), int(nthreads)))


def hash_labels_numpy(labels, base=0) -> int:
    """The same hash in NumPy (small volumes; pins the C version in tests/test_oracle_golden.py)."""
    lab = np.ascontiguousarray(labels, dtype=np.uint8).ravel().astype(np.uint64)
    with np.errstate(over="ignore"):
        z = ((np.arange(lab.size, dtype=np.uint64) + np.uint64(base)) << np.uint64(3)) | lab
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
        return int(z.sum(dtype=np.uint64))


def vrg_oracle_c(data, value_map, H=2.25, max_segment_size=5000, iter_max=200, record_tables=False,
                 nthreads=0):
    """Same contract as ``oracle.vrg_oracle.vrg_oracle`` (dict result), at C speed."""
    lib = _load()
    data = np.ascontiguousarray(data, dtype=np.float64)
    labels = np.ascontiguousarray(value_map).astype(np.uint8)  # copy
    Z, Y, X = data.shape
    cap = int(iter_max) + 2
    trace = np.zeros((cap, 3), dtype=np.int64)
    quirk = np.zeros(4, dtype=np.int64)
    it = ctypes.c_int64(0)
    ex = ctypes.c_int64(0)
    nt = ctypes.c_int64(0)
    nl = ctypes.c_int64(0)
    levels = np.zeros(65536, dtype=np.float64)
    tables = None
    tcap = 0
    if record_tables:
        L = len(np.unique(data))
        tcap = cap
        tables = np.zeros((tcap, 2, L), dtype=np.float64)
    rc = lib.vrg_oracle_run(
        data.ctypes.data, labels.ctypes.data, Z, Y, X, float(H), int(min(max_segment_size, 2 ** 62)),
        int(iter_max), ctypes.addressof(it), ctypes.addressof(ex), trace.ctypes.data, cap,
        ctypes.addressof(nt), quirk.ctypes.data, tables.ctypes.data if tables is not None else None, tcap,
        levels.ctypes.data, levels.size, ctypes.addressof(nl), int(nthreads))
    if rc != 0:
        raise ValueError("oracle: " + ERRORS.get(rc, "error %d" % rc))
    res = {
        "labels": labels, "seg": labels <= 1, "iterations": int(it.value), "exit": int(ex.value),
        "trace": trace[: nt.value].copy(), "levels": levels[: nl.value].copy(),
        "quirk_potential": dict(zip(("add_to_inside", "remove_to_outside", "cancel_repromoted", "cancelled"),
                                    (int(q) for q in quirk))),
    }
    if record_tables:
        res["tables"] = [(tables[i, 0], tables[i, 1]) for i in range(min(int(it.value), tcap))]
    return res
