"""TEST INFRASTRUCTURE ONLY -- ctypes loader for ``oracle/vrg_strict_oracle.c``: the reference's variational region
growing WITH its list order.  Only ``tests/`` may import this.

Parity status: PINNED.  ``tests/test_strict_oracle.py`` checks it against fixtures written by
``tests/golden/make_golden_strict.py`` from the UNMODIFIED reference on noisy inputs where the reference's result
differs from the order-free restatement (final valueMap including its stale band labels, the row order of
``segmented``, both band lists in list order, iteration count, the per-iteration trace, and the running Parzen sums
of every iteration to 1e-11 relative).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "vrg_strict_oracle.c")
LIB = os.path.join(HERE, "_build", "libvrg_strict_oracle.so")
EXIT_RUNNING = -1


def build(force: bool = False) -> str:
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        os.makedirs(os.path.dirname(LIB), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-o", LIB, SRC, "-lm"])
    return LIB


_lib = None


def _load():
    global _lib
    if _lib is None:
        lib = ctypes.CDLL(build())
        i64, p = ctypes.c_int64, ctypes.c_void_p
        lib.vrg_strict_create.restype = p
        lib.vrg_strict_create.argtypes = [p, p, i64, p, i64, i64, i64, ctypes.c_double, i64, i64]
        lib.vrg_strict_destroy.argtypes = [p]
        lib.vrg_strict_destroy.restype = None
        lib.vrg_strict_step.argtypes = [p]
        lib.vrg_strict_step.restype = i64
        lib.vrg_strict_info.argtypes = [p, p]
        lib.vrg_strict_min_margin.argtypes = [p]
        lib.vrg_strict_min_margin.restype = ctypes.c_double
        lib.vrg_strict_get.argtypes = [p, p, p, p, p, p]
        lib.vrg_strict_get.restype = None
        lib.vrg_strict_list.argtypes = [p, ctypes.c_int, p]
        lib.vrg_strict_list.restype = i64
        _lib = lib
    return _lib


def vrg_strict_oracle(data, value_map, H=2.25, max_segment_size=5000, iter_max=200, record_band=False):
    """List-order restatement of ``variationalRegionGrowing`` (VRG:10-121).  Returns a dict:

    ``value_map``   the reference's final valueMap, stale band labels included (uint8)
    ``seg``         segmentedMap (bool);  ``segmented`` rows (z, y, x) in the reference's list order
    ``inner``, ``outer``  the band lists in list order (flat voxel indices)
    ``iterations``, ``exit``, ``trace`` as in ``oracle.vrg_oracle.vrg_oracle``
    ``pin``, ``pout`` the running sums innerProb / outerProb (VRG:132-133) after the last update
    ``bands``       with ``record_band``: per decision (flat band voxel indices in allBnd order, pin/n_in, pout/n_out)
    ``skipped``     listed flips that were in neither band at their turn;  ``dropped`` Q3 count;  ``min_margin``
    """
    lib = _load()
    data = np.ascontiguousarray(data, dtype=np.float64)
    vm = np.ascontiguousarray(value_map).astype(np.uint8)
    if not np.isin(vm, (0, 3, 4)).all():
        raise ValueError("strict oracle: initial valueMap may only hold labels 0, 3 and 4")
    Z, Y, X = data.shape
    levels, idx = np.unique(data, return_inverse=True)
    lev = np.ascontiguousarray(idx.reshape(-1), dtype=np.int32)
    levels = np.ascontiguousarray(levels, dtype=np.float64)
    h = lib.vrg_strict_create(lev.ctypes.data, levels.ctypes.data, len(levels), vm.ctypes.data, Z, Y, X, float(H),
                              int(min(max_segment_size, 2 ** 62)), int(iter_max))
    N = data.size
    info = np.zeros(10, dtype=np.int64)
    pin = np.zeros(N)
    pout = np.zeros(N)
    bands = []

    def lists():
        out = []
        for which in range(3):
            buf = np.zeros(N, dtype=np.int64)
            n = lib.vrg_strict_list(h, which, buf.ctypes.data)
            out.append(buf[:n].copy())
        return out
    try:
        lib.vrg_strict_info(h, info.ctypes.data)
        if info[6] == 0:
            raise ValueError("strict oracle: empty seed set (reference raises IndexError at VRG:88)")
        if info[4] + info[5] == 0:
            raise ValueError("strict oracle: seed has no boundary (reference raises IndexError at VRG:88)")
        while True:
            if record_band:
                lib.vrg_strict_info(h, info.ctypes.data)
                lib.vrg_strict_get(h, None, None, pin.ctypes.data, pout.ctypes.data, None)
                inner, outer, _ = lists()
                b = np.concatenate([inner, outer])
                bands.append((b, pin[b] / info[2], pout[b] / info[3]))
            if lib.vrg_strict_step(h) != EXIT_RUNNING:
                break
        lib.vrg_strict_info(h, info.ctypes.data)
        out_vm = np.zeros(N, dtype=np.uint8)
        sm = np.zeros(N, dtype=np.uint8)
        trace = np.zeros((int(info[7]), 3), dtype=np.int64)
        lib.vrg_strict_get(h, out_vm.ctypes.data, sm.ctypes.data, pin.ctypes.data, pout.ctypes.data, trace.ctypes.data)
        inner, outer, seg = lists()
        mm = float(lib.vrg_strict_min_margin(h))
    finally:
        lib.vrg_strict_destroy(h)
    return {
        "value_map": out_vm.reshape(data.shape), "seg": sm.reshape(data.shape).astype(bool),
        "segmented": np.stack(np.unravel_index(seg, data.shape), axis=1).astype(np.int64).reshape(-1, 3),
        "inner": inner, "outer": outer, "iterations": int(info[0]), "exit": int(info[1]), "trace": trace,
        "pin": pin.reshape(data.shape), "pout": pout.reshape(data.shape), "bands": bands,
        "skipped": int(info[8]), "dropped": int(info[9]), "min_margin": mm,
    }
