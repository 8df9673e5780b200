"""Host side of the strict (list-order) engine, ``vrg_strict_*`` in include/vrg_b200.h: the reference's variational
region growing WITH the effects of its list processing order (VRG:156-259).  Pure ctypes
pass-through; the work is in ``csrc/vrg_strict.cu``.  No CPU fallback."""
from __future__ import annotations

import ctypes

import numpy as np

from . import _native as nat


class StrictEngine:
    def __init__(self, shape, H=2.25, max_segment_size=5000, iter_max=200, device=0, max_seconds=0.0):
        self.lib = nat.load()
        self.shape = tuple(int(s) for s in shape)
        if len(self.shape) != 3:
            raise ValueError("StrictEngine: shape must be (Z, Y, X)")
        self._h = nat.vp()
        sh = (nat.i64 * 3)(*self.shape)
        nat.check(self.lib.vrg_strict_create(int(device), sh, float(H), int(iter_max), int(min(max_segment_size, 2 ** 62)),
                                             float(max_seconds), ctypes.byref(self._h)))

    def close(self):
        if self._h:
            self.lib.vrg_strict_destroy(self._h)
            self._h = nat.vp()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def init(self, data, value_map):
        data = np.ascontiguousarray(data, dtype=np.float64)
        vm = np.ascontiguousarray(value_map, dtype=np.uint8)
        if data.shape != self.shape or vm.shape != self.shape:
            raise ValueError("StrictEngine.init: arrays must have shape %s" % (self.shape,))
        nat.check(self.lib.vrg_strict_init(self._h, data.ctypes.data, vm.ctypes.data))

    @staticmethod
    def _res(r):
        return {k: int(getattr(r, k)) for k, _ in nat.StrictResult._fields_}

    def step(self):
        r = nat.StrictResult()
        nat.check(self.lib.vrg_strict_step(self._h, ctypes.byref(r)))
        return self._res(r)

    def run(self):
        r = nat.StrictResult()
        nat.check(self.lib.vrg_strict_run(self._h, ctypes.byref(r)))
        return self._res(r)

    def value_map(self):
        out = np.empty(self.shape, dtype=np.uint8)
        nat.check(self.lib.vrg_strict_download(self._h, out.ctypes.data, None))
        return out

    def segmented_map(self):
        out = np.empty(self.shape, dtype=np.uint8)
        nat.check(self.lib.vrg_strict_download(self._h, None, out.ctypes.data))
        return out

    def _list(self, which, sums):
        n = nat.i64(0)
        cap = 1 << 16
        while True:
            vox = np.empty(cap, dtype=np.int64)
            pin = np.empty(cap) if sums else None
            pout = np.empty(cap) if sums else None
            rc = self.lib.vrg_strict_list(self._h, which, vox.ctypes.data, pin.ctypes.data if sums else None,
                                          pout.ctypes.data if sums else None, cap, ctypes.byref(n))
            if rc == nat.ERR_ARG and n.value > cap:
                cap = int(n.value)
                continue
            nat.check(rc)
            k = int(n.value)
            return (vox[:k], pin[:k], pout[:k]) if sums else vox[:k]

    def band(self):
        """(flat voxel indices in allBnd order, innerProb/innerSize, outerProb/outerSize) -- what the next decision reads."""
        return self._list(0, True)

    def segmented(self):
        """Rows (z, y, x) in the reference's list order (VRG:126,172,200)."""
        vox = self._list(1, False)
        return np.stack(np.unravel_index(vox, self.shape), axis=1).astype(np.int64).reshape(-1, 3)

    def trace(self):
        n = nat.i64(0)
        nat.check(self.lib.vrg_strict_get_trace(self._h, None, 0, ctypes.byref(n)))
        rows = np.zeros((int(n.value), 3), dtype=np.int64)
        nat.check(self.lib.vrg_strict_get_trace(self._h, rows.ctypes.data, rows.shape[0], ctypes.byref(n)))
        return rows

    def sums(self):
        pin = np.empty(self.shape)
        pout = np.empty(self.shape)
        nat.check(self.lib.vrg_strict_get_sums(self._h, pin.ctypes.data, pout.ctypes.data))
        return pin, pout
