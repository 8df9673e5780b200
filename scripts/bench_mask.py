#!/usr/bin/env python
"""Timings of the mask-side operations on one B200, device-resident inputs, CUDA events:

    python scripts/bench_mask.py [--shape 170x512x512] [--reps 5] [--cpu]

Prints one JSON line: EDT of the vessel mask (thin foreground) and of a brain-sized ellipsoid (solid foreground), 26-connected
labelling, and the whole vesselness -> vessel-mask rule, each as ms and Gvoxel/s with the algorithmic bytes per voxel
(EDT: 1 B mask + 8 B distance; rule: 8 B vesselness + 1 B brain mask + 1 B vessel mask).  --cpu adds SciPy's times for the
same inputs on this host (the reference's own implementation of these calls)."""
import argparse
import ctypes
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="170x512x512")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--cpu", action="store_true")
    a = ap.parse_args()
    import torch
    import bench
    from arterynetwork_b200 import _native as nat
    shape = tuple(int(t) for t in a.shape.split("x"))
    Z, Y, X = shape
    n = Z * Y * X
    lib = nat.load()
    torch.cuda.set_device(0)
    d_data, _ = bench.device_phantom(shape, 0, 0, Z, 0)
    vessel = (d_data > 0.5).to(torch.uint8).contiguous()          # thin foreground + speckle
    z, y, x = torch.meshgrid(torch.arange(Z, device="cuda"), torch.arange(Y, device="cuda"), torch.arange(X, device="cuda"), indexing="ij")
    brain = ((((z - Z / 2) / (Z / 2 - 2)) ** 2 + ((y - Y / 2) / (Y / 2 - 4)) ** 2 + ((x - X / 2) / (X / 2 - 4)) ** 2) <= 1.0).to(torch.uint8).contiguous()
    del z, y, x
    dist = torch.empty(shape, dtype=torch.float64, device="cuda")
    labels = torch.empty(shape, dtype=torch.int32, device="cuda")
    out = torch.empty(shape, dtype=torch.uint8, device="cuda")
    shp = (nat.i64 * 3)(*shape)
    ncomp = ctypes.c_int64(0)
    info = (nat.i64 * 2)()

    def timed(fn):
        fn()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(a.reps):
            fn()
        e.record()
        torch.cuda.synchronize()
        return s.elapsed_time(e) / a.reps

    res = {"shape": shape, "voxels": n, "reps": a.reps}
    ms = timed(lambda: nat.check(lib.vrg_edt_device(0, vessel.data_ptr(), ctypes.addressof(shp), dist.data_ptr(), None)))
    res["edt_vessel_mask"] = {"ms": ms, "Gvox_s": n / ms / 1e6, "GBps_9B": 9.0 * n / ms / 1e6, "max_distance": float(dist.max())}
    ms = timed(lambda: nat.check(lib.vrg_edt_device(0, brain.data_ptr(), ctypes.addressof(shp), dist.data_ptr(), None)))
    res["edt_brain_mask"] = {"ms": ms, "Gvox_s": n / ms / 1e6, "GBps_9B": 9.0 * n / ms / 1e6, "max_distance": float(dist.max())}
    ms = timed(lambda: nat.check(lib.vrg_label_components_device(0, vessel.data_ptr(), ctypes.addressof(shp), labels.data_ptr(),
                                                                 ctypes.byref(ncomp), None, 0, None)))
    res["label26"] = {"ms": ms, "Gvox_s": n / ms / 1e6, "components": int(ncomp.value)}
    ms = timed(lambda: nat.check(lib.vrg_vessel_mask_device(0, d_data.data_ptr(), brain.data_ptr(), ctypes.addressof(shp), 10.0, 0.8, 0.7,
                                                           150, out.data_ptr(), ctypes.addressof(info), None, None)))
    res["vessel_mask_rule"] = {"ms": ms, "Gvox_s": n / ms / 1e6, "GBps_10B": 10.0 * n / ms / 1e6, "voxels_kept": int(info[0]),
                               "components_kept": int(info[1])}
    if a.cpu:
        from scipy import ndimage as ndi
        hv, hb, hd = vessel.cpu().numpy(), brain.cpu().numpy(), d_data.cpu().numpy()
        t = time.perf_counter(); ev = ndi.distance_transform_edt(hv); res["edt_vessel_mask"]["scipy_ms"] = 1e3 * (time.perf_counter() - t)
        res["edt_vessel_mask"]["equals_scipy"] = bool(np.array_equal(ev, (lib.vrg_edt_device(0, vessel.data_ptr(), ctypes.addressof(shp), dist.data_ptr(), None), dist.cpu().numpy())[1]))
        t = time.perf_counter(); eb = ndi.distance_transform_edt(hb); res["edt_brain_mask"]["scipy_ms"] = 1e3 * (time.perf_counter() - t)
        res["edt_brain_mask"]["equals_scipy"] = bool(np.array_equal(eb, (lib.vrg_edt_device(0, brain.data_ptr(), ctypes.addressof(shp), dist.data_ptr(), None), dist.cpu().numpy())[1]))
        t = time.perf_counter(); lab, k = ndi.label(hv, structure=np.ones((3, 3, 3), dtype=int)); res["label26"]["scipy_ms"] = 1e3 * (time.perf_counter() - t)
        res["label26"]["equals_scipy"] = bool(k == ncomp.value and np.array_equal(lab, labels.cpu().numpy()))
        t = time.perf_counter()
        v = hd.copy(); lo, hi = v.min(), v.max()
        v[np.logical_and(eb <= 10, v <= lo + 0.8 * (hi - lo))] = 0
        v[v <= lo + 0.7 * (hi - lo)] = 0
        v[v != 0] = 1
        lab, k = ndi.label(v, structure=np.ones((3, 3, 3), dtype=int))
        cnt = np.bincount(lab.ravel())
        v[cnt[lab] <= 150] = 0
        res["vessel_mask_rule"]["scipy_ms_without_edt"] = 1e3 * (time.perf_counter() - t)
        res["vessel_mask_rule"]["equals_scipy"] = bool(np.array_equal(v.astype(np.uint8), out.cpu().numpy()))
        res["cpu_threads"] = 1
    print(json.dumps(res))


if __name__ == "__main__":
    main()
