"""Generate the golden fixtures by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py            # all small cases (~2 min)
    python tests/golden/make_golden.py c1_128     # config C1, ~10 min single core

Each fixture ``tests/golden/<case>.npz`` holds the inputs (as an integer lattice
``k`` with ``data = k / quantum``; float64 is rebuilt exactly), the reference's
outputs (final valueMap as uint8, printed iteration count, stdout), the
per-iteration trace (n_flips, n_in, n_out) captured around ``update``
(VRG:124), the normalised Parzen sums innerProb/innerSize, outerProb/outerSize
per intensity level at band voxels (VRG:79-82), and the reference's quirk
counters.
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from arterynetwork_b200.phantom import make_phantom  # noqa: E402
from oracle.ref_harness import run_reference  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def case_straight_line():  # VRG:284-298
    vol = np.zeros((50, 50, 150), dtype=int)
    vol[20:22, 20:22, 20:40] = 1
    vm = np.full(vol.shape, 3)
    vm[20:22, 20:22, 22:25] = 0
    return vol, 1, vm, dict(H=2.25, max_segment_size=5000)


def case_sphere():  # VRG:300-314
    x, y, z = np.mgrid[:50, :50, :50]
    vol = ((x - 25) ** 2 + (y - 25) ** 2 + (z - 25) ** 2 <= 100).astype(int)
    vm = np.full(vol.shape, 3)
    vm[25:27, 25:27, 25:27] = 0
    return vol, 1, vm, dict(H=2.25, max_segment_size=5000)


def _noisy_tube(seed, q=128):
    data = np.zeros((16, 16, 28))
    data[6:10, 6:10, 4:24] = 1.0
    rng = np.random.default_rng(seed)
    return np.round((data + rng.normal(0, 0.12, data.shape)) * q).astype(np.int64)


def case_tube_clean():
    k = _noisy_tube(0)
    vm = np.full(k.shape, 3)
    vm[7:9, 7:9, 13:15] = 0
    return k, 128, vm, dict(H=2.25, max_segment_size=None)


def case_edge():  # tube running into the array face: out-of-bounds neighbours are dropped
    data = np.zeros((16, 16, 24))
    data[6:10, 6:10, :] = 1.0
    rng = np.random.default_rng(0)
    k = np.round((data + rng.normal(0, 0.1, data.shape)) * 64).astype(np.int64)
    vm = np.full(k.shape, 3)
    vm[7:9, 7:9, 0:2] = 0
    return k, 64, vm, dict(H=2.25, max_segment_size=None)


def case_maxseg():  # VRG:101 -- cap tested before the flips are applied
    k, q, _, _ = case_edge()
    vm = np.full(k.shape, 3)
    vm[7:9, 7:9, 10:12] = 0
    return k, q, vm, dict(H=2.25, max_segment_size=50)


def case_h1():  # non-default H
    k = _noisy_tube(1)
    vm = np.full(k.shape, 3)
    vm[7:9, 7:9, 13:15] = 0
    return k, 128, vm, dict(H=1.0, max_segment_size=None)


def _phantom_k(shape, seed, **kw):
    data, vm, info = make_phantom(shape, seed=seed, **kw)
    q = info["quantum"]
    return np.rint(data * q).astype(np.int64), q, vm


def case_forest40():
    k, q, vm = _phantom_k((40, 40, 40), 0, cell=(40, 40, 40), margin=3, depth=3, root_r2=9,
                          min_len=8, max_len=16)
    return k, q, vm.astype(np.int64), dict(H=2.25, max_segment_size=None)


def _removal(seed, sigma_k):
    k, q, _ = _phantom_k((32, 32, 32), seed, cell=(32, 32, 32), margin=3, depth=2, root_r2=4,
                         min_len=8, max_len=12, quantum=16, sigma_k=sigma_k)
    vm = np.full(k.shape, 3)
    vm[10:22, 10:22, 10:22] = 0
    return k, q, vm, dict(H=2.25, max_segment_size=None)


def case_removal32():  # big seed cube shrinking onto a thin tube: removals
    return _removal(0, 2)


def case_cancel32():  # removals + cancelled additions (Q1) with no order-dependent follow-up
    return _removal(24, 3)


def case_excl32():  # label-4 voxels and the 4->3 absorption (VRG:137,167-168,177-179)
    data, vm, info = make_phantom((32, 32, 32), seed=0, cell=(32, 32, 32), margin=3, depth=3,
                                  root_r2=9, min_len=8, max_len=12, exclude_below_k=40)
    return np.rint(data * 256).astype(np.int64), 256, vm.astype(np.int64), dict(H=2.25, max_segment_size=None)


def case_tube_cont():  # continuous intensities (no lattice): the brute-force Parzen path, SURVEY.md section 8(f) N1
    data = np.zeros((16, 16, 28))
    data[6:10, 6:10, 4:24] = 1.0
    data = data + np.random.default_rng(0).normal(0, 0.12, data.shape)
    vm = np.full(data.shape, 3)
    vm[7:9, 7:9, 13:15] = 0
    return data, 0, vm, dict(H=2.25, max_segment_size=None)


def case_forest_cont():
    data, vm, _ = make_phantom((24, 24, 24), seed=0, cell=(24, 24, 24), margin=3, depth=2, root_r2=4, min_len=6, max_len=9)
    data = data + np.random.default_rng(5).normal(0, 1e-3, data.shape)  # break the lattice: 13824 distinct values
    return data, 0, vm.astype(np.int64), dict(H=2.25, max_segment_size=None)


def case_excl_cont():  # continuous intensities AND label 4: absorbed voxels join the outside sums (VRG:235,247)
    data = np.zeros((16, 16, 28))
    data[6:10, 6:10, 4:24] = 1.0
    data = data + np.random.default_rng(3).normal(0, 0.12, data.shape)
    vm = np.full(data.shape, 3)
    vm[data <= 0.05] = 4  # the darker two thirds of the background (cf. the commented-out initialisation at VRG:41-43)
    vm[7:9, 7:9, 13:15] = 0
    return data, 0, vm, dict(H=2.25, max_segment_size=None)


def case_forest64_cont():  # a 64^3 forest with the lattice broken by hash noise (262144 distinct values): the continuous mode at
    # a size where its lists and its volume partition are exercised; stored as lattice + noise scale (tests/golden_util.py)
    data, vm, _ = make_phantom((64, 64, 64), seed=2, cell=(64, 64, 64), margin=5, depth=3, root_r2=9, min_len=10, max_len=18)
    return np.rint(data * 256).astype(np.int64), 256, vm.astype(np.int64), dict(H=2.25, max_segment_size=None, noise_scale=1e-3)


def _tube_k(seed, q=128, sigma=0.12, shape=(16, 16, 28)):
    data = np.zeros(shape)
    data[6:10, 6:10, 4:24] = 1.0
    return np.round((data + np.random.default_rng(seed).normal(0, sigma, shape)) * q).astype(np.int64)


def case_forest48_s1():  # a second forest, other seed and size
    k, q, vm = _phantom_k((48, 48, 48), 1, cell=(48, 48, 48), margin=4, depth=3, root_r2=9, min_len=9, max_len=18)
    return k, q, vm.astype(np.int64), dict(H=2.25, max_segment_size=None)


def case_forest_two_trees():  # two trees (two seeds), anisotropic volume, rows longer than 64 voxels
    k, q, vm = _phantom_k((40, 56, 72), 2, cell=(40, 56, 36), margin=4, depth=3, root_r2=9, min_len=8, max_len=14)
    return k, q, vm.astype(np.int64), dict(H=2.25, max_segment_size=None)


def case_tube_h4():  # narrow Parzen kernel
    k = _tube_k(5)
    vm = np.full(k.shape, 3)
    vm[7:9, 7:9, 13:15] = 0
    return k, 128, vm, dict(H=4.0, max_segment_size=None)


def case_tube_two_seeds():  # two seeds in one tube: the fronts meet and merge
    k = _tube_k(6)
    vm = np.full(k.shape, 3)
    vm[7:9, 7:9, 6:8] = 0
    vm[7:9, 7:9, 20:22] = 0
    return k, 128, vm, dict(H=2.25, max_segment_size=None)


def case_int_levels():  # integer data, a handful of levels (the self-tests' dtype, VRG:285,302)
    k = _tube_k(7, q=1, sigma=0.0) * 3 + np.random.default_rng(7).integers(0, 2, (16, 16, 28))
    vm = np.full(k.shape, 3)
    vm[7:9, 7:9, 13:15] = 0
    return k, 1, vm, dict(H=2.25, max_segment_size=None)


def case_excl32_b():  # label 4 on the darkest third only, other tree
    data, vm, info = make_phantom((32, 32, 32), seed=3, cell=(32, 32, 32), margin=3, depth=3, root_r2=9, min_len=8,
                                  max_len=12, exclude_below_k=-10)
    return np.rint(data * 256).astype(np.int64), 256, vm.astype(np.int64), dict(H=2.25, max_segment_size=None)


def case_tube_fat_seed():  # seed cube wider than the tube: removals at its sides while it grows along the tube
    k = _tube_k(8)
    vm = np.full(k.shape, 3)
    vm[5:11, 5:11, 10:16] = 0
    return k, 128, vm, dict(H=2.25, max_segment_size=None)


C1_KW = dict(cell=(128, 128, 128), margin=8, depth=4, root_r2=16, min_len=16, max_len=34)


def case_c1_128():  # BASELINE.json configs[0]
    k, q, vm = _phantom_k((128, 128, 128), 0, **C1_KW)
    return k, q, vm.astype(np.int64), dict(H=2.25, max_segment_size=None)


CASES = {
    "straight_line": case_straight_line,
    "sphere": case_sphere,
    "tube_clean": case_tube_clean,
    "edge": case_edge,
    "maxseg": case_maxseg,
    "h1": case_h1,
    "forest40": case_forest40,
    "removal32": case_removal32,
    "cancel32": case_cancel32,
    "excl32": case_excl32,
    "c1_128": case_c1_128,
    "tube_cont": case_tube_cont,
    "forest_cont": case_forest_cont,
    "excl_cont": case_excl_cont,
    "forest64_cont": case_forest64_cont,
    "forest48_s1": case_forest48_s1,
    "forest_two_trees": case_forest_two_trees,
    "tube_h4": case_tube_h4,
    "tube_two_seeds": case_tube_two_seeds,
    "int_levels": case_int_levels,
    "excl32_b": case_excl32_b,
    "tube_fat_seed": case_tube_fat_seed,
}
SMALL = [c for c in CASES if c not in ("c1_128", "forest64_cont")]


def level_tables(res, data, max_levels=None):
    """Collapse the recorded band samples to one (p_in, p_out) per intensity level."""
    it, lev, pin, pout = [], [], [], []
    for i, s in enumerate(res["band_samples"]):
        v = data[tuple(s["coords"].T)].astype(np.float64)
        u, first = np.unique(v, return_index=True)
        if max_levels is not None and len(u) > max_levels:
            pick = np.linspace(0, len(u) - 1, max_levels).astype(int)
            u, first = u[pick], first[pick]
        it.append(np.full(len(u), i, dtype=np.int32))
        lev.append(u)
        pin.append(s["p_in"][first])
        pout.append(s["p_out"][first])
    return (np.concatenate(it), np.concatenate(lev), np.concatenate(pin), np.concatenate(pout))


def generate(name):
    k, q, vm, kw = CASES[name]()
    continuous = q == 0  # the case returns float64 data itself
    data = k if q in (0, 1) else k.astype(np.float64) / q
    noise_scale = float(kw.get("noise_scale", 0.0))
    if noise_scale > 0:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from golden_util import hash_noise
        data = data + hash_noise(k.size, noise_scale).reshape(k.shape)
    t0 = time.time()
    res = run_reference(data, vm, H=kw["H"], max_segment_size=kw["max_segment_size"],
                        check_drift=(name not in ("c1_128", "forest64_cont")))
    wall = time.time() - t0
    tb_it, tb_lev, tb_pin, tb_pout = level_tables(res, data, max_levels=48 if k.size > 50 ** 3 else None)
    seg_rows = res["segmented"]
    ok_set = np.array_equal(np.sort(np.ravel_multi_index(tuple(seg_rows.T), k.shape)),
                            np.flatnonzero(res["segmented_map"].ravel() == 1))
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"),
        k=(np.zeros(1, np.int16) if continuous else k.astype(np.int16)), quantum=np.int64(q), data_is_int=np.bool_(q == 1),
        noise_scale=np.float64(noise_scale),
        data_f64=(data.astype(np.float64) if continuous else np.zeros(1)),
        value_map_in=np.asarray(vm, dtype=np.uint8),
        H=np.float64(kw["H"]),
        max_segment_size=np.int64(-1 if kw["max_segment_size"] is None else kw["max_segment_size"]),
        labels=res["value_map"].astype(np.uint8),
        seg=np.packbits(res["segmented_map"].astype(bool).ravel()),
        iterations=np.int64(res["iterations"]),
        stdout=np.array(res["stdout"]),
        trace=res["trace"],
        tb_iter=tb_it, tb_level=tb_lev, tb_pin=tb_pin, tb_pout=tb_pout,
        Q2_voxels=np.int64(res["quirks"]["Q2_voxels"]),
        Q3_dropped=np.int64(res["quirks"]["Q3_dropped"]),
        max_drift=np.float64(res["quirks"]["max_drift"]),
        segmented_is_set_of_map=np.bool_(ok_set),
        reference_wall_s=np.float64(wall),
        numpy_version=np.array(np.__version__),
    )
    print("%-14s it=%s n_seg=%d Q2=%d Q3=%d drift=%.2e wall=%.1fs" % (
        name, res["iterations"], len(seg_rows), res["quirks"]["Q2_voxels"],
        res["quirks"]["Q3_dropped"], res["quirks"]["max_drift"], wall), flush=True)


if __name__ == "__main__":
    for n in (sys.argv[1:] or SMALL):
        generate(n)
