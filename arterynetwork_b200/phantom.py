"""Synthetic vessel-tree phantoms for the VRG hot path.

Everything that decides a voxel's value is INTEGER arithmetic, so the NumPy
generator here and the CUDA generator in ``csrc/vrg_phantom.cu`` produce the
same volume bit for bit (the 2048x2048x1024 config does not fit host RAM and is
generated on the device, slab by slab):

* geometry: a forest of branching tube trees, one per grid cell, given as
  integer segments ``(a, b, r2)``; a voxel p is inside a tube iff its exact
  squared distance to the segment is <= r2 (tested without division);
* noise: a counter-based hash (splitmix64 of the global voxel index) whose four
  16-bit fields are summed (Irwin-Hall, ~Gaussian) and scaled by an integer;
* value: ``(quantum * inside + noise) / quantum`` as float64 -- a k/quantum
  lattice, the reference's dtype with a few hundred
  distinct levels.

Seeds for the region growing are 2x2x2 cubes of label 0 at each tree's root in
an otherwise all-3 ``valueMap`` (the reference's own test set-up, VRG:288-289).
Trees live in disjoint cells with a gap, so growth fronts of different seeds
never merge (merging fronts trigger the reference's order-dependent
bookkeeping).
"""
from __future__ import annotations

import numpy as np

IH_SIGMA = 37837  # std of the sum of four uniform 16-bit fields, 2*65536/sqrt(12)
IH_MEAN = 131070  # 4 * 65535 / 2
_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)
_GOLD = np.uint64(0x9E3779B97F4A7C15)

SEG_COLS = 8  # az, ay, ax, bz, by, bx, r2, tree


def splitmix64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        z = x.astype(np.uint64)
        z = (z ^ (z >> np.uint64(30))) * _M1
        z = (z ^ (z >> np.uint64(27))) * _M2
        return z ^ (z >> np.uint64(31))


def noise_k(lin_index: np.ndarray, seed: int, sigma_k: int) -> np.ndarray:
    """Integer noise (in lattice units) at global linear voxel indices."""
    with np.errstate(over="ignore"):
        key = lin_index.astype(np.uint64) + np.uint64(seed + 1) * _GOLD
    z = splitmix64(key)
    m = np.uint64(0xFFFF)
    s = (z & m) + ((z >> np.uint64(16)) & m) + ((z >> np.uint64(32)) & m) + (z >> np.uint64(48))
    s = s.astype(np.int64)
    return (s * sigma_k) // IH_SIGMA - (IH_MEAN * sigma_k) // IH_SIGMA


def forest_segments(shape, seed=0, cell=(160, 160, 160), margin=6, depth=4,
                    root_r2=36, min_len=18, max_len=46):
    """Integer tube segments of one tree per cell, plus each tree's root voxel.

    Returns ``(segments[int64, n x 8], roots[int64, t x 3])`` in (z, y, x) order.
    """
    rng = np.random.default_rng(seed)
    Z, Y, X = shape
    ncell = [max(1, int(round(n / c))) for n, c in zip(shape, cell)]
    csize = [n / k for n, k in zip(shape, ncell)]
    radii = [root_r2, 16, 9, 4, 2, 1, 1, 1]
    segs, roots = [], []
    tree = 0
    for cz in range(ncell[0]):
        for cy in range(ncell[1]):
            for cx in range(ncell[2]):
                lo = np.array([int(cz * csize[0]), int(cy * csize[1]), int(cx * csize[2])]) + margin
                hi = np.array([int((cz + 1) * csize[0]), int((cy + 1) * csize[1]),
                               int((cx + 1) * csize[2])]) - margin - 1
                if np.any(hi - lo < 8):
                    continue
                root = (lo + hi) // 2
                roots.append(root)
                frontier = [(root, None, 0)]
                while frontier:
                    node, direction, lvl = frontier.pop()
                    nchild = 3 if lvl == 0 else 2
                    for _ in range(nchild):
                        v = rng.normal(size=3)
                        if direction is not None:
                            v = v * 0.9 + direction * 1.1
                        v /= np.linalg.norm(v) + 1e-12
                        length = rng.integers(min_len, max_len + 1) * (0.85 ** lvl)
                        end = np.rint(node + v * length).astype(np.int64)
                        end = np.minimum(np.maximum(end, lo), hi)
                        if np.all(end == node):
                            continue
                        segs.append([*node, *end, radii[min(lvl, len(radii) - 1)], tree])
                        if lvl + 1 < depth:
                            frontier.append((end, v, lvl + 1))
                tree += 1
    segments = np.asarray(segs, dtype=np.int64).reshape(-1, SEG_COLS)
    return segments, np.asarray(roots, dtype=np.int64).reshape(-1, 3)


def rasterize(shape, segments, z0=0, nz=None) -> np.ndarray:
    """Boolean tube mask of planes [z0, z0+nz) -- exact integer point/segment test."""
    Z, Y, X = shape
    nz = Z - z0 if nz is None else nz
    mask = np.zeros((nz, Y, X), dtype=bool)
    for az, ay, ax, bz, by, bx, r2, _ in np.asarray(segments, dtype=np.int64):
        r = int(np.ceil(np.sqrt(r2))) + 1
        zlo, zhi = max(min(az, bz) - r, z0), min(max(az, bz) + r + 1, z0 + nz)
        ylo, yhi = max(min(ay, by) - r, 0), min(max(ay, by) + r + 1, Y)
        xlo, xhi = max(min(ax, bx) - r, 0), min(max(ax, bx) + r + 1, X)
        if zlo >= zhi or ylo >= yhi or xlo >= xhi:
            continue
        zz, yy, xx = np.meshgrid(np.arange(zlo, zhi), np.arange(ylo, yhi), np.arange(xlo, xhi),
                                 indexing="ij", sparse=True)
        inside = segment_test(zz, yy, xx, az, ay, ax, bz, by, bx, r2)
        mask[zlo - z0:zhi - z0, ylo:yhi, xlo:xhi] |= inside
    return mask


def segment_test(pz, py, px, az, ay, ax, bz, by, bx, r2):
    """dist(p, segment ab)^2 <= r2 in exact int64 arithmetic (no division)."""
    dz, dy, dx = bz - az, by - ay, bx - ax
    wz, wy, wx = pz - az, py - ay, px - ax
    c1 = wz * dz + wy * dy + wx * dx
    c2 = dz * dz + dy * dy + dx * dx
    w2 = wz * wz + wy * wy + wx * wx
    ez, ey, ex = pz - bz, py - by, px - bx
    e2 = ez * ez + ey * ey + ex * ex
    mid = (w2 * c2 - c1 * c1) <= r2 * c2
    return np.where(c1 <= 0, w2 <= r2, np.where(c1 >= c2, e2 <= r2, mid))


def seed_value_map(shape, roots, z0=0, nz=None, dtype=np.uint8) -> np.ndarray:
    """All-3 valueMap with a 2x2x2 cube of 0 at every root (cf. VRG:288-289)."""
    Z, Y, X = shape
    nz = Z - z0 if nz is None else nz
    vm = np.full((nz, Y, X), 3, dtype=dtype)
    for rz, ry, rx in np.asarray(roots, dtype=np.int64):
        for z in (rz, rz + 1):
            if z0 <= z < z0 + nz:
                vm[z - z0, ry:ry + 2, rx:rx + 2] = 0
    return vm


def make_phantom(shape, seed=0, quantum=256, sigma_k=31, exclude_below_k=None,
                 z0=0, nz=None, **forest_kw):
    """Float64 intensity volume (k/quantum lattice) and the seeded valueMap.

    ``exclude_below_k``: voxels with lattice value <= that are given label 4
    (the commented-out initialisation at VRG:41-43), exercising the 4->3 path.
    Returns ``(data, value_map, info)``; ``z0``/``nz`` select a z-slab of the same
    global phantom (bit-identical to slicing the whole volume).
    """
    Z, Y, X = shape
    nz = Z - z0 if nz is None else nz
    segments, roots = forest_segments(shape, seed=seed, **forest_kw)
    tube = rasterize(shape, segments, z0, nz)
    lin = (np.arange(z0, z0 + nz, dtype=np.int64)[:, None, None] * Y
           + np.arange(Y, dtype=np.int64)[None, :, None]) * X + np.arange(X, dtype=np.int64)[None, None, :]
    k = tube.astype(np.int64) * quantum + noise_k(lin, seed, sigma_k)
    data = k.astype(np.float64) / float(quantum)
    vm = seed_value_map(shape, roots, z0, nz)
    if exclude_below_k is not None:
        vm[(k <= exclude_below_k) & (vm != 0)] = 4
    info = {"segments": segments, "roots": roots, "tube_voxels": int(tube.sum()),
            "quantum": quantum, "sigma_k": sigma_k, "seed": seed}
    return data, vm, info
