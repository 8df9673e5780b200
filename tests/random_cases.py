"""Seeded random small VRG inputs for differential tests (GPU vs oracle, oracle vs oracle): odd shapes, rows that cross the
32- and 64-voxel word boundaries, blobs and tubes in noise on different lattices, several seeds, optional label 4, different
H and maxSegmentSize."""
import numpy as np


def random_case(i):
    rng = np.random.default_rng(1000 + i)
    shape = (int(rng.integers(2, 14)), int(rng.integers(3, 20)), int(rng.choice([5, 17, 31, 32, 33, 47, 64, 65, 70, 97])))
    Z, Y, X = shape
    z, y, x = np.ogrid[:Z, :Y, :X]
    bright = np.zeros(shape, dtype=bool)
    centres = []
    for _ in range(int(rng.integers(1, 4))):
        c = np.array([rng.integers(0, Z), rng.integers(0, Y), rng.integers(0, X)])
        centres.append(c)
        if rng.random() < 0.5:  # a tube along x through the centre
            r = int(rng.integers(1, 3))
            bright |= (abs(z - c[0]) <= r) & (abs(y - c[1]) <= r) & (abs(x - c[2]) <= int(rng.integers(3, X)))
        else:
            r = int(rng.integers(1, 5))
            bright |= (z - c[0]) ** 2 + (y - c[1]) ** 2 + (x - c[2]) ** 2 <= r * r
    q = int(rng.choice([1, 8, 32, 256]))
    contrast, sigma = (6, 1.0) if q == 1 else (1.0, float(rng.choice([0.08, 0.15, 0.25])))
    data = bright * contrast + rng.normal(0, sigma, shape)
    data = np.rint(data * q) / q if q > 1 else np.rint(data)
    vm = np.full(shape, 3, dtype=np.uint8)
    if rng.random() < 0.4:
        vm[data <= np.quantile(data, rng.uniform(0.1, 0.6))] = 4
    for c in centres:
        lo = np.maximum(c - rng.integers(0, 2, 3), 0)
        hi = np.minimum(c + 1 + rng.integers(0, 2, 3), shape)
        vm[lo[0]:hi[0], lo[1]:hi[1], lo[2]:hi[2]] = 0
    H = float(rng.choice([1.0, 2.25, 4.0]))
    max_seg = int(rng.choice([10 ** 12, 10 ** 12, 5000, 60]))
    return data.astype(np.float64), vm, H, max_seg


def table_shift_case():
    """A vessel whose first stretch is darker than the rest, with in-between intensities along its wall: the inside mean
    rises while the region grows, the crossover of the two Parzen densities moves, and the decision bits of the in-between
    levels change in the middle of the run (iterations 7 and 11; voxels that entered leave again).  The pipelined run has to
    notice (a sweep is repeated on the new table)."""
    d = np.zeros((20, 24, 64))
    d[8:12, 10:14, 4:60] = 1.0
    d[8:12, 10:14, 4:10] = 0.75
    d[7, 10:14, 4:60:3] = 0.4375
    d[12, 10:14, 5:60:4] = 0.40625
    d[2:4, 2:4, 2:30] = 0.4375
    vm = np.full(d.shape, 3, dtype=np.uint8)
    vm[9:11, 11:13, 5:7] = 0
    return d, vm
