"""B200: the mask-side operations (EDT, 26-connected labelling, vesselness -> vessel mask; SURVEY.md section 8(f) N2 /
N3) through the C-ABI, against the SciPy golden fixtures, the oracle, and size-independent properties."""
import contextlib
import io

import numpy as np
import pytest

from mask_golden_util import load_mask_case, mask_names, rule_names

pytestmark = pytest.mark.gpu


def gvv():
    from arterynetwork_b200 import generateVesselVolume
    return generateVesselVolume


@pytest.mark.parametrize("name", mask_names())
def test_edt_equals_scipy_fixture(name):
    g = load_mask_case(name)
    out = gvv().distance_transform_edt(g["mask"])
    assert out.dtype == np.float64 and out.shape == g["shape"]
    assert np.array_equal(out, g["edt"])


@pytest.mark.parametrize("name", mask_names())
def test_labels_equal_scipy_fixture(name):
    g = load_mask_case(name)
    labeled, result = gvv().labelVolume(g["mask"].astype(int), minSize=10, maxHop=3)
    assert np.array_equal(labeled, g["labels"])
    counts = np.bincount(g["labels"].ravel())
    assert result == [(int(k), int(counts[k])) for k in np.nonzero(counts)[0]]  # GVV:127-131


@pytest.mark.parametrize("name", rule_names())
def test_vessel_mask_equals_fixture(name):
    g = load_mask_case(name)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        out, info = gvv().vesselnessToVesselMask(g["vesselness"], g["brain"], minComponentSize=g["min_size"], return_info=True)
    assert out.dtype == np.uint8
    assert np.array_equal(out, g["vessel_mask"])
    assert buf.getvalue() == "Number of voxels in segmentation: %d\n" % int(g["vessel_mask"].sum())  # GVV:211
    assert info["voxels"] == int(g["vessel_mask"].sum())


def test_f_ordered_inputs_give_the_same_result():
    g = load_mask_case("rule_a")
    with contextlib.redirect_stdout(io.StringIO()):
        a = gvv().vesselnessToVesselMask(np.asfortranarray(g["vesselness"]), np.asfortranarray(g["brain"]),
                                         minComponentSize=g["min_size"])
    assert np.array_equal(a, g["vessel_mask"])
    m = load_mask_case("tubes_to_faces")
    assert np.array_equal(gvv().distance_transform_edt(np.asfortranarray(m["mask"])), m["edt"])


def _random_blobs(shape, seed, n_balls, rmax):
    rng = np.random.default_rng(seed)
    m = np.zeros(shape, dtype=bool)
    z, y, x = np.ogrid[: shape[0], : shape[1], : shape[2]]
    for _ in range(n_balls):
        c = rng.integers(0, shape)
        r = rng.integers(1, rmax + 1)
        m |= (z - c[0]) ** 2 + (y - c[1]) ** 2 + (x - c[2]) ** 2 <= r * r
    return m


@pytest.mark.parametrize("shape,seed", [((40, 52, 70), 0), ((33, 65, 129), 1), ((64, 64, 64), 2)])
def test_edt_and_labels_equal_oracle_on_blobs(shape, seed):
    from oracle import mask_oracle as mo
    m = _random_blobs(shape, seed, 60, 9)
    assert np.array_equal(gvv().distance_transform_edt(m), mo.edt_oracle(m))
    labeled, result = gvv().labelVolume(m)
    olab, ores = mo.label_oracle(m)
    assert np.array_equal(labeled, olab) and result == ores


def test_edt_large_solid_mask_equals_scipy():
    """Distances far larger than any vessel radius (the brain-mask case, GVV:183): a solid ellipsoid of 192 x 256 x 320."""
    ndi = pytest.importorskip("scipy.ndimage")
    Z, Y, X = 192, 256, 320
    z, y, x = np.ogrid[:Z, :Y, :X]
    m = ((z - Z / 2) / (Z / 2 - 3)) ** 2 + ((y - Y / 2) / (Y / 2 - 3)) ** 2 + ((x - X / 2) / (X / 2 - 3)) ** 2 <= 1.0
    out = gvv().distance_transform_edt(m)
    assert np.array_equal(out, ndi.distance_transform_edt(m))
    assert out.max() > 80


def test_edt_properties_at_c2_size():
    """Config C2 shape (512 x 512 x 170): no CPU reference at this size inside a test budget, so check what the domain
    offers -- 0 exactly on the background, >= 1 on the foreground, squared distances are integers, 1-Lipschitz along every
    axis, and every foreground voxel at distance d has a background voxel at exactly that distance (sampled)."""
    from arterynetwork_b200.phantom import forest_segments, rasterize
    shape = (170, 512, 512)
    segs, _ = forest_segments(shape, seed=0)
    m = rasterize(shape, segs)
    d = gvv().distance_transform_edt(m)
    assert np.all(d[~m] == 0) and np.all(d[m] >= 1)
    sq = np.rint(d * d)
    assert np.array_equal(np.sqrt(sq), d)
    for ax in range(3):
        assert np.abs(np.diff(d, axis=ax)).max() <= 1.0 + 1e-12
    rng = np.random.default_rng(0)
    fg = np.argwhere(m)
    bg_mask = ~m
    for p in fg[rng.choice(len(fg), 200, replace=False)]:
        r = int(np.ceil(d[tuple(p)]))
        lo = np.maximum(p - r, 0); hi = np.minimum(p + r + 1, shape)
        sub = bg_mask[lo[0]:hi[0], lo[1]:hi[1], lo[2]:hi[2]]
        zz, yy, xx = np.nonzero(sub)
        d2 = (zz + lo[0] - p[0]) ** 2 + (yy + lo[1] - p[1]) ** 2 + (xx + lo[2] - p[2]) ** 2
        assert d2.min() == sq[tuple(p)]


def test_vessel_mask_c2_size_against_oracle_components():
    """C2-sized vesselness phantom: the GPU mask must equal the rule restated with the GPU's own EDT (checked above)
    and the oracle's flood fill; idempotence: masking the vesselness with the result and re-running changes nothing."""
    from arterynetwork_b200.phantom import make_phantom
    from oracle import mask_oracle as mo
    shape = (85, 256, 256)
    data, _, _ = make_phantom(shape, seed=3, cell=(85, 128, 128))
    z, y, x = np.ogrid[: shape[0], : shape[1], : shape[2]]
    brain = (((z - 42) / 40.0) ** 2 + ((y - 128) / 120.0) ** 2 + ((x - 128) / 120.0) ** 2 <= 1.0).astype(np.uint8)
    with contextlib.redirect_stdout(io.StringIO()):
        out = gvv().vesselnessToVesselMask(data, brain, minComponentSize=150)
        edt = gvv().distance_transform_edt(brain)
    ref = mo.vessel_mask_oracle(data, brain, min_size=150, brain_edt=edt)
    assert np.array_equal(out, ref)
    assert 0 < out.sum() < out.size
    with contextlib.redirect_stdout(io.StringIO()):
        again = gvv().vesselnessToVesselMask(np.where(out, data, data.min()), brain, minComponentSize=150)
    assert np.array_equal(again & out, again)


def test_argument_errors():
    from arterynetwork_b200 import _native as nat
    with pytest.raises(ValueError):
        gvv().distance_transform_edt(np.ones((4, 4, 4)))          # no background voxel
    with pytest.raises(ValueError):
        gvv().distance_transform_edt(np.ones((4, 4)))             # not 3-D
    with pytest.raises(ValueError):
        gvv().labelVolume(np.ones((4, 4, 4)), maxHop=1)
    assert nat.load().vrg_edt(0, None, None, None) == nat.ERR_ARG


def test_component_merge_on_every_pair_of_seven_voxel_rows():
    """The kernel-side twin of tests/test_mask_host.py::test_run_pair_pruning_rule_of_the_component_merge (which checks a Python
    transcription of k_cc_merge's linking rule): all 2^14 pairs of 7-voxel rows, each pair isolated by background, go through
    vrg_label_components itself; labels and sizes must equal the flood-fill oracle's, numbering included."""
    import itertools
    from oracle import mask_oracle as mo
    W = 7
    rows = np.array(list(itertools.product([0, 1], repeat=W)), dtype=bool)  # 128 rows
    per_plane = 64
    n_pairs = len(rows) ** 2
    planes = n_pairs // per_plane
    m = np.zeros((2 * planes, 3 * per_plane, W + 2), dtype=bool)  # every second plane and every third row stay empty
    a_idx, b_idx = np.divmod(np.arange(n_pairs), len(rows))
    z = 2 * (np.arange(n_pairs) // per_plane)
    y = 3 * (np.arange(n_pairs) % per_plane)
    m[z, y, 1:W + 1] = rows[a_idx]
    m[z, y + 1, 1:W + 1] = rows[b_idx]
    labeled, result = gvv().labelVolume(m)
    olab, ores = mo.label_oracle(m)
    assert np.array_equal(labeled, olab)
    assert [tuple(map(int, r)) for r in result] == [tuple(map(int, r)) for r in ores]


def test_edt_lines_fuzz_against_the_oracle():
    """The kernel-side twin of tests/test_mask_host.py::test_edt_line_pass_transcription_equals_brute_force: the same four
    kinds of adversarial lines (sparse zeros, lines without any zero, ramps, dense zeros) as thin volumes along each axis, so
    that k_edt_rows and both k_edt_lines passes each see them, against the min-plus oracle."""
    from oracle import mask_oracle as mo
    rng = np.random.default_rng(0)
    for trial in range(60):
        n = int(rng.integers(1, 70))
        kind = trial % 4
        if kind == 0:
            line = rng.random(n) < 0.85
        elif kind == 1:
            line = np.ones(n, dtype=bool)  # no zero on the line: the distance comes from the other axes (or is "infinite")
        elif kind == 2:
            line = np.ones(n, dtype=bool); line[[0, -1]] = False
        else:
            line = rng.random(n) < 0.5
        for axis in range(3):
            shape = [int(rng.integers(1, 4)), int(rng.integers(1, 4)), int(rng.integers(1, 4))]
            shape[axis] = n
            m = np.ones(shape, dtype=bool)
            idx = [slice(None)] * 3
            for j in range(int(np.prod(shape)) // n):  # every line of the volume along `axis` gets a variant of the pattern
                pos = list(np.unravel_index(j, [s for i, s in enumerate(shape) if i != axis]))
                pos.insert(axis, slice(None))
                v = np.roll(line, j)
                if kind == 1 and j == 0:
                    v = v.copy(); v[n // 2] = False  # one zero in the whole volume: every other line is "infinite" until it sees it
                m[tuple(pos)] = v
            got = gvv().distance_transform_edt(m)
            want = mo.edt_oracle(m)
            if not m.all():  # SciPy's (and the oracle's) result for a mask without any zero is not a distance; skipped
                assert np.array_equal(got, want), (trial, axis, shape)
