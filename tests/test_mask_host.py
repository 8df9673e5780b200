"""CPU: host-side logic of arterynetwork_b200/generateVesselVolume.py that runs before any GPU call."""
import numpy as np
import pytest

from arterynetwork_b200 import generateVesselVolume as gvv


def test_mask_volume_matches_reference_semantics():
    """GVV:86-106: a copy with the voxels outside the mask zeroed; the input is untouched."""
    rng = np.random.default_rng(0)
    vol = rng.normal(size=(3, 4, 5))
    mask = rng.integers(0, 3, size=vol.shape)
    keep = vol.copy()
    out = gvv.maskVolume(vol, mask)
    assert np.array_equal(vol, keep) and out is not vol
    assert np.array_equal(out, np.where(mask == 0, 0.0, vol))


def test_isotropic_view_uses_the_transpose_of_f_ordered_volumes():
    a = np.asfortranarray(np.arange(24, dtype=np.float64).reshape(2, 3, 4))
    v, transposed = gvv._isotropic_view(a, np.float64)
    assert transposed and v.flags.c_contiguous and v.shape == (4, 3, 2)
    assert np.shares_memory(v, a)            # no copy: nibabel-style F-ordered input is used through its transpose
    assert np.array_equal(v.T, a)
    c, transposed = gvv._isotropic_view(np.ascontiguousarray(a), np.float64)
    assert not transposed and c.shape == (2, 3, 4)


def test_argument_errors_raise_before_any_gpu_call():
    with pytest.raises(ValueError):
        gvv.distance_transform_edt(np.ones((4, 4)))
    with pytest.raises(ValueError):
        gvv.labelVolume(np.ones((4, 4, 4)), maxHop=2)
    with pytest.raises(ValueError):
        gvv.vesselnessToVesselMask(np.zeros((2, 3, 4)), np.ones((2, 3, 5)))


def test_run_pair_pruning_rule_of_the_component_merge():
    """DESIGN CHECK of the idea, not of the shipped kernel (its kernel-side twin, same inputs through vrg_label_components, is
    tests/test_gpu_mask.py::test_component_merge_on_every_pair_of_seven_voxel_rows).
    The rule k_cc_merge (csrc/vrg_mask.cu) uses to link a voxel to the previous row -- straight across only if the voxel
    or its neighbour starts an x-run, the left diagonal only across a background voxel and under the same condition, the
    right diagonal across a background voxel always -- transcribed to Python and checked exhaustively on all pairs of rows
    of 7 voxels: it links exactly the pairs of runs that touch (26-connectivity within two rows), each at least once."""
    import itertools
    W = 7

    def runs(row):
        out, start = [], None
        for x, v in enumerate(list(row) + [0]):
            if v and start is None:
                start = x
            if not v and start is not None:
                out.append((start, x - 1))
                start = None
        return out

    def run_of(rs, x):
        return next(i for i, (s, e) in enumerate(rs) if s <= x <= e)

    for A in itertools.product([0, 1], repeat=W):
        ra = runs(A)
        for B in itertools.product([0, 1], repeat=W):
            rb = runs(B)
            touching = {(i, j) for i, (a0, a1) in enumerate(ra) for j, (b0, b1) in enumerate(rb) if a0 - 1 <= b1 and b0 <= a1 + 1}
            linked = set()
            for x in range(W):
                if not A[x]:
                    continue
                pstart = x == 0 or not A[x - 1]
                ql = x > 0 and B[x - 1]
                if B[x]:
                    if pstart or not ql:
                        linked.add((run_of(ra, x), run_of(rb, x)))
                else:
                    if ql and (pstart or x < 2 or not B[x - 2]):
                        linked.add((run_of(ra, x), run_of(rb, x - 1)))
                    if x + 1 < W and B[x + 1]:
                        linked.add((run_of(ra, x), run_of(rb, x + 1)))
            assert linked == touching, (A, B)


def test_edt_line_pass_transcription_equals_brute_force():
    """DESIGN CHECK of the idea, not of the shipped kernel (kernel-side twin: tests/test_gpu_mask.py::
    test_edt_lines_fuzz_against_the_oracle).
    The line pass of the distance transform (k_edt_lines, csrc/vrg_edt.cu) -- lower envelope of parabolas in which the
    zeros inside a stretch of background are neither pushed nor popped -- transcribed to Python and fuzzed against the plain
    double loop, infinite inputs (rows without a zero voxel) included."""
    INF = 1 << 30

    def line_pass(f):
        n = len(f)
        ss, tt, gg = [0] * n, [0] * n, [0] * n
        q, sq, tq, fq, prev = -1, 0, 0, 0, 0
        for u in range(n):
            fu = f[u]
            wanted = (prev > 0 or (u + 1 < n and f[u + 1] > 0)) if fu == 0 else fu < INF
            prev = fu
            if not wanted:
                continue
            while q >= 0:
                if (tq - sq) ** 2 + fq <= (tq - u) ** 2 + fu:
                    break
                q -= 1
                if q >= 0:
                    sq, tq, fq = ss[q], tt[q], gg[q]
            w = 0
            if q >= 0:
                w = 1 + int((u * u - sq * sq + fu - fq) / (2 * (u - sq)))  # truncating division, as in C++
            if w < n:
                q += 1
                sq, tq, fq = u, w, fu
                ss[q], tt[q], gg[q] = u, w, fu
        out = [0] * n
        for u in range(n - 1, -1, -1):
            while q > 0 and u < tq:
                q -= 1
                sq, tq, fq = ss[q], tt[q], gg[q]
            d = 0 if f[u] == 0 else ((u - sq) ** 2 + fq if q >= 0 else INF)
            out[u] = min(d, INF)
        return out

    def brute(f):
        return [min([INF] + [(u - i) ** 2 + v for i, v in enumerate(f) if v < INF]) for u in range(len(f))]

    rng = np.random.default_rng(0)
    for _ in range(3000):
        n = int(rng.integers(1, 40))
        kind = int(rng.integers(0, 4))
        if kind == 0:
            f = rng.integers(0, 3, n) * rng.integers(0, 30, n)
        elif kind == 1:
            f = np.where(rng.random(n) < 0.3, INF, rng.integers(0, 50, n))
        elif kind == 2:
            f = (np.minimum(np.arange(n), np.arange(n)[::-1]) ** 2) * int(rng.integers(0, 2))
        else:
            f = np.where(rng.random(n) < 0.5, 0, rng.integers(1, 400, n))
        f = [int(v) for v in f]
        assert line_pass(f) == brute(f), f


def test_label_volume_refuses_multi_valued_input_before_any_gpu_call():
    """skimage.measure.label connects equal values only; the drop-in labels the non-zero mask: exact for binary volumes, and it
    says so instead of returning different labels for a label map."""
    v = np.zeros((3, 4, 5)); v[0, 0, 0] = 1; v[0, 0, 1] = 2
    with pytest.raises(ValueError):
        gvv.labelVolume(v)
