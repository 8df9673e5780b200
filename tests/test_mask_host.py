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
