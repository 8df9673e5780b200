"""Rule fixtures produced by EXECUTING the reference's own ``main()`` (Code/generateVesselVolume.py:138-229), unmodified,
in the build container:

    python tests/golden/mask/make_golden_rule_ref.py

``main()`` is file-driven (nibabel in, nibabel out) and imports packages that are not installed here, so the harness
stands in for the storage layer only:

* ``nibabel``           -- an in-memory module: ``load(path)`` hands out the volume registered under the file's name
                           (``get_data()`` / ``.affine``), ``save(img, path)`` keeps the image that ``main()`` writes;
* ``skimage.measure``   -- ``label(volume, return_num=True, connectivity=3)`` is served by ``scipy.ndimage.label`` with the
                           full 3x3x3 structure (26-connectivity, raster-order numbering: on the 0/1 volume ``main()``
                           passes the two agree label for label);
* ``matplotlib``        -- empty stand-ins (imported at GVV:10-11, never called).

The reference's source is compiled from a scratch copy under the system temp dir (``main()`` writes its distance-transform
cache next to its own file, GVV:181-184, and /root/reference is read-only); nothing of it is copied into the repository.
Every array operation between loading and saving -- the EDT of the brain mask, both cut-offs, the binarisation, the
labelling call, the 150-voxel filter (GVV:176-200) -- is the reference's own code.  Each ``rule_ref_<case>.npz`` holds the
inputs (vesselness as integer lattice k / quantum, brain mask) and the uint8 mask ``main()`` saved.
"""
import os
import shutil
import sys
import tempfile
import types

import numpy as np
import scipy
from scipy import ndimage as ndi

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden_mask import rule_case  # noqa: E402  (the same synthetic vesselness volumes)

REF = "/root/reference/Code/generateVesselVolume.py"
FULL = np.ones((3, 3, 3), dtype=int)


def run_reference_main(vesselness, brain):
    """vesselVolumeMask as the unmodified main() writes it for these two volumes."""
    store = {"vesselnessFiltered.nii.gz": vesselness, "brainVolumeMask.nii.gz": brain.astype(np.uint8),
             "brainVolume.nii.gz": vesselness * 0.0, "401 3D MRA BRAIN.nii.gz": vesselness * 0.0}
    saved = {}

    class Img:
        def __init__(self, data, affine=None):
            self._d, self.affine = data, (np.eye(4) if affine is None else affine)

        def get_data(self):
            return self._d

    nib = types.ModuleType("nibabel")
    nib.load = lambda path: Img(np.array(store[os.path.basename(path)], copy=True))
    nib.Nifti1Image = Img
    nib.save = lambda img, path: saved.__setitem__(os.path.basename(path), img.get_data())
    sk = types.ModuleType("skimage")
    skm = types.ModuleType("skimage.measure")

    def label(volume, return_num=False, connectivity=None):
        assert connectivity == 3
        lab, n = ndi.label(volume, structure=FULL)
        return (lab, n) if return_num else lab
    skm.label = label
    sk.measure = skm
    mpl = types.ModuleType("matplotlib")
    plt = types.ModuleType("matplotlib.pyplot")
    mpl.pyplot = plt
    stubs = {"nibabel": nib, "skimage": sk, "skimage.measure": skm, "matplotlib": mpl, "matplotlib.pyplot": plt}
    keep = {k: sys.modules.get(k) for k in stubs}
    tmp = tempfile.mkdtemp(prefix="gvv_ref_")
    try:
        sys.modules.update(stubs)
        path = os.path.join(tmp, "generateVesselVolume.py")
        shutil.copyfile(REF, path)  # scratch copy: main() writes its cache beside __file__
        mod = types.ModuleType("gvv_reference")
        mod.__file__ = path
        with open(path) as f:
            exec(compile(f.read(), path, "exec"), mod.__dict__)
        mod.main()
    finally:
        for k, v in keep.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
        shutil.rmtree(tmp, ignore_errors=True)
    return saved["vesselVolumeMask.nii.gz"]


def main():
    import warnings
    warnings.simplefilter("ignore")
    for name, (seed, shape, nt) in {"rule_ref_a": (1, (24, 40, 56), 14), "rule_ref_b": (5, (40, 48, 64), 30),
                                    "rule_ref_c": (3, (16, 64, 64), 10)}.items():
        k, q, brain = rule_case(seed, shape, nt)
        if name == "rule_ref_b":  # one fat structure so that a component survives the 150-voxel filter
            k[15:25, 20:28, 10:54] = 60
        vessel = run_reference_main(k.astype(np.float64) / q, brain)
        assert vessel.dtype == np.uint8
        np.savez_compressed(os.path.join(HERE, name + ".npz"), shape=np.array(shape), k=k.astype(np.int16), quantum=q,
                            brain=np.packbits(brain), vessel_mask=np.packbits(vessel.astype(bool)), min_size=150,
                            brain_edt=ndi.distance_transform_edt(brain), scipy_version=scipy.__version__,
                            produced_by="Code/generateVesselVolume.py main(), unmodified (storage layer stubbed)")
        print(name, shape, "vessel voxels", int(vessel.sum()))


if __name__ == "__main__":
    main()
