"""CPU: the NIfTI-1 writer (predict.py:204-214 -> pred-%04d.nii.gz) round-trips data and affine, and its header follows the
NIfTI-1.1 layout (field offsets checked against the specification, not against our own reader)."""
import gzip
import struct

import numpy as np


def test_nifti_roundtrip_and_header_layout(tmp_path):
    from b200seg.utils.nifti import load_nifti, save_nifti
    rng = np.random.default_rng(0)
    affine = np.array([[0.0, -0.8, 0.0, 12.5], [1.2, 0.0, 0.0, -30.0], [0.0, 0.0, 2.5, 7.0], [0, 0, 0, 1.0]])
    for dtype, ext in ((np.uint8, ".nii.gz"), (np.float32, ".nii"), (np.int16, ".nii.gz")):
        vol = (rng.random((5, 7, 3)) * 100).astype(dtype)
        path = str(tmp_path / ("a" + ext))
        save_nifti(path, vol[None], affine)          # [1, W, H, D] like the reference's prediction tensors
        back, aff = load_nifti(path)
        assert back.dtype == dtype and np.array_equal(back, vol) and np.allclose(aff, affine, atol=1e-6)
        raw = (gzip.open if ext.endswith(".gz") else open)(path, "rb").read()
        assert struct.unpack("<i", raw[0:4])[0] == 348                       # sizeof_hdr
        assert struct.unpack("<8h", raw[40:56])[:4] == (3, 5, 7, 3)          # dim
        assert struct.unpack("<f", raw[108:112])[0] == 352.0                 # vox_offset
        assert raw[344:348] == b"n+1\0"                                      # magic
        assert np.allclose(struct.unpack("<8f", raw[76:108])[1:4], [1.2, 0.8, 2.5], atol=1e-6)   # pixdim = column norms
        assert struct.unpack("<2h", raw[252:256]) == (1, 1)                  # qform_code, sform_code
        # first voxel values follow the header in Fortran (x fastest) order
        first = np.frombuffer(raw, dtype=dtype, count=5, offset=352)
        assert np.array_equal(first, vol[:, 0, 0])
    multi = rng.random((2, 4, 4, 4)).astype(np.float32)
    save_nifti(str(tmp_path / "m.nii.gz"), multi)
    back, _ = load_nifti(str(tmp_path / "m.nii.gz"))
    assert np.array_equal(back, multi)
