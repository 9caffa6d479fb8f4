"""Minimal NIfTI-1 writer / reader for the prediction volumes (predict.py:204-214 saves `pred-%04d.nii.gz` through
torchio.ScalarImage.save, i.e. nibabel / SimpleITK -- neither is in this image).  Single-file .nii / .nii.gz, header per
the NIfTI-1.1 specification (348 bytes + 4 extension bytes, vox_offset 352), sform + qform from the 4 x 4 RAS affine.

Host-side IO, not on the GPU path: the label volume arrives here after its one device -> host copy."""
import gzip
import struct

import numpy as np

_DTYPES = {np.dtype("uint8"): (2, 8), np.dtype("int16"): (4, 16), np.dtype("int32"): (8, 32), np.dtype("float32"): (16, 32),
           np.dtype("float64"): (64, 64), np.dtype("int8"): (256, 8), np.dtype("uint16"): (512, 16), np.dtype("int64"): (1024, 64)}


def _quaternion(rot):
    """(qfac, b, c, d) of a proper rotation matrix (NIfTI-1 'method 2')."""
    r = np.array(rot, dtype=np.float64)
    qfac = 1.0
    if np.linalg.det(r) < 0:
        r[:, 2] = -r[:, 2]
        qfac = -1.0
    a = r[0, 0] + r[1, 1] + r[2, 2] + 1.0
    if a > 0.5:
        a = 0.5 * np.sqrt(a)
        b, c, d = 0.25 * (r[2, 1] - r[1, 2]) / a, 0.25 * (r[0, 2] - r[2, 0]) / a, 0.25 * (r[1, 0] - r[0, 1]) / a
    else:
        xd, yd, zd = 1.0 + r[0, 0] - (r[1, 1] + r[2, 2]), 1.0 + r[1, 1] - (r[0, 0] + r[2, 2]), 1.0 + r[2, 2] - (r[0, 0] + r[1, 1])
        if xd > 1.0:
            b = 0.5 * np.sqrt(xd)
            c, d, a = 0.25 * (r[0, 1] + r[1, 0]) / b, 0.25 * (r[0, 2] + r[2, 0]) / b, 0.25 * (r[2, 1] - r[1, 2]) / b
        elif yd > 1.0:
            c = 0.5 * np.sqrt(yd)
            b, d, a = 0.25 * (r[0, 1] + r[1, 0]) / c, 0.25 * (r[1, 2] + r[2, 1]) / c, 0.25 * (r[0, 2] - r[2, 0]) / c
        else:
            d = 0.5 * np.sqrt(zd)
            b, c, a = 0.25 * (r[0, 2] + r[2, 0]) / d, 0.25 * (r[1, 2] + r[2, 1]) / d, 0.25 * (r[1, 0] - r[0, 1]) / d
        if a < 0:
            b, c, d = -b, -c, -d
    return qfac, float(b), float(c), float(d)


def save_nifti(path, volume, affine=None):
    """volume: [W, H, D] or [C, W, H, D] array-like (a leading singleton channel is dropped, like torchio does for 3-D
    images); affine: 4 x 4 voxel -> RAS-mm matrix (identity when None).  `.nii.gz` paths are gzip-compressed."""
    vol = np.asarray(volume)
    if vol.ndim == 4 and vol.shape[0] == 1:
        vol = vol[0]
    if vol.dtype == np.bool_:
        vol = vol.astype(np.uint8)
    if vol.dtype not in _DTYPES:
        raise ValueError("unsupported dtype %s" % vol.dtype)
    affine = np.eye(4) if affine is None else np.asarray(affine, dtype=np.float64)
    code, bitpix = _DTYPES[vol.dtype]
    spatial = vol.shape[-3:] if vol.ndim >= 3 else vol.shape
    dim = [vol.ndim if vol.ndim <= 3 else 5] + list(spatial) + [1] * (7 - len(spatial))
    if vol.ndim == 4:          # [C, W, H, D] -> NIfTI stores vector components in the 5th dimension
        dim = [5] + list(spatial) + [1, vol.shape[0], 1, 1]
        data = np.moveaxis(vol, 0, -1)[:, :, :, None, :]
    else:
        data = vol
    zooms = np.sqrt((affine[:3, :3] ** 2).sum(0))
    qfac, qb, qc, qd = _quaternion(affine[:3, :3] / np.where(zooms == 0, 1, zooms))
    pixdim = [qfac] + [float(z) for z in zooms] + [1.0] * 4
    hdr = struct.pack("<i10s18sihcB", 348, b"", b"", 0, 0, b"r", 0)
    hdr += struct.pack("<8h", *dim[:8])
    hdr += struct.pack("<3f", 0.0, 0.0, 0.0)                       # intent_p1..3
    hdr += struct.pack("<4h", 0, code, bitpix, 0)                    # intent_code, datatype, bitpix, slice_start
    hdr += struct.pack("<8f", *pixdim)
    hdr += struct.pack("<f", 352.0)                                  # vox_offset
    hdr += struct.pack("<2f", 1.0, 0.0)                              # scl_slope, scl_inter
    hdr += struct.pack("<hBB", 0, 0, 2)                              # slice_end, slice_code, xyzt_units (mm)
    hdr += struct.pack("<4f", 0.0, 0.0, 0.0, 0.0)                    # cal_max, cal_min, slice_duration, toffset
    hdr += struct.pack("<2i", 0, 0)                                  # glmax, glmin
    hdr += struct.pack("<80s24s", b"b200seg", b"")                   # descrip, aux_file
    hdr += struct.pack("<2h", 1, 1)                                  # qform_code, sform_code (scanner anat)
    hdr += struct.pack("<6f", qb, qc, qd, *[float(v) for v in affine[:3, 3]])
    hdr += struct.pack("<12f", *[float(v) for v in affine[:3, :].reshape(-1)])
    hdr += struct.pack("<16s4s", b"", b"n+1\0")
    assert len(hdr) == 348, len(hdr)
    payload = hdr + b"\0\0\0\0" + np.asfortranarray(data).tobytes(order="F")
    opener = gzip.open if str(path).endswith(".gz") else open
    with opener(path, "wb") as f:
        f.write(payload)


def load_nifti(path):
    """Inverse of save_nifti for its own files (and other little-endian single-file NIfTI-1): returns (array, affine)."""
    opener = gzip.open if str(path).endswith(".gz") else open
    with opener(path, "rb") as f:
        raw = f.read()
    if struct.unpack("<i", raw[:4])[0] != 348 or raw[344:347] != b"n+1":
        raise ValueError("not a little-endian single-file NIfTI-1 image")
    dim = struct.unpack("<8h", raw[40:56])
    code, bitpix = struct.unpack("<2h", raw[70:74])
    vox_offset = int(struct.unpack("<f", raw[108:112])[0])
    dtype = {v[0]: k for k, v in _DTYPES.items()}[code]
    shape = [d for d in dim[1:dim[0] + 1]]
    data = np.frombuffer(raw, dtype=dtype, count=int(np.prod(shape)), offset=vox_offset).reshape(shape, order="F")
    affine = np.eye(4)
    affine[:3, :] = np.array(struct.unpack("<12f", raw[280:328])).reshape(3, 4)
    if len(shape) == 5:
        data = np.moveaxis(data[:, :, :, 0, :], -1, 0)
    return data, affine
