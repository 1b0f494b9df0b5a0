"""Dependency-free NIfTI-1 single-file (.nii / .nii.gz) reader and writer.

The reference reads volumes through SimpleITK and hands them to the generator as
``(x, y, z)``-ordered tensors (``fetalsyngen/utils/image_reading.py:32-55``: the SimpleITK
``(z, y, x)`` array is permuted ``(2, 1, 0)``).  Neither SimpleITK nor nibabel exists in the
target image, so the host side of this path carries its own codec.  Arrays are returned in
``(x, y, z)`` order, C-contiguous (z fastest), which is the layout every kernel assumes.
"""
from __future__ import annotations

import gzip
import struct
from pathlib import Path

import numpy as np

_CODE_TO_DTYPE = {
    2: np.uint8,
    4: np.int16,
    8: np.int32,
    16: np.float32,
    64: np.float64,
    256: np.int8,
    512: np.uint16,
    768: np.uint32,
}
_DTYPE_TO_CODE = {np.dtype(v): k for k, v in _CODE_TO_DTYPE.items()}


class NiftiError(ValueError):
    pass


def read_nifti(path, with_affine: bool = False):
    """Read a 3-D NIfTI-1 volume.  Returns ``arr[x, y, z]`` (C-contiguous) and optionally the sform."""
    path = str(path)
    raw = gzip.open(path, "rb").read() if path.endswith(".gz") else Path(path).read_bytes()
    if len(raw) < 352:
        raise NiftiError(f"{path}: too short for a NIfTI-1 header")
    hdr_len = struct.unpack("<i", raw[0:4])[0]
    endian = "<"
    if hdr_len != 348:
        if struct.unpack(">i", raw[0:4])[0] != 348:
            raise NiftiError(f"{path}: sizeof_hdr is not 348")
        endian = ">"
    dim = struct.unpack(endian + "8h", raw[40:56])
    if not 3 <= dim[0] <= 7 or any(d != 1 for d in dim[4 : 1 + dim[0]]):
        raise NiftiError(f"{path}: expected a 3-D volume, dim={dim}")
    code = struct.unpack(endian + "h", raw[70:72])[0]
    if code not in _CODE_TO_DTYPE:
        raise NiftiError(f"{path}: unsupported datatype code {code}")
    vox_offset = int(struct.unpack(endian + "f", raw[108:112])[0])
    slope, inter = struct.unpack(endian + "2f", raw[112:120])
    nx, ny, nz = dim[1:4]
    dt = np.dtype(_CODE_TO_DTYPE[code]).newbyteorder(endian)
    flat = np.frombuffer(raw, dtype=dt, count=nx * ny * nz, offset=vox_offset)
    # file order is x fastest; present as [x, y, z] with z fastest
    arr = np.ascontiguousarray(flat.reshape((nz, ny, nx)).transpose(2, 1, 0)).astype(dt.newbyteorder("="), copy=False)
    if slope not in (0.0, 1.0) or inter != 0.0:
        if np.isfinite(slope) and slope != 0.0:
            arr = arr.astype(np.float32) * np.float32(slope) + np.float32(inter)
    if not with_affine:
        return arr
    aff = np.eye(4)
    qform_code, sform_code = struct.unpack(endian + "2h", raw[252:256])
    pix = struct.unpack(endian + "8f", raw[76:108])
    if sform_code > 0:
        aff[:3] = np.array(struct.unpack(endian + "12f", raw[280:328]), dtype=np.float64).reshape(3, 4)
    elif qform_code > 0:  # quaternion form (NIfTI-1 standard, method 2)
        b, c, d, qx, qy, qz = struct.unpack(endian + "6f", raw[256:280])
        a = np.sqrt(max(0.0, 1.0 - (b * b + c * c + d * d)))
        rot = np.array([[a * a + b * b - c * c - d * d, 2 * (b * c - a * d), 2 * (b * d + a * c)],
                        [2 * (b * c + a * d), a * a + c * c - b * b - d * d, 2 * (c * d - a * b)],
                        [2 * (b * d - a * c), 2 * (c * d + a * b), a * a + d * d - b * b - c * c]])
        qfac = -1.0 if pix[0] < 0 else 1.0
        aff[:3, :3] = rot * np.array([pix[1], pix[2], pix[3] * qfac])[None, :]
        aff[:3, 3] = (qx, qy, qz)
    else:
        aff[0, 0], aff[1, 1], aff[2, 2] = pix[1:4]
    return arr, aff


def ras_axes(affine) -> tuple[tuple[int, int, int], tuple[bool, bool, bool]]:
    """(perm, flip): output axis w of the RAS-oriented array is input axis perm[w], reversed when flip[w] — the
    closest axis permutation to the affine's rotation (what monai's ``Orientation("RAS")`` / nibabel's
    ``io_orientation`` compute; the reference applies it to every volume it loads, datasets.py:284-286,
    rand_gmm.py:91-96)."""
    from scipy.optimize import linear_sum_assignment

    r = np.asarray(affine, dtype=np.float64)[:3, :3]
    norm = np.sqrt((r**2).sum(0))
    norm[norm == 0] = 1.0
    rn = r / norm
    world, vox = linear_sum_assignment(-np.abs(rn))  # world axis w <-> voxel axis vox[w]
    perm = tuple(int(v) for v in vox[np.argsort(world)])
    flip = tuple(bool(rn[w, perm[w]] < 0) for w in range(3))
    return perm, flip


def to_ras(arr: np.ndarray, affine) -> tuple[np.ndarray, np.ndarray]:
    """Reorient ``arr[x, y, z]`` to RAS storage order; returns (array, updated affine).  Identity (no copy) for
    volumes that are stored RAS already, like the reference's bundled data."""
    perm, flip = ras_axes(affine)
    aff = np.asarray(affine, dtype=np.float64).copy()
    if perm == (0, 1, 2) and not any(flip):
        return arr, aff
    out = np.transpose(arr, perm)
    aff[:3, :3] = aff[:3, :3][:, list(perm)]
    for w in range(3):
        if flip[w]:
            out = np.flip(out, axis=w)
            aff[:3, 3] = aff[:3, 3] + aff[:3, w] * (out.shape[w] - 1)
            aff[:3, w] = -aff[:3, w]
    return np.ascontiguousarray(out), aff


def write_nifti(path, arr: np.ndarray, affine: np.ndarray | None = None) -> None:
    """Write ``arr[x, y, z]`` as NIfTI-1 (gzip if the name ends in .gz)."""
    arr = np.asarray(arr)
    if arr.ndim != 3:
        raise NiftiError("write_nifti expects a 3-D array")
    if arr.dtype == np.int64:
        arr = arr.astype(np.int32)
    if arr.dtype == np.bool_:
        arr = arr.astype(np.uint8)
    if arr.dtype not in _DTYPE_TO_CODE:
        raise NiftiError(f"unsupported dtype {arr.dtype}")
    aff = np.eye(4) if affine is None else np.asarray(affine, dtype=np.float64)
    hdr = bytearray(352)
    struct.pack_into("<i", hdr, 0, 348)
    struct.pack_into("<8h", hdr, 40, 3, arr.shape[0], arr.shape[1], arr.shape[2], 1, 1, 1, 1)
    struct.pack_into("<h", hdr, 70, _DTYPE_TO_CODE[arr.dtype])
    struct.pack_into("<h", hdr, 72, arr.dtype.itemsize * 8)
    vox = np.sqrt((aff[:3, :3] ** 2).sum(0))
    struct.pack_into("<8f", hdr, 76, 1.0, vox[0], vox[1], vox[2], 1.0, 1.0, 1.0, 1.0)
    struct.pack_into("<f", hdr, 108, 352.0)
    struct.pack_into("<2f", hdr, 112, 1.0, 0.0)
    hdr[123] = 2  # xyzt_units: mm
    struct.pack_into("<hh", hdr, 252, 0, 1)  # qform_code, sform_code
    struct.pack_into("<12f", hdr, 280, *aff[:3].reshape(-1))
    hdr[344:348] = b"n+1\0"
    payload = np.ascontiguousarray(arr.transpose(2, 1, 0)).tobytes()
    path = str(path)
    if path.endswith(".gz"):
        with gzip.open(path, "wb", compresslevel=1) as f:
            f.write(bytes(hdr))
            f.write(payload)
    else:
        with open(path, "wb") as f:
            f.write(bytes(hdr))
            f.write(payload)
