"""Minimal NIfTI-1 single-file (.nii / .nii.gz) reader and writer.

nibabel is not installed in this image and the reference uses only a sliver of it:
`nib.load(p).dataobj / .affine / .header`, `nib.Nifti1Image(arr, affine, header)` and
`nib.save(img, p)` (reference DataClass_GP.py:48, vae_reg_GP.py:618-620,
build_model_recons.py:88,113-116).  `vaegam.compat.install_nibabel()` registers this module as
`nibabel` ONLY when no real nibabel is installed, so those call sites keep working unchanged
 and an installed nibabel is never shadowed.
"""
from __future__ import annotations

import gzip
import struct

import numpy as np

_DTYPES = {2: np.uint8, 4: np.int16, 8: np.int32, 16: np.float32, 64: np.float64, 256: np.int8,
           512: np.uint16, 768: np.uint32, 1024: np.int64, 1280: np.uint64}
_CODES = {np.dtype(v).str[1:]: k for k, v in _DTYPES.items()}
SYNTHETIC_PREFIX = "synthetic://"
__version__ = "0.vaegam"


def _qform_affine(raw, endian, pixdim):
    """NIfTI-1 'method 2': rotation from the quaternion (b, c, d), qfac = pixdim[0], offsets qoffset_*."""
    b, c, d, qx, qy, qz = struct.unpack(endian + "6f", raw[256:280])
    a2 = 1.0 - (b * b + c * c + d * d)
    a = np.sqrt(a2) if a2 > 0 else 0.0
    if a2 <= 0:                                   # normalise (b, c, d) for a 180-degree rotation
        n = np.sqrt(b * b + c * c + d * d)
        b, c, d = b / n, c / n, d / n
    R = np.array([[a * a + b * b - c * c - d * d, 2 * (b * c - a * d), 2 * (b * d + a * c)],
                  [2 * (b * c + a * d), a * a + c * c - b * b - d * d, 2 * (c * d - a * b)],
                  [2 * (b * d - a * c), 2 * (c * d + a * b), a * a + d * d - b * b - c * c]])
    qfac = -1.0 if pixdim[0] < 0 else 1.0
    aff = np.eye(4)
    aff[:3, :3] = R * np.array([pixdim[1], pixdim[2], pixdim[3] * qfac])
    aff[:3, 3] = (qx, qy, qz)
    return aff


class Nifti1Header:
    def __init__(self, pixdim=(1.0, 1.0, 1.0, 1.0), sform=None, descrip=b"vaegam-b200"):
        self.pixdim = tuple(float(p) for p in pixdim)
        self.sform = None if sform is None else np.asarray(sform, dtype=np.float64)
        self.descrip = descrip

    def copy(self):
        return Nifti1Header(self.pixdim, None if self.sform is None else self.sform.copy(), self.descrip)


class Nifti1Image:
    def __init__(self, dataobj, affine=None, header=None):
        self.dataobj = np.asarray(dataobj)
        self.affine = np.eye(4) if affine is None else np.asarray(affine, dtype=np.float64)
        self.header = header.copy() if isinstance(header, Nifti1Header) else Nifti1Header()

    def get_fdata(self):
        return np.asarray(self.dataobj, dtype=np.float64)

    @property
    def shape(self):
        return self.dataobj.shape


def _open(path, mode):
    return gzip.open(path, mode) if str(path).endswith(".gz") else open(path, mode)


def load(path) -> Nifti1Image:
    path = str(path)
    if path.startswith(SYNTHETIC_PREFIX):            # synthetic cohorts carry no file: header only
        return Nifti1Image(np.zeros((0,), np.float32), np.eye(4), Nifti1Header())
    with _open(path, "rb") as f:
        raw = f.read()
    if len(raw) < 352:
        raise ValueError(f"{path}: too short for a NIfTI-1 file")
    endian = "<" if struct.unpack("<i", raw[:4])[0] == 348 else ">"
    if struct.unpack(endian + "i", raw[:4])[0] != 348:
        raise ValueError(f"{path}: not a NIfTI-1 header")
    if raw[344:347] != b"n+1":
        raise ValueError(f"{path}: only single-file NIfTI-1 ('n+1') is supported")
    dim = struct.unpack(endian + "8h", raw[40:56])
    datatype, bitpix = struct.unpack(endian + "hh", raw[70:74])
    pixdim = struct.unpack(endian + "8f", raw[76:108])
    vox_offset = int(struct.unpack(endian + "f", raw[108:112])[0])
    slope, inter = struct.unpack(endian + "ff", raw[112:120])
    qform_code, sform_code = struct.unpack(endian + "hh", raw[252:256])
    srow = np.array(struct.unpack(endian + "12f", raw[280:328]), dtype=np.float64).reshape(3, 4)
    if datatype not in _DTYPES:
        raise ValueError(f"{path}: unsupported NIfTI datatype {datatype}")
    shape = tuple(int(d) for d in dim[1:1 + dim[0]])
    dt = np.dtype(_DTYPES[datatype]).newbyteorder(endian)
    n = int(np.prod(shape))
    data = np.frombuffer(raw, dtype=dt, count=n, offset=max(vox_offset, 352)).reshape(shape, order="F")
    if slope not in (0.0, 1.0) or inter != 0.0:
        if slope != 0.0:
            data = data.astype(np.float64) * slope + inter
    affine = np.eye(4)
    if sform_code > 0:                     # nibabel's precedence: sform, then qform, then pixdim
        affine[:3, :] = srow
    elif qform_code > 0:
        affine = _qform_affine(raw, endian, pixdim)
    else:
        affine[0, 0], affine[1, 1], affine[2, 2] = pixdim[1], pixdim[2], pixdim[3]
    return Nifti1Image(data, affine, Nifti1Header(pixdim[1:5], affine if sform_code > 0 else None))


def save(img: Nifti1Image, path):
    data = np.asarray(img.dataobj)
    if data.dtype == np.float16:
        data = data.astype(np.float32)
    key = data.dtype.str[1:]
    if key not in _CODES:
        data = data.astype(np.float64)
        key = data.dtype.str[1:]
    code = _CODES[key]
    hdr = bytearray(352)
    struct.pack_into("<i", hdr, 0, 348)
    dim = [data.ndim] + list(data.shape) + [1] * (7 - data.ndim)
    struct.pack_into("<8h", hdr, 40, *dim)
    struct.pack_into("<hh", hdr, 70, code, data.dtype.itemsize * 8)
    aff = np.asarray(img.affine, dtype=np.float64)
    pix = [float(np.linalg.norm(aff[:3, i])) or 1.0 for i in range(3)]
    tr = img.header.pixdim[3] if isinstance(img.header, Nifti1Header) and len(img.header.pixdim) > 3 else 1.0
    struct.pack_into("<8f", hdr, 76, 1.0, pix[0], pix[1], pix[2], tr, 1.0, 1.0, 1.0)
    struct.pack_into("<f", hdr, 108, 352.0)
    struct.pack_into("<ff", hdr, 112, 1.0, 0.0)
    struct.pack_into("<hh", hdr, 252, 0, 1)                     # sform_code = 1 (scanner)
    struct.pack_into("<12f", hdr, 280, *aff[:3, :].reshape(-1))
    hdr[344:348] = b"n+1\0"
    with _open(path, "wb") as f:
        f.write(bytes(hdr))
        f.write(np.asfortranarray(data).astype(data.dtype.newbyteorder("<"), copy=False).tobytes(order="F"))
