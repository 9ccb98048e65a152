"""Oracle restatement of the reference's volume reader (test infrastructure).

The reference reads every volume with `nib.load(path).get_fdata()` (pkg/utils/dataloader.py:206-207, 226-227,
239-241) and wraps it with `torch.tensor(...)`.  nibabel (nibabel=4.0.2, environment.yml:154) is a third-party
dependency ABSENT from /root/reference and from this image, and the reference ships no image files: PARITY UNPINNED
for the file format.  What is restated here is nibabel's published behaviour for single-file NIfTI-1 images
(nibabel/nifti1.py `header_dtype`, `Nifti1Header.get_slope_inter`; nibabel/volumeutils.py `array_from_file`,
`apply_read_scaling`):
  * 348-byte header, little or big endian (sizeof_hdr == 348 decides), magic 'n+1', data at vox_offset (>= 352);
  * voxel array stored in Fortran order (dim[1] fastest), returned with shape dim[1:dim[0]+1];
  * get_fdata(): float64; scl_slope == 0 or non-finite -> no scaling, else `arr * slope + inter` (two numpy ops);
  * gzip by file extension.
`write_nifti` produces such files for the tests (the CUDA/C++ staging path is compared with `read_fdata` on them).
"""
import gzip
import struct

import numpy as np

DTYPES = {2: np.uint8, 4: np.int16, 8: np.int32, 16: np.float32, 64: np.float64, 256: np.int8, 512: np.uint16,
          768: np.uint32, 1024: np.int64, 1280: np.uint64}
CODES = {np.dtype(v).name: k for k, v in DTYPES.items()}


def _open(path, mode):
    return gzip.open(path, mode) if str(path).endswith(".gz") else open(path, mode)


def read_header(buf):
    (sz,) = struct.unpack("<i", buf[:4])
    end = "<"
    if sz != 348:
        (sz_be,) = struct.unpack(">i", buf[:4])
        if sz_be != 348:
            raise ValueError("not a NIfTI-1 file")
        end = ">"
    if buf[344:347] != b"n+1":
        raise ValueError("not a single-file NIfTI-1 image")
    dim = struct.unpack(end + "8h", buf[40:56])
    datatype, bitpix = struct.unpack(end + "2h", buf[70:74])
    vox_offset, slope, inter = struct.unpack(end + "3f", buf[108:120])
    return dict(endian=end, shape=tuple(int(d) for d in dim[1:dim[0] + 1]), datatype=datatype, bitpix=bitpix,
                vox_offset=max(int(vox_offset), 352), scl_slope=float(slope), scl_inter=float(inter))


def read_fdata(path):
    """== nib.load(path).get_fdata(): float64 array of shape dim[1..ndim]."""
    with _open(path, "rb") as f:
        buf = f.read()
    h = read_header(buf[:348])
    dt = np.dtype(DTYPES[h["datatype"]]).newbyteorder(h["endian"])
    n = int(np.prod(h["shape"]))
    raw = np.frombuffer(buf, dtype=dt, count=n, offset=h["vox_offset"]).reshape(h["shape"], order="F")
    slope, inter = h["scl_slope"], h["scl_inter"]
    if slope == 0 or not np.isfinite(slope):            # Nifti1Header.get_slope_inter -> (None, None)
        return raw.astype(np.float64)
    arr = raw.astype(np.float64)
    if slope != 1.0:                                    # volumeutils.apply_read_scaling
        arr = arr * slope
    if inter != 0.0:
        arr = arr + inter
    return arr


def write_nifti(path, array, scl_slope=float("nan"), scl_inter=float("nan"), big_endian=False, vox_offset=352,
                pad_dims=0):
    """Single-file NIfTI-1 image holding `array` (stored as is, Fortran order on disk).  `pad_dims` appends singleton
    axes to dim[] (4-D headers of 3-D volumes, as FSL/ANTs write them)."""
    end = ">" if big_endian else "<"
    a = np.asarray(array)
    code = CODES[a.dtype.name]
    shape = tuple(a.shape) + (1,) * pad_dims
    dim = [len(shape)] + list(shape) + [1] * (7 - len(shape))
    hdr = bytearray(348)
    struct.pack_into(end + "i", hdr, 0, 348)
    struct.pack_into(end + "8h", hdr, 40, *dim)
    struct.pack_into(end + "2h", hdr, 70, code, a.dtype.itemsize * 8)
    struct.pack_into(end + "8f", hdr, 76, 1.0, 2.0, 2.0, 2.0, 1.0, 1.0, 1.0, 1.0)      # pixdim (2 mm MNI grid)
    struct.pack_into(end + "3f", hdr, 108, float(vox_offset), scl_slope, scl_inter)
    hdr[344:348] = b"n+1\0"
    data = a.astype(a.dtype.newbyteorder(end)).tobytes(order="F")
    with _open(path, "wb") as f:
        f.write(bytes(hdr) + b"\0" * (vox_offset - 348) + data)
