"""ctypes binding of the host-side input staging library (include/adni_staging.h, csrc/stage/nifti_stage.cpp):
gzip NIfTI-1 volumes decoded by native threads straight into (pinned) torch buffers.  The reference's counterpart
is `nib.load(path).get_fdata()` inside 32 DataLoader worker processes (pkg/utils/dataloader.py:206-241,
train_anat_cnn.py:187-198).  There is no Python/numpy fallback: a missing library raises."""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "csrc", "libadni_stage.so")
_lib = None

SYMBOLS = ["adni_stage_last_error", "adni_stage_version", "adni_nifti_read_info", "adni_nifti_read_f64",
           "adni_nifti_read_f32", "adni_nifti_read_mask_u8", "adni_stage_volumes"]


class NiftiInfo(ctypes.Structure):
    _fields_ = [("ndim", ctypes.c_int32), ("dim", ctypes.c_int64 * 7), ("datatype", ctypes.c_int32),
                ("bitpix", ctypes.c_int32), ("swapped", ctypes.c_int32), ("scaled", ctypes.c_int32),
                ("scl_slope", ctypes.c_double), ("scl_inter", ctypes.c_double), ("vox_offset", ctypes.c_int64),
                ("nvox", ctypes.c_int64)]

    @property
    def shape(self):
        return tuple(int(self.dim[i]) for i in range(self.ndim))


class StagingError(RuntimeError):
    pass


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise RuntimeError(f"{_LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(the staging path has no Python fallback)")
        lib = ctypes.CDLL(_LIB_PATH)
        lib.adni_stage_last_error.restype = ctypes.c_char_p
        for name in ("adni_nifti_read_f64", "adni_nifti_read_f32", "adni_nifti_read_mask_u8"):
            getattr(lib, name).argtypes = [ctypes.c_char_p, ctypes.c_void_p, ctypes.c_int64, ctypes.POINTER(NiftiInfo)]
        lib.adni_nifti_read_info.argtypes = [ctypes.c_char_p, ctypes.POINTER(NiftiInfo)]
        lib.adni_stage_volumes.argtypes = [ctypes.POINTER(ctypes.c_char_p), ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                           ctypes.c_int64, ctypes.c_int64, ctypes.c_int, ctypes.POINTER(ctypes.c_int)]
        _lib = lib
    return _lib


def _check(rc):
    if rc != 0:
        msg = load().adni_stage_last_error().decode(errors="replace")
        if rc == -5:
            raise FileNotFoundError(msg) if "cannot open" in msg else StagingError(msg)
        if rc == -1:
            raise ValueError(msg)
        raise StagingError(msg)


def _squeeze_trailing(shape):
    shape = list(shape)
    while len(shape) > 3 and shape[-1] == 1:
        shape.pop()
    return tuple(shape)


def read_info(path):
    info = NiftiInfo()
    _check(load().adni_nifti_read_info(os.fsencode(path), ctypes.byref(info)))
    return info


def read_volume(path, dtype=torch.float32, out=None, pin=False):
    """torch.tensor(nib.load(path).get_fdata()) as fp32 (staging format), fp64 (the reference's own dtype) or uint8
    (brain mask).  `out` = a preallocated contiguous CPU tensor (e.g. a slot of a pinned batch buffer)."""
    info = read_info(path)
    shape = info.shape
    if out is None:
        out = torch.empty(shape, dtype=dtype, pin_memory=pin)
    assert out.device.type == "cpu" and out.is_contiguous()
    fn = {torch.float32: "adni_nifti_read_f32", torch.float64: "adni_nifti_read_f64",
          torch.uint8: "adni_nifti_read_mask_u8"}[out.dtype]
    _check(getattr(load(), fn)(os.fsencode(path), out.data_ptr(), out.numel(), ctypes.byref(info)))
    return out


def stage_volumes(paths, out, threads=8):
    """Decode len(paths) files in parallel into out[i] (out: contiguous CPU tensor (n, ...), fp32 = intensities, uint8 =
    masks; pinned memory makes the following H2D copy asynchronous).  A None path leaves out[i] untouched."""
    n = len(paths)
    assert out.device.type == "cpu" and out.is_contiguous() and out.shape[0] == n and out.dtype in (torch.float32, torch.uint8)
    arr = (ctypes.c_char_p * n)(*[None if p is None else os.fsencode(p) for p in paths])
    status = (ctypes.c_int * n)()
    per = out[0].numel() if n else 0
    rc = load().adni_stage_volumes(arr, n, 0 if out.dtype == torch.float32 else 1, out.data_ptr(),
                                   per * out.element_size(), per, int(threads), status)
    _check(rc)
    return out
