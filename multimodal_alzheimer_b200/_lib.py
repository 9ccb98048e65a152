"""ctypes binding of libadni_b200.so (the C-ABI declared in include/adni_b200.h).

There is NO fallback: if the shared library is missing or a call fails, an exception is raised.
``ADNI_ENOMEM`` is re-raised as ``torch.cuda.OutOfMemoryError`` (the reference's HPO loop catches that,
pkg/models/mri_models/train_anat_cnn.py:148-150); ``ADNI_EINVAL`` as ``ValueError``.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libadni_b200.so")

ENGINE_AUTO, ENGINE_TCGEN05, ENGINE_DIRECT, ENGINE_MMA_SYNC = 0, 1, 2, 3

_lib = None


class AdniError(RuntimeError):
    pass


class ConvGeom(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int) for n in ("N", "D", "H", "W", "Cin", "Cout", "k", "stride", "pad", "dil")]


_P = ctypes.c_void_p
_I = ctypes.c_int
_LL = ctypes.c_longlong
_D = ctypes.c_double
_F = ctypes.c_float
_SZ = ctypes.c_size_t

# name -> argtypes (every function returns int unless listed in _RESTYPES)
_SIGNATURES = {
    "adni_conv3d_out_extent": [_I, _I, _I, _I, _I],
    "adni_conv3d_plan_info": [ctypes.POINTER(ConvGeom), _I, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_double)],
    "adni_conv3d_wgrad_schedule": [ctypes.POINTER(ConvGeom), ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int), _I],
    "adni_conv3d_fprop": [ctypes.POINTER(ConvGeom), _P, _P, _P, _P, _P, _P, _I, _P],
    "adni_conv3d_dgrad": [ctypes.POINTER(ConvGeom), _P, _P, _P, _P, _I, _P],
    "adni_conv3d_wgrad": [ctypes.POINTER(ConvGeom), _P, _P, _P, _P, _P, _I, _P],
    "adni_conv3d_wgrad_scratch_floats": [ctypes.POINTER(ConvGeom)],
    "adni_stem_x8_elems": [_I, _I, _I, _I],
    "adni_stem_expand": [_P, _I, _I, _I, _I, _P, _P],
    "adni_stem_weights": [_P, _P, _P],
    "adni_stem_fprop": [_P, _I, _I, _I, _I, _P, _P, _P, _P, _P],
    "adni_stem_wgrad": [_P, _P, _I, _I, _I, _I, _P, _P, _P],
    "adni_weights_to_kernel_layout": [_P, _I, _I, _I, _P, _P, _P],
    "adni_wgrad_to_param_layout": [_P, _I, _I, _I, _P, _I, _P],
    "adni_peer_buffer_bytes": [_I, _I],
    "adni_peer_allreduce_f64": [_P, _I, _P, _P, _I, _I, _I, _P],
    "adni_peer_grad_flag_bytes": [],
    "adni_peer_allreduce_f32": [_LL, _LL, _P, _P, _P, _P, _I, _I, _I, _P],
    "adni_volumes_to_ndhwc_bf16": [_P, _I, _I, _I, _LL, _P, _P],
    "adni_maxout_fwd": [_P, _P, _P, _LL, _P],
    "adni_maxout_bwd": [_P, _P, _P, _P, _P, _LL, _P],
    "adni_concat_channels": [_P, _I, _P, _I, _LL, _P, _P],
    "adni_split_channels": [_P, _I, _I, _LL, _P, _P, _P],
    "adni_pad_volume_high": [_P, _I, _I, _I, _I, _I, _I, _I, _I, _P, _P],
    "adni_crop_volume_high": [_P, _I, _I, _I, _I, _I, _I, _I, _I, _P, _P],
    "adni_bootstrap_metrics": [_P, _I, _P, _P, _I, _I, _I, _P, _P, _P, _P, _P],
    "adni_adam_max_tensors_per_launch": [],
    "adni_adam_step_multi": [_I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _D, _D, _D, _P],
    "adni_weights_multi_job_bytes": [],
    "adni_weights_to_kernel_layout_multi": [_P, _I, _I, _P],
    "adni_bn_finalize": [_P, _P, _D, _I, _P, _P, _F, _F, _P, _P, _P, _P, _P, _P, _P],
    "adni_bn_apply": [_P, _P, _P, _P, _P, _LL, _I, _I, _P, _P, _P],
    "adni_bn_bwd_reduce": [_P, _P, _P, _P, _P, _P, _P, _LL, _I, _I, _P, _P],
    "adni_bn_bwd_apply": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _D, _LL, _I, _I, _P, _P, _P, _P, _D, _I, _P],
    "adni_conv3d_dgrad_bnred_profitable": [ctypes.POINTER(ConvGeom)],
    "adni_conv3d_dgrad_bnred": [ctypes.POINTER(ConvGeom), _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P],
    "adni_bn_train_apply": [_P, _P, _P, _D, _P, _P, _F, _F, _P, _P, _P, _P, _P, _LL, _I, _I, _P],
    "adni_channel_stats": [_P, _LL, _I, _P, _P, _P],
    "adni_maxpool3d_fwd": [_P, _I, _I, _I, _I, _I, _I, _I, _I, _P, _P, _P],
    "adni_maxpool3d_bwd": [_P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P, _P],
    "adni_relu_maxpool_fwd": [_P, _I, _I, _I, _I, _I, _I, _I, _P, _P, _P],
    "adni_relu_maxpool_bwd": [_P, _P, _P, _I, _I, _I, _I, _I, _I, _P, _P],
    "adni_bn_relu_maxpool_fwd": [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P],
    "adni_maxpool_bn_bwd_reduce": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P, _P],
    "adni_maxpool_bn_bwd_apply": [_P, _P, _P, _P, _P, _P, _D, _I, _I, _I, _I, _I, _I, _I, _I, _P, _P],
    "adni_gap_fwd": [_P, _I, _LL, _I, _P, _P],
    "adni_gap_bwd": [_P, _I, _LL, _I, _P, _P],
    "adni_linear_fwd": [_P, _I, _P, _P, _P, _I, _I, _I, _I, _I, _P],
    "adni_linear_bwd": [_P, _I, _P, _P, _I, _P, _I, _P, _I, _I, _P, _P, _I, _I, _I, _I, _P],
    "adni_rows_stats_f32": [_P, _I, _I, _I, _P, _P],
    "adni_bn1d_apply": [_P, _I, _P, _P, _P, _I, _I, _I, _I, _P],
    "adni_bn1d_bwd_reduce": [_P, _I, _P, _I, _P, _I, _P, _P, _I, _I, _I, _P, _P],
    "adni_bn1d_bwd_apply": [_P, _I, _P, _I, _P, _I, _P, _P, _P, _P, _D, _I, _I, _I, _P, _I, _P, _P, _P],
    "adni_loss_fwd": [_P, _I, _I, _P, _I, _I, _D, _P, _P, _P, _P],
    "adni_loss_bwd": [_P, _I, _I, _P, _I, _I, _P, _P, _P, _P, _I, _P],
    "adni_bn_param_grads": [_P, _I, _P, _P, _P],
    "adni_bn_eval_params": [_P, _P, _P, _P, _F, _I, _P, _P, _P],
    "adni_dropout": [_P, _P, _LL, _I, _D, ctypes.c_ulonglong, _P, _P],
    "adni_relu_fwd": [_P, _P, _LL, _P],
    "adni_relu_f32": [_P, _P, _P, _LL, _P],
    "adni_relu_bwd": [_P, _P, _P, _LL, _P],
    "adni_quantile_workspace_bytes": [_I],
    "adni_quantile_minmax_normalize": [_P, _P, _I, _LL, _D, _P, _P, _P, _P, _P, _SZ, _P],
    "adni_standardize": [_P, _P, _LL, _D, _D, _P, _P, _P],
    "adni_scan_moments": [_P, _I, _LL, _P, _P],
    "adni_masked_std_mean": [_P, _P, _I, _LL, _P, _P],
    "adni_cast_f32_to_bf16": [_P, _P, _LL, _P],
    "adni_cast_f64_to_bf16": [_P, _P, _LL, _P],
    "adni_cast_bf16_to_f32": [_P, _P, _LL, _P],
    "adni_last_error_string": [],
    "adni_version": [],
    "adni_launch_count": [],
}
_RESTYPES = {
    "adni_last_error_string": ctypes.c_char_p,
    "adni_launch_count": ctypes.c_longlong,
    "adni_stem_x8_elems": ctypes.c_longlong,
    "adni_quantile_workspace_bytes": ctypes.c_size_t,
    "adni_peer_buffer_bytes": ctypes.c_size_t,
    "adni_peer_grad_flag_bytes": ctypes.c_size_t,
    "adni_conv3d_wgrad_scratch_floats": ctypes.c_longlong,
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES.keys())


def load():
    """Load the shared library (raises if it has not been built — there is no CPU fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise AdniError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(multimodal_alzheimer_b200 has no CPU or PyTorch fallback path)")
    lib = ctypes.CDLL(LIB_PATH)
    for name, argtypes in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, ctypes.c_int)
    _lib = lib
    return lib


def launch_count():
    return int(load().adni_launch_count())


def _check(rc, what):
    if rc == 0:
        return
    msg = load().adni_last_error_string().decode("utf-8", "replace")
    if rc == -4:
        raise torch.cuda.OutOfMemoryError(f"{what}: {msg}")
    if rc == -1:
        raise ValueError(f"{what}: {msg}")
    if rc == -2:
        raise NotImplementedError(f"{what}: {msg}")
    raise AdniError(f"{what} failed ({rc}): {msg}")


def ptr(t):
    """Device pointer of a tensor (None -> NULL). The tensor must be a contiguous CUDA tensor."""
    if t is None:
        return None
    if not t.is_cuda:
        raise AdniError("adni_b200 kernels need CUDA tensors (no CPU fallback)")
    return ctypes.c_void_p(t.data_ptr())


def stream_ptr():
    # torch.cuda.current_stream() costs ~14 us per call (device-index plumbing); the raw getter is ~0.3 us
    return ctypes.c_void_p(torch._C._cuda_getCurrentRawStream(torch.cuda.current_device()))


# optional per-entry-point device timing (bench.py --shape-profile): name -> [n, list of (start, end) events]
CALL_TIMING = None


def call(name, *args):
    lib = load()
    if CALL_TIMING is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = getattr(lib, name)(*args)
        e1.record()
        CALL_TIMING.setdefault(name, []).append((e0, e1))
    else:
        rc = getattr(lib, name)(*args)
    _check(rc, name)


def collect_call_timing():
    """Returns {entry point: (calls, total ms)} and resets the recorder."""
    global CALL_TIMING
    rec, CALL_TIMING = CALL_TIMING, None
    torch.cuda.synchronize()
    return {k: (len(v), sum(a.elapsed_time(b) for a, b in v)) for k, v in (rec or {}).items()}


def geom(N, D, H, W, Cin, Cout, k, stride, pad, dil):
    return ConvGeom(N, D, H, W, Cin, Cout, k, stride, pad, dil)


def out_extent(n, k, stride, pad, dil):
    return (n + 2 * pad - dil * (k - 1) - 1) // stride + 1
