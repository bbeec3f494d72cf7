"""ctypes binding of libafsl.so - the only way the package reaches the GPU kernels.

There is deliberately no fallback: if the shared object is missing or a call
fails, an exception is raised (a product path that silently ran on the CPU or
in eager PyTorch would void every parity and performance claim).
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_double, c_float, c_int, c_longlong, c_void_p

import torch

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG_DIR, "libafsl.so")

_P, _I, _F, _D = c_void_p, c_int, c_float, c_double

# name -> argument ctypes, mirroring include/afsl.h exactly
SIGNATURES = {
    "afsl_prototypes_fwd_f32": [_P, _P, _P, _I, _I, _I, _I, _P],
    "afsl_prototypes_bwd_f32": [_P, _P, _P, _I, _I, _I, _I, _P],
    "afsl_proto_scores_fwd_f32": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P],
    "afsl_proto_scores_bwd_f32": [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P],
    "afsl_proto_head_fwd_f32": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P],
    "afsl_proto_head_bwd_f32": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P],
    "afsl_l2_normalize_fwd_f32": [_P, _P, _I, _I, _F, _P],
    "afsl_l2_normalize_bwd_f32": [_P, _P, _P, _I, _I, _F, _P],
    "afsl_cpl_fwd_f32": [_P, _P, _P, _P, _F, _P, _I, _I, _I, _I, _P],
    "afsl_cpl_bwd_f32": [_P, _P, _P, _P, _F, _P, _P, _P, _I, _I, _I, _I, _P],
    "afsl_cpl_fwd_save_f32": [_P, _P, _P, _P, _F, _P, _P, _P, _I, _I, _I, _I, _P],
    "afsl_cpl_bwd_saved_f32": [_P, _P, _P, _P, _F, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P],
    "afsl_angular_fwd_f32": [_P, _P, _P, _F, _F, _I, _I, _P, _I, _I, _I, _I, _P],
    "afsl_angular_bwd_f32": [_P, _P, _P, _F, _F, _I, _I, _P, _P, _P, _I, _I, _I, _I, _P],
    "afsl_specaug_views_f32": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _F, _I, _I, _I, _I, _I, _P],
    "afsl_view_fusion_grid": [_I, _I],
    "afsl_view_fusion_fwd_f32": [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P],
    "afsl_view_fusion_bwd_f32": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P],
    "afsl_gbn_stats_f32": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _F, _P],
    "afsl_gbn_relu_pool_fwd_f32": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "afsl_gbn_relu_pool_bwd_f32": [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "afsl_stage1_moments_f64": [_P, _P, _I, _I, _I, _I, _I, _P],
    "afsl_stage1_fwd_f32": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "afsl_stage1_bwd_f32": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    "afsl_gbn_stats_nhwc_f32": [_P, _P, _I, _P, _P, _P, _I, _I, _I, _I, _I, _F, _P],
    "afsl_gbn_relu_pool_nhwc_fwd_f32": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "afsl_gbn_relu_pool_nhwc_bwd_f32": [_P, _P, _P, _P, _P, _P, _P, _P, _I, _P, _I, _I, _I, _I, _I, _I, _P],
    "afsl_bn_running_update_f32": [_P, _P, _P, _P, _P, _P, _F, _F, _I, _I, _P],
    "afsl_stage1_finalize_f64": [_P, _I, _P, _P, _P, _F, _D, _P, _P, _P, _P, _P, _P, _P, _I, _P],
    "afsl_stage1_dw_f32": [_P, _I, _I, _P, _P, _P, _P, _P, _P, _D, _I, _P, _P, _P, _P, _P],
    "afsl_eval_vote_i32": [_P, _P, _P, _P, _P, _I, _P, _P, _I, _P],
    "afsl_logmel_f32": [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _F, _F, _F, _P],
    "afsl_transpose_f32": [_P, _P, _I, _I, _I, _P],
    "afsl_linear_fwd_f32": [_P, _P, _P, _P, _I, _I, _I, _I, _P],
    "afsl_linear_bwd_f32": [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _P],
}

_lib = None


class AfslError(RuntimeError):
    pass


def load() -> ctypes.CDLL:
    """Load libafsl.so (once).  Raises if it has not been built - never falls back."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise AfslError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(or `python audio-few-shot-learning_b200/build.py`); there is no CPU / eager fallback")
    lib = ctypes.CDLL(LIB_PATH)
    lib.afsl_version.restype = c_int
    lib.afsl_last_error.restype = c_char_p
    lib.afsl_launch_count.restype = c_longlong
    lib.afsl_view_fusion_weight_floats.restype = c_int
    lib.afsl_view_fusion_param_floats.restype = c_int
    lib.afsl_stage1_channels.restype = c_int
    lib.afsl_stage1_acc_slots.restype = c_int
    lib.afsl_gbn_nhwc_parts.restype = c_int
    lib.afsl_gbn_nhwc_parts.argtypes = [c_int]
    lib.afsl_linear_bwd_workspace_floats.restype = c_longlong
    lib.afsl_linear_bwd_workspace_floats.argtypes = [c_int, c_int, c_int]
    lib.afsl_cpl_saved_supported.restype = c_int
    lib.afsl_cpl_saved_supported.argtypes = [c_int, c_int, c_int]
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = c_int
    _lib = lib
    return lib


def launch_count() -> int:
    """Kernel launches issued through the library so far (bench bookkeeping)."""
    return int(load().afsl_launch_count())


# devices of the tensors whose pointers were taken for the launch being assembled (ptr() ... then call())
_pending_devices = []


class _CurrentStream:
    """Placeholder returned by stream_ptr(): call() replaces it with the current stream OF THE TENSORS' DEVICE."""


_STREAM = _CurrentStream()


def ptr(t, channels_last: bool = False):
    """Device pointer of a tensor or None; enforces the ABI's layout rules (``channels_last``: the tensor must be
    dense in NHWC order - the entry points that take it say so) and records the tensor's device for call()."""
    if t is None:
        return None
    try:
        if not t.is_cuda:
            raise AfslError("libafsl operates on CUDA tensors only (no CPU path); got a %s tensor" % t.device)
        if channels_last:
            if not t.is_contiguous(memory_format=torch.channels_last):
                raise AfslError("this libafsl entry point needs a channels-last (NHWC) tensor")
        elif not t.is_contiguous():
            raise AfslError("libafsl needs contiguous tensors")
        p = t.data_ptr()
        if p % 16 and t.numel():
            raise AfslError("libafsl needs 16-byte aligned tensors")
    except AfslError:
        _pending_devices.clear()
        raise
    _pending_devices.append(t.device.index)
    return p


def stream_ptr():
    """The launch stream argument: resolved by call() to the current stream of the device that owns the tensors
    (the reference picks ``cuda:{gpu_index}`` without ``set_device``, src/train_test.py:44-45, so the current
    device need not be the tensors' device)."""
    return _STREAM


def call(name: str, *args) -> None:
    """Launch one libafsl entry point on the device of its tensor arguments (all must share one device), on that
    device's current stream."""
    lib = load()
    devices = set(_pending_devices)
    _pending_devices.clear()
    if len(devices) > 1:
        raise AfslError(f"{name}: tensor arguments live on different devices {sorted(devices)}")
    dev = devices.pop() if devices else torch.cuda.current_device()
    fn = getattr(lib, name)
    if dev == torch.cuda.current_device():
        rc = fn(*[torch.cuda.current_stream(dev).cuda_stream if a is _STREAM else a for a in args])
    else:
        with torch.cuda.device(dev):              # grid sizing and the launch itself follow cudaGetDevice()
            rc = fn(*[torch.cuda.current_stream(dev).cuda_stream if a is _STREAM else a for a in args])
    if rc != 0:
        raise AfslError(f"{name} failed (code {rc}): {lib.afsl_last_error().decode()}")
