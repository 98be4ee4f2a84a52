"""ctypes binding of libwakeword_b200.so (include/wakeword_b200.h).  No CPU fallback: a missing
library or a non-sm_100 device raises."""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libwakeword_b200.so")

WW_CONV_SPLIT2, WW_CONV_FP32, WW_CONV_FP16 = 0, 1, 2
CONV_MODES = {"split2": WW_CONV_SPLIT2, "fp32": WW_CONV_FP32, "fp16": WW_CONV_FP16}

AUG_NORM_IN, AUG_SHIFT, AUG_SPEED, AUG_NOISE, AUG_GAIN, AUG_NORM_OUT = (1 << i for i in range(6))


class WWConfig(C.Structure):
    _fields_ = [("sample_rate", C.c_int32), ("n_samples", C.c_int32), ("n_fft", C.c_int32),
                ("win_length", C.c_int32), ("hop_length", C.c_int32), ("n_mels", C.c_int32),
                ("fmin", C.c_float), ("fmax", C.c_float), ("hidden_size", C.c_int32),
                ("num_layers", C.c_int32), ("num_classes", C.c_int32), ("threshold", C.c_float),
                ("conv_mode", C.c_int32), ("chunk_clips", C.c_int32)]


class WWPvoc(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("rate", "rs_orig", "rs_new", "crop_off")]


class WWAug(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("flags", "shift", "rs_orig", "rs_new", "crop_off",
                                          "noise_idx", "noise_off", "snr_db", "gain")]


EXPORTS = ["ww_abi_version", "ww_create", "ww_destroy", "ww_last_error", "ww_n_frames", "ww_set_weights",
           "ww_prepare_resample", "ww_augment", "ww_logmel", "ww_forward", "ww_score", "ww_score_stream",
           "ww_score_host", "ww_kernel_launches", "ww_conv_mode", "ww_normalize", "ww_profile", "ww_profile_read",
           "ww_train_backward", "ww_train_apply", "ww_train_step", "ww_train_reset", "ww_train_n_params",
           "ww_train_grad_buffer", "ww_train_param_range", "ww_get_weights",
           "ww_train_get_moments", "ww_train_set_moments", "ww_train_get_step", "ww_train_set_step",
           "ww_host_alloc", "ww_host_free", "ww_host_numa_node", "ww_time_stretch", "ww_add_gaussian_noise", "ww_set_threshold",
           "ww_augment_pcm16", "ww_logmel_pcm16", "ww_score_pcm16", "ww_score_stream_pcm16", "ww_score_host_pcm16"]

_lib = None
_lock = threading.Lock()


class WakewordB200Error(RuntimeError):
    pass


def load():
    """Load the shared library (building is __graft_entry__.build()'s job; we never fall back)."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise WakewordB200Error(
                f"{LIB_PATH} is missing: build it with `python -m wakeword_jupyterlab_b200.build` "
                "(there is no CPU fallback)")
        lib = C.CDLL(LIB_PATH)
        vp, i32, i64 = C.c_void_p, C.c_int, C.c_int64
        lib.ww_abi_version.restype = C.c_int
        lib.ww_create.argtypes = [C.POINTER(vp), i32, C.POINTER(WWConfig)]
        lib.ww_destroy.argtypes = [vp]
        lib.ww_destroy.restype = None
        lib.ww_last_error.argtypes = [vp]
        lib.ww_last_error.restype = C.c_char_p
        lib.ww_n_frames.argtypes = [vp]
        lib.ww_set_weights.argtypes = [vp, C.c_char_p, vp, C.POINTER(i64), i32]
        lib.ww_prepare_resample.argtypes = [vp, i32, i32]
        lib.ww_augment.argtypes = [vp, vp, vp, i32, i64, C.POINTER(WWAug), vp, i32, vp]
        lib.ww_logmel.argtypes = [vp, vp, i64, vp, i32, i32, vp]
        lib.ww_forward.argtypes = [vp, vp, vp, i32, vp]
        lib.ww_score.argtypes = [vp, vp, vp, i32, i64, C.POINTER(WWAug), i32, vp, vp, vp, i32, vp]
        lib.ww_score_stream.argtypes = [vp, vp, i64, i32, vp, vp, i64, vp]
        lib.ww_score_host.argtypes = [vp, vp, vp, i32, i64, C.POINTER(WWAug), i32, vp, vp, vp, i32]
        for base in ("ww_augment", "ww_logmel", "ww_score", "ww_score_stream", "ww_score_host"):
            getattr(lib, base + "_pcm16").argtypes = getattr(lib, base).argtypes
        f32 = C.c_float
        lib.ww_train_backward.argtypes = [vp, vp, vp, i32, vp, vp, vp, vp, vp]
        lib.ww_train_apply.argtypes = [vp, f32, f32, f32, f32, f32, f32, vp]
        lib.ww_train_step.argtypes = [vp, vp, vp, i32, vp, f32, vp, i32, vp]
        lib.ww_train_reset.argtypes = [vp]
        lib.ww_train_n_params.argtypes = [vp]
        lib.ww_train_n_params.restype = i64
        lib.ww_train_grad_buffer.argtypes = [vp]
        lib.ww_train_grad_buffer.restype = vp
        lib.ww_train_param_range.argtypes = [vp, C.c_char_p, C.POINTER(i64), C.POINTER(i64)]
        lib.ww_get_weights.argtypes = [vp, C.c_char_p, vp]
        lib.ww_train_get_moments.argtypes = [vp, C.c_char_p, vp, vp]
        lib.ww_train_set_moments.argtypes = [vp, C.c_char_p, vp, vp]
        lib.ww_train_get_step.argtypes = [vp]
        lib.ww_train_get_step.restype = i64
        lib.ww_train_set_step.argtypes = [vp, i64]
        lib.ww_time_stretch.argtypes = [vp, vp, C.POINTER(WWPvoc), vp, i32, vp]
        lib.ww_add_gaussian_noise.argtypes = [vp, vp, i64, C.c_float, C.c_uint64, vp]
        lib.ww_host_alloc.argtypes = [vp, C.c_size_t, C.POINTER(C.c_int)]
        lib.ww_host_alloc.restype = vp
        lib.ww_host_free.argtypes = [vp, vp]
        lib.ww_host_free.restype = None
        lib.ww_host_numa_node.argtypes = [vp]
        lib.ww_kernel_launches.argtypes = [vp]
        lib.ww_kernel_launches.restype = i64
        lib.ww_conv_mode.argtypes = [vp]
        lib.ww_set_threshold.argtypes = [vp, C.c_float]
        lib.ww_normalize.argtypes = [vp, vp, vp, i64, vp]
        lib.ww_profile.argtypes = [vp, i32]
        lib.ww_profile_read.argtypes = [vp, i32, C.POINTER(C.c_double), C.POINTER(i64)]
        _lib = lib
        return lib


def check(lib, ctx, rc, what):
    if rc != 0:
        msg = lib.ww_last_error(ctx)
        raise WakewordB200Error(f"{what} failed (code {rc}): {msg.decode() if msg else ''}")
