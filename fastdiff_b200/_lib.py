"""ctypes binding of the C-ABI in include/fastdiff_b200.h (no torch types cross this boundary)."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_LIB = os.path.join(_HERE, "csrc", "libfastdiff_b200.so")

FD_MODE_FP32_SIMT, FD_MODE_TC_3XTF32, FD_MODE_TC_TF32, FD_MODE_TC_3XF16 = 0, 1, 2, 3
MODE_NAMES = {"fp32_simt": 0, "tc_3xtf32": 1, "tc_tf32": 2, "tc_3xf16": 3}


class fd_config(C.Structure):
    _fields_ = [
        ("audio_channels", C.c_int32), ("inner_channels", C.c_int32), ("cond_channels", C.c_int32),
        ("n_upsample", C.c_int32), ("upsample_ratios", C.c_int32 * 4),
        ("lvc_layers_each_block", C.c_int32), ("lvc_kernel_size", C.c_int32),
        ("kpnet_hidden_channels", C.c_int32), ("kpnet_conv_size", C.c_int32),
        ("diffusion_step_embed_dim_in", C.c_int32), ("diffusion_step_embed_dim_mid", C.c_int32),
        ("diffusion_step_embed_dim_out", C.c_int32),
    ]


class fd_step(C.Structure):
    _fields_ = [
        ("t", C.c_float), ("coef_eps", C.c_float), ("div", C.c_float), ("sigma", C.c_float),
        ("c1", C.c_float), ("c2", C.c_float), ("c3", C.c_float), ("add_noise", C.c_int32),
    ]


# every symbol include/fastdiff_b200.h declares: (restype, argtypes)
_P = C.c_void_p
SYMBOLS = {
    "fd_create": (C.c_int, [C.POINTER(fd_config), C.c_int, C.POINTER(_P)]),
    "fd_load_weights": (C.c_int, [_P, _P, C.c_size_t]),
    "fd_load_weights_dev": (C.c_int, [_P, _P, C.c_size_t, _P]),
    "fd_workspace_bytes": (C.c_int, [_P, C.c_int, C.c_int, C.POINTER(C.c_size_t)]),
    "fd_set_mode": (C.c_int, [_P, C.c_int]),
    "fd_get_mode": (C.c_int, [_P]),
    "fd_set_option": (C.c_int, [_P, C.c_char_p, C.c_int64]),
    "fd_denoise": (C.c_int, [_P, _P, _P, _P, _P, C.c_int, C.c_int, _P, C.c_size_t, _P]),
    "fd_sample": (C.c_int, [_P, _P, _P, C.POINTER(fd_step), C.c_int, _P, C.c_int, C.c_uint64, C.c_int, C.c_int, _P,
                            C.c_int, C.c_int, _P, C.c_size_t, _P]),
    "fd_reverse_update": (C.c_int, [_P, _P, _P, _P, C.POINTER(fd_step), C.c_int, C.c_uint64, C.c_uint32, _P, C.c_size_t, _P]),
    "fd_wav_int16": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, _P, _P]),
    "fd_mel_frontend": (C.c_int, [_P, _P, C.c_int, C.c_int, _P, _P, _P, _P]),
    "fd_debug_read": (C.c_int, [_P, C.c_char_p, _P, C.POINTER(C.c_size_t), C.c_int, C.c_int, _P, _P]),
    "fd_set_noise_window": (C.c_int, [_P, C.c_int64, C.c_int64]),
    "fd_launch_count": (C.c_uint64, [_P]),
    "fd_check_saturation": (C.c_int, [_P, C.POINTER(C.c_int), C.c_int, _P]),
    "fd_timing_enable": (C.c_int, [_P, C.c_int]),
    "fd_timing_report": (C.c_int, [_P, C.c_char_p, C.c_size_t]),
    "fd_last_error": (C.c_char_p, [_P]),
    "fd_destroy": (None, [_P]),
    "fd_version": (C.c_char_p, []),
}

_cache = {}


def load(path: str | None = None) -> C.CDLL:
    """Load the shared library and type its entry points.  Fails loudly when the extension is absent:
    there is no CPU fallback for the product path."""
    path = os.path.abspath(path or os.environ.get("FASTDIFF_B200_LIB", DEFAULT_LIB))
    if path in _cache:
        return _cache[path]
    if not os.path.exists(path):
        raise RuntimeError(
            f"fastdiff_b200: CUDA extension not built ({path} missing). Run `python -c 'import __graft_entry__ as g; g.build()'` "
            "at the repo root (nvcc, sm_100a). There is no CPU fallback.")
    lib = C.CDLL(path)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _cache[path] = lib
    return lib


class FdError(RuntimeError):
    pass


def check(lib, handle, rc: int, what: str):
    if rc != 0:
        msg = lib.fd_last_error(handle)
        raise FdError(f"{what} failed ({rc}): {msg.decode() if msg else '?'}")
