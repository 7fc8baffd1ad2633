"""ctypes binding of libdct3d.so (include/dct3d.h).  Loading fails loudly when the CUDA library
is missing: there is no CPU fallback in this package."""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

OK, E_INVALID, E_CUDA, E_OVERFLOW, E_STREAM, E_NEED_MORE = 0, -1, -2, -3, -4, -5

# name -> (restype, argtypes); every symbol include/dct3d.h declares
_vp, _u64p, _szp = C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_size_t)
SYMBOLS = {
    "dct3d_device_count": (C.c_int, []),
    "dct3d_list_devices": (C.c_int, [C.c_char_p, C.c_size_t]),
    "dct3d_create": (C.c_int, [C.POINTER(_vp), C.c_int, C.c_int, C.c_int, C.c_int]),
    "dct3d_destroy": (None, [_vp]),
    "dct3d_last_error": (C.c_char_p, [_vp]),
    "dct3d_eg_locate": (C.c_int, [_vp, _vp, C.c_size_t, C.c_uint64, C.c_size_t, C.POINTER(C.c_uint64)]),
    "dct3d_eg_locate_dev": (C.c_int, [_vp, _vp, C.c_size_t, C.c_uint64, C.c_size_t, C.POINTER(C.c_uint64), _vp]),
    "dct3d_rgb_split": (C.c_int, [_vp, _vp, C.c_size_t, _vp, _vp, _vp]),
    "dct3d_rgb_mix": (C.c_int, [_vp, _vp, _vp, _vp, C.c_size_t, _vp]),
    "dct3d_rgb_split_dev": (C.c_int, [_vp, _vp, C.c_size_t, _vp, _vp, _vp, _vp]),
    "dct3d_rgb_mix_dev": (C.c_int, [_vp, _vp, _vp, _vp, C.c_size_t, _vp, _vp]),
    "dct3d_encode_u8_range": (C.c_int, [_vp, _vp, C.c_int, _u64p]),
    "dct3d_encode_u8_place": (C.c_int, [_vp, C.c_uint64, C.c_int, _vp, C.c_size_t, C.POINTER(C.c_uint8)]),
    "dct3d_decode_u8_range": (C.c_int, [_vp, _vp, C.c_size_t, C.c_uint64, C.c_uint64, C.c_int, _vp, _u64p]),
    "dct3d_stream_shift_dev": (C.c_int, [_vp, _vp, C.c_uint64, C.c_uint, _vp, C.c_size_t, _vp]),
    "dct3d_host_register": (C.c_int, [_vp, C.c_size_t]),
    "dct3d_host_unregister": (C.c_int, [_vp]),
    "dct3d_multi_create": (C.c_int, [C.POINTER(_vp), C.POINTER(C.c_int), C.c_int, C.c_int, C.c_int, C.c_int]),
    "dct3d_multi_destroy": (None, [_vp]),
    "dct3d_multi_last_error": (C.c_char_p, [_vp]),
    "dct3d_multi_set_option": (C.c_int, [_vp, C.c_char_p, C.c_long]),
    "dct3d_multi_context": (C.c_void_p, [_vp, C.c_int]),
    "dct3d_multi_encode_u8": (C.c_int, [_vp, _vp, C.c_int, _vp, C.c_size_t, _u64p, _szp, _u64p]),
    "dct3d_multi_locate": (C.c_int, [_vp, _vp, C.c_size_t, C.c_int, _u64p]),
    "dct3d_multi_decode_u8": (C.c_int, [_vp, _vp, C.c_size_t, C.c_int, _vp, _u64p]),
    "dct3d_multi_set_weights": (C.c_int, [_vp, C.POINTER(C.c_double)]),
    "dct3d_multi_probe_links": (C.c_int, [_vp, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "dct3d_multi_stream_begin": (C.c_int, [_vp]),
    "dct3d_multi_stream_encode": (C.c_int, [_vp, _vp, C.c_int, C.c_int, _vp, C.c_size_t, _szp]),
    "dct3d_multi_stream_decode": (C.c_int, [_vp, _vp, C.c_size_t, _u64p, C.c_int, _vp]),
    "dct3d_host_alloc": (C.c_void_p, [C.c_size_t]),
    "dct3d_host_free": (None, [_vp]),
    "dct3d_set_option": (C.c_int, [_vp, C.c_char_p, C.c_long]),
    "dct3d_get_stat": (C.c_long, [_vp, C.c_char_p]),
    "dct3d_encode_u8": (C.c_int, [_vp, _vp, C.c_int, _vp, C.c_size_t, _u64p, _szp]),
    "dct3d_decode_u8": (C.c_int, [_vp, _vp, C.c_size_t, C.c_int, _vp]),
    "dct3d_stream_begin": (C.c_int, [_vp]),
    "dct3d_stream_encode": (C.c_int, [_vp, _vp, C.c_int, C.c_int, _vp, C.c_size_t, _szp]),
    "dct3d_stream_decode": (C.c_int, [_vp, _vp, C.c_size_t, _u64p, C.c_int, _vp]),
    "dct3d_forward_f32": (C.c_int, [_vp, _vp, _vp, C.c_int]),
    "dct3d_inverse_f32": (C.c_int, [_vp, _vp, _vp, C.c_int]),
    "dct3d_forward_f64": (C.c_int, [_vp, _vp, _vp, C.c_int]),
    "dct3d_inverse_f64": (C.c_int, [_vp, _vp, _vp, C.c_int]),
    "dct3d_quantize_u8": (C.c_int, [_vp, _vp, C.c_int, _vp]),
    "dct3d_reconstruct_i16": (C.c_int, [_vp, _vp, C.c_int, _vp]),
    "dct3d_eg_encode_i16": (C.c_int, [_vp, _vp, C.c_size_t, C.c_uint64, _vp, C.c_size_t, _u64p]),
    "dct3d_eg_decode_i16": (C.c_int, [_vp, _vp, C.c_size_t, C.c_uint64, C.c_size_t, _vp, _u64p]),
    "dct3d_encode_u8_dev": (C.c_int, [_vp, _vp, C.c_int, _vp, C.c_size_t, C.c_uint64, _u64p, _vp]),
    "dct3d_decode_u8_dev": (C.c_int, [_vp, _vp, C.c_size_t, C.c_uint64, C.c_int, _vp, _u64p, _vp]),
    "dct3d_forward_f32_dev": (C.c_int, [_vp, _vp, _vp, C.c_int, _vp]),
    "dct3d_inverse_f32_dev": (C.c_int, [_vp, _vp, _vp, C.c_int, _vp]),
    "dct3d_forward_f64_dev": (C.c_int, [_vp, _vp, _vp, C.c_int, _vp]),
    "dct3d_inverse_f64_dev": (C.c_int, [_vp, _vp, _vp, C.c_int, _vp]),
    "dct3d_quantize_u8_dev": (C.c_int, [_vp, _vp, C.c_int, _vp, _vp]),
    "dct3d_reconstruct_i16_dev": (C.c_int, [_vp, _vp, C.c_int, _vp, _vp]),
    "dct3d_eg_encode_i16_dev": (C.c_int, [_vp, _vp, C.c_size_t, C.c_uint64, _vp, C.c_size_t, _u64p, _vp]),
    "dct3d_eg_decode_i16_dev": (C.c_int, [_vp, _vp, C.c_size_t, C.c_uint64, C.c_size_t, _vp, _u64p, _vp]),
}

_LIB = None


def lib_path() -> str:
    """libdct3d.so built in-tree; DCT3D_LIB names another build of the same sources (kernel-variant experiments of
    profiles/tools/step_time.py).  Either way it is this CUDA library or nothing."""
    return os.environ.get("DCT3D_LIB") or _build.LIB


def load() -> C.CDLL:
    """dlopen libdct3d.so (built in-tree by build.build()) and bind every entry point."""
    global _LIB
    if _LIB is None:
        path = lib_path()
        if path == _build.LIB and _build.stale() and os.path.exists(os.path.join(_build.CSRC, "dct3d_api.cu")):
            try:
                _build.build()
            except Exception:   # noqa: BLE001  (no nvcc on this box: the shipped library is used as it is)
                pass
        if not os.path.exists(path):
            raise RuntimeError(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(nvcc, sm_100a). This package has no CPU fallback.")
        L = C.CDLL(path)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)   # AttributeError if the library does not export it
            fn.restype = res
            fn.argtypes = args
        _LIB = L
    return _LIB
