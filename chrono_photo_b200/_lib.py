"""ctypes binding of libchrono_b200.so (include/chrono_b200.h). There is no Python or CPU fallback: if the shared
library is missing or no CUDA device is present, the calls fail loudly."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CHB_LIB", os.path.join(HERE, "libchrono_b200.so"))  # CHB_LIB: tuning builds


class ChbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libchrono_b200 error {code}: {msg}")
        self.code = code


class Fade(C.Structure):
    _fields_ = [("is_none", C.c_uint8), ("mode", C.c_uint8), ("absolute", C.c_uint8), ("_pad", C.c_uint8),
                ("offset", C.c_int32), ("n_values", C.c_int32), ("values", C.POINTER(C.c_float))]


class OutlierParams(C.Structure):
    _fields_ = [("thr_absolute", C.c_uint8), ("background", C.c_uint8), ("outlier", C.c_uint8), ("_pad", C.c_uint8),
                ("thr_min", C.c_float), ("thr_max", C.c_float), ("thr_scale", C.c_float),
                ("weights", C.c_float * 4), ("fade", Fade), ("sample_count", C.c_int32),
                ("seed", C.c_uint64), ("pixel_offset", C.c_uint64), ("block_pixels", C.c_uint64), ("block_skip", C.c_uint64)]


class SimpleParams(C.Structure):
    _fields_ = [("darker", C.c_uint8), ("_pad", C.c_uint8 * 3), ("weights", C.c_float * 4), ("fade", Fade)]


class DebugPlanes(C.Structure):
    _fields_ = [("median", C.POINTER(C.c_float)), ("q1", C.POINTER(C.c_float)), ("q3", C.POINTER(C.c_float)),
                ("n_outliers", C.POINTER(C.c_int32))]


# every symbol include/chrono_b200.h declares: name -> (restype, argtypes)
_vp, _i, _u8p, _i32p, _f32p, _u64p = C.c_void_p, C.c_int, C.POINTER(C.c_uint8), C.POINTER(C.c_int32), C.POINTER(C.c_float), C.POINTER(C.c_uint64)
SYMBOLS = {
    "chb_last_error": (C.c_char_p, []),
    "chb_version": (_i, []),
    "chb_ctx_create": (_i, [C.POINTER(C.c_int), _i, C.POINTER(_vp)]),
    "chb_ctx_destroy": (_i, [_vp]),
    "chb_ctx_device_count": (_i, [_vp]),
    "chb_ctx_mem_info": (_i, [_vp, _i, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "chb_ctx_set_stream": (_i, [_vp, _i, _vp]),
    "chb_stack_create": (_i, [_vp, _i, _i, _i, _i, C.POINTER(_vp)]),
    "chb_stack_destroy": (_i, [_vp]),
    "chb_stack_device_bytes": (C.c_size_t, [_vp, _i]),
    "chb_stack_upload": (_i, [_vp, _i, _vp, C.c_size_t, _i, _i]),
    "chb_stack_upload_pinned": (_i, [_vp, _i, _vp, C.c_size_t, _i, _i]),
    "chb_stack_download": (_i, [_vp, _i, _vp, C.c_size_t]),
    "chb_stack_sync": (_i, [_vp]),
    "chb_stack_fill_synthetic": (_i, [_vp, _i, C.c_uint64, _i, _i]),
    "chb_stack_fill_synthetic_blocks": (_i, [_vp, _i, C.c_uint64, _i, _i, _i, _i]),
    "chb_synth_frame_host": (_i, [_i, C.c_uint64, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "chb_outlier": (_i, [_vp, C.POINTER(OutlierParams), _i32p, _i, _vp, _vp, _u64p]),
    "chb_outlier_debug": (_i, [_vp, C.POINTER(OutlierParams), _i32p, _i, _vp, _vp, _u64p, C.POINTER(DebugPlanes)]),
    "chb_simple": (_i, [_vp, C.POINTER(SimpleParams), _i32p, _i, _vp]),
    "chb_outlier_device": (_i, [_vp, C.POINTER(OutlierParams), _i32p, _i, _i, _f32p]),
    "chb_simple_device": (_i, [_vp, C.POINTER(SimpleParams), _i32p, _i, _f32p]),
    "chb_fetch_last": (_i, [_vp, _vp, _vp, _u64p]),
    "chb_fetch_last_device": (_i, [_vp, _i, _vp, _vp]),
    "chb_set_tuning": (_i, [C.c_char_p, _i]),
    "chb_stack_upload_jpeg": (_i, [_vp, _i, _vp, C.c_size_t, _i, _i]),
    "chb_encode_jpeg": (_i, [_vp, _vp, _i, _i, C.c_size_t, _i, _vp, C.c_size_t, C.POINTER(C.c_size_t)]),
    "chb_decode_jpeg": (_i, [_vp, _vp, C.c_size_t, _vp, C.c_size_t, C.c_size_t, C.POINTER(_i), C.POINTER(_i)]),
    "chb_outlier_enqueue": (_i, [_vp, C.POINTER(OutlierParams), _i32p, _i, _i]),
    "chb_stack_wait": (_i, [_vp, _f32p, _u64p]),
    "chb_outlier_video": (_i, [_vp, C.POINTER(OutlierParams), _i, _i, _i, _vp, _vp, _u64p]),
    "chb_outlier_video_device": (_i, [_vp, C.POINTER(OutlierParams), _i, _i, _i, _i, _f32p]),
    "chb_launch_count": (C.c_uint64, []),
    "chb_launch_count_reset": (None, []),
    "chb_last_slow_pixels": (C.c_uint64, []),
    "chb_last_hard_pixels": (C.c_uint64, []),
    "chb_last_main_kernel_ms": (C.c_float, []),
    "chb_sample_positions": (_i, [C.c_uint64, _i, _i, _i32p]),
    "chb_threshold_new": (None, [_i, C.c_float, C.c_float, _f32p, _f32p, _f32p]),
    "chb_fade_build": (_i, [_i32p, _f32p, _i, _f32p, _i, _i32p]),
    "chb_crop_create": (_i, [_i32p, _i, _i, _i, _i32p, _i32p, _i32p]),
    "chb_shake_create": (_i, [_vp, _i, _i, _i, _i32p, _i, _i, _i, _vp, C.c_size_t, C.POINTER(_vp)]),
    "chb_shake_offset": (_i, [_vp, _vp, C.c_size_t, _i32p, _i32p, _i32p]),
    "chb_shake_destroy": (_i, [_vp]),
    "chb_video_windows": (_i, [_i] * 11 + [_i32p, _i32p, _i32p, _i]),
}

_lib = None


def lib():
    """Loads the shared library (once). Raises if it has not been built: the CUDA path is the only path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: run `python -m chrono_photo_b200.build` (nvcc, sm_100a). "
                              "chrono_photo_b200 has no CPU fallback.")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(rc):
    if rc != 0:
        raise ChbError(rc, lib().chb_last_error().decode("utf-8", "replace"))
