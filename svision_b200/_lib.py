"""ctypes binding of ``libsvx.so`` (``include/svx.h``).  There is no CPU fallback: if the
library is missing or a call fails, this module raises."""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsvx.so")

IMAGE_F32, IMAGE_F16 = 0, 1
PRECISION_3PASS, PRECISION_1PASS = 0, 1

#: every symbol include/svx.h declares (tests check the .so exports exactly these)
SYMBOLS = (
    "svx_create", "svx_destroy", "svx_encode", "svx_forward", "svx_classify_device",
    "svx_classify", "svx_debug_activation", "svx_gemm_selftest", "svx_conv_selftest", "svx_debug_counters", "svx_set_profiling",
    "svx_profile_read", "svx_launch_count",
    "svx_launch_count_reset", "svx_max_batch", "svx_device", "svx_last_error", "svx_version",
    "svx_bed_count_rows", "svx_bed_parse", "svx_pairs_generate",
    "svx_classify_device_calls", "svx_exchange_create", "svx_exchange_export", "svx_exchange_attach",
    "svx_classify_exchange", "svx_exchange_status", "svx_exchange_destroy",
    "svx_calls_aggregate", "svx_np_mean_f32", "svx_np_std_i64",
    "svx_multi_create", "svx_multi_classify", "svx_multi_device_count", "svx_multi_last_split",
    "svx_multi_destroy", "svx_selftest_last_ms",
)
IPC_HANDLE_BYTES = 72


class SvxWeights(ctypes.Structure):
    _fields_ = [(f"{l}_{k}", ctypes.c_void_p)
                for l in ("conv1", "conv2", "conv3", "conv4", "conv5", "fc6", "fc7", "fc8")
                for k in ("w", "b")]


class SvxError(RuntimeError):
    pass


_lib = None


def load() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SvxError(f"{LIB_PATH} not found: build it with `python -m svision_b200.build` "
                       "(there is no CPU fallback for the encode+classify path)")
    lib = ctypes.CDLL(LIB_PATH)
    vp, i32, i64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64
    lib.svx_create.argtypes = [ctypes.POINTER(SvxWeights), i32, i64, i32, ctypes.POINTER(vp)]
    lib.svx_create.restype = i32
    lib.svx_destroy.argtypes = [vp]
    lib.svx_destroy.restype = None
    lib.svx_encode.argtypes = [vp, vp, i64, vp, i32, vp]
    lib.svx_encode.restype = i32
    lib.svx_forward.argtypes = [vp, vp, i32, i64, vp, vp]
    lib.svx_forward.restype = i32
    lib.svx_classify_device.argtypes = [vp, vp, i64, vp, vp, vp, vp]
    lib.svx_classify_device.restype = i32
    lib.svx_classify.argtypes = [vp, vp, i64, vp, vp]
    lib.svx_classify.restype = i32
    lib.svx_debug_activation.argtypes = [vp, ctypes.c_char_p, i64, vp]
    lib.svx_debug_activation.restype = i32
    lib.svx_gemm_selftest.argtypes = [i32, vp, vp, vp, i64, i64, i64, i32, i32, vp]
    lib.svx_gemm_selftest.restype = i32
    lib.svx_conv_selftest.argtypes = [i32, vp, vp, vp, i64, i64, i64, i32, vp, i32, i32, vp]
    lib.svx_conv_selftest.restype = i32
    lib.svx_debug_counters.argtypes = [vp, vp, i32]
    lib.svx_debug_counters.restype = i32
    lib.svx_set_profiling.argtypes = [vp, i32]
    lib.svx_set_profiling.restype = i32
    lib.svx_profile_read.argtypes = [vp, vp, vp, i32]
    lib.svx_profile_read.restype = i32
    lib.svx_launch_count.argtypes = []
    lib.svx_launch_count.restype = i64
    lib.svx_launch_count_reset.argtypes = []
    lib.svx_launch_count_reset.restype = None
    lib.svx_bed_count_rows.argtypes = [vp, i64, ctypes.POINTER(i64)]
    lib.svx_bed_count_rows.restype = i32
    lib.svx_bed_parse.argtypes = [vp, i64, i64, vp, vp, vp, vp]
    lib.svx_bed_parse.restype = i32
    lib.svx_pairs_generate.argtypes = [i64, vp, vp, vp, i64, vp, vp, ctypes.POINTER(i64)]
    lib.svx_pairs_generate.restype = i32
    lib.svx_classify_device_calls.argtypes = [vp, vp, i64, vp, vp]
    lib.svx_classify_device_calls.restype = i32
    lib.svx_exchange_create.argtypes = [vp, i32, i32, i64, ctypes.POINTER(vp)]
    lib.svx_exchange_create.restype = i32
    lib.svx_exchange_export.argtypes = [vp, vp]
    lib.svx_exchange_export.restype = i32
    lib.svx_exchange_attach.argtypes = [vp, vp]
    lib.svx_exchange_attach.restype = i32
    lib.svx_classify_exchange.argtypes = [vp, vp, vp, i64, ctypes.POINTER(vp), vp]
    lib.svx_classify_exchange.restype = i32
    lib.svx_exchange_status.argtypes = [vp]
    lib.svx_exchange_status.restype = i32
    lib.svx_exchange_destroy.argtypes = [vp]
    lib.svx_exchange_destroy.restype = None
    lib.svx_calls_aggregate.argtypes = [vp, i64, i64, vp, vp, vp, vp, vp, vp, vp, i64, vp, vp, vp,
                                        ctypes.POINTER(i64), ctypes.POINTER(i64)]
    lib.svx_calls_aggregate.restype = i32
    lib.svx_np_mean_f32.argtypes = [vp, i64, vp]
    lib.svx_np_mean_f32.restype = i32
    lib.svx_np_std_i64.argtypes = [vp, i64, vp]
    lib.svx_np_std_i64.restype = i32
    lib.svx_multi_create.argtypes = [ctypes.POINTER(SvxWeights), ctypes.POINTER(i32), i32, i64, i32,
                                     ctypes.POINTER(vp)]
    lib.svx_multi_create.restype = i32
    lib.svx_multi_classify.argtypes = [vp, vp, i64, vp, vp]
    lib.svx_multi_classify.restype = i32
    lib.svx_multi_device_count.argtypes = [vp]
    lib.svx_multi_device_count.restype = i32
    lib.svx_multi_last_split.argtypes = [vp, vp]
    lib.svx_multi_last_split.restype = i32
    lib.svx_multi_destroy.argtypes = [vp]
    lib.svx_multi_destroy.restype = None
    lib.svx_selftest_last_ms.argtypes = []
    lib.svx_selftest_last_ms.restype = ctypes.c_float
    lib.svx_max_batch.argtypes = [vp]
    lib.svx_max_batch.restype = i64
    lib.svx_device.argtypes = [vp]
    lib.svx_device.restype = i32
    lib.svx_last_error.argtypes = []
    lib.svx_last_error.restype = ctypes.c_char_p
    lib.svx_version.argtypes = []
    lib.svx_version.restype = ctypes.c_char_p
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().svx_last_error().decode(errors="replace")
        raise SvxError(f"{what} failed (status {rc}): {msg}")
