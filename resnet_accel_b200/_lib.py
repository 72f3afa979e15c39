"""ctypes binding of libaccel_b200.so (the C ABI declared in include/accel_b200.h).

The library is the only compute path of this package: if it is missing or fails to load the
import raises - there is no CPU or PyTorch fallback.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ACCEL_B200_LIB") or os.path.join(_HERE, "libaccel_b200.so")      # override: developer builds

# status codes (accel_status; mirrors AcceleratorError::Code, accelerator_driver.hpp:337-345)
OK, INIT_FAILED, TIMEOUT, DMA_ERROR, ILLEGAL_COMMAND, INVALID_CONFIG, MEMORY_ERROR, NOT_READY = 0, -1, -2, -3, -4, -5, -6, -7
STATUS_NAMES = {0: "OK", -1: "INIT_FAILED", -2: "TIMEOUT", -3: "DMA_ERROR", -4: "ILLEGAL_COMMAND",
                -5: "INVALID_CONFIG", -6: "MEMORY_ERROR", -7: "NOT_READY"}

RELU, OUT_I8, OUT_I32, OUT_F32, RELU_OUT = 1, 2, 4, 8, 16


class AcceleratorError(RuntimeError):
    """Python twin of resnet_accel::AcceleratorError (accelerator_driver.hpp:330-360)."""

    def __init__(self, code: int, message: str):
        super().__init__(f"[{STATUS_NAMES.get(code, code)}] {message}")
        self.code = code


class OutLayout(C.Structure):
    _fields_ = [("rows_per_image", C.c_int64), ("image_stride", C.c_int64), ("chan_stride", C.c_int64),
                ("row_stride", C.c_int64), ("row_len", C.c_int64), ("row_pitch", C.c_int64)]


class Epilogue(C.Structure):
    _fields_ = [("flags", C.c_int32), ("n_channels", C.c_int32), ("chan_scale", C.c_void_p), ("bias", C.c_void_p),
                ("residual", C.c_void_p), ("res_scale_main", C.c_float), ("res_scale_res", C.c_float),
                ("res_scale_out", C.c_float), ("sat_count", C.c_void_p), ("chan_absmax", C.c_void_p), ("acc_bound", C.c_int32)]


class ConvGeom(C.Structure):
    _fields_ = [("batch", C.c_int32), ("c_in", C.c_int32), ("h", C.c_int32), ("w", C.c_int32),
                ("ksize", C.c_int32), ("stride", C.c_int32), ("pad", C.c_int32), ("in_row_pitch", C.c_int32)]


# every symbol include/accel_b200.h declares: name -> (restype, argtypes)
_P, _I32, _I64, _F, _SZ = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_size_t
SYMBOLS = {
    "accel_last_error_string": (C.c_char_p, []),
    "accel_version": (C.c_char_p, []),
    "accel_device_check": (C.c_int, []),
    "accel_debug_set_timeline": (None, [_P]),
    "accel_debug_counter": (C.c_longlong, [C.c_int]),
    "accel_plan_create": (C.c_int, [_P, _P, _I32, _I32, _I32, _I32, C.POINTER(_P), C.POINTER(_SZ)]),
    "accel_plan_upload": (C.c_int, [_P, _P, _P, _SZ, _P]),
    "accel_plan_destroy": (None, [_P]),
    "accel_plan_num_blocks": (_I64, [_P]),
    "accel_plan_num_mma": (_I64, [_P]),
    "accel_plan_num_tiles": (_I64, [_P]),
    "accel_plan_export_ops": (_I64, [_P, _P, _I64]),
    "accel_plan_export_mma": (_I64, [_P, _P, _I64]),
    "accel_plan_conv_ws_bytes": (C.c_int, [_P, _I32, _I32, _I32, C.POINTER(_SZ)]),
    "accel_plan_conv_ws_prepare": (C.c_int, [_P, _P, _I32, _I32, _I32, _P, _SZ, _P]),
    "accel_plan_conv_ws_release": (None, [_P]),
    "accel_plan_gemm_ws_bytes": (C.c_int, [_P, C.POINTER(_SZ)]),
    "accel_plan_gemm_ws_prepare": (C.c_int, [_P, _P, _P, _SZ, _P]),
    "accel_plan_gemm_ws_release": (None, [_P]),
    "accel_plan_gemm_ws_live_chunks": (_I64, [_P, _I32]),
    "accel_bsr_gemm_i8": (C.c_int, [_P, _P, _I64, _I64, _I64, C.POINTER(Epilogue), _P, C.POINTER(OutLayout), _P]),
    "accel_conv_bsr_i8": (C.c_int, [_P, _P, C.POINTER(ConvGeom), C.POINTER(Epilogue), _P, C.POINTER(OutLayout), _P]),
    "accel_conv_bsr_i8_dual": (C.c_int, [_P, _P, _P, C.POINTER(ConvGeom), C.POINTER(Epilogue), _P, C.POINTER(Epilogue), _P,
                                         C.POINTER(OutLayout), _P]),
    "accel_conv_pool_bsr_i8": (C.c_int, [_P, _P, C.POINTER(ConvGeom), C.POINTER(Epilogue), _I32, _I32, _I32, _P, _I32, _P]),
    "accel_bsr_gemm_generic": (C.c_int, [_P, _I64, _I64, _I64, _P, _P, _P, _I32, _I32, _I32, _I32, _I64, _P, _I64, _P]),
    "accel_block_l1_i8": (C.c_int, [_P, _I64, _I64, _I64, _I32, _P, _P]),
    "accel_block_l2_f32": (C.c_int, [_P, _I64, _I64, _I64, _I32, _I32, _P, _P]),
    "accel_bsr_scan": (C.c_int, [_P, _I32, _I32, _P, _P, _P]),
    "accel_bsr_gather_i8": (C.c_int, [_P, _I64, _I64, _I64, _I32, _P, _I32, _I32, _P, _P, _P]),
    "accel_bsr_gather_f32": (C.c_int, [_P, _I64, _I64, _I64, _I32, _I32, _P, _I32, _I32, _P, _P, _P]),
    "accel_quantize_rows_f32": (C.c_int, [_P, _I64, _I64, _I64, _P, _P, _P]),
    "accel_symmetric_scales_f32": (C.c_int, [_P, _I64, _P, _P]),
    "accel_row_absmax_f32": (C.c_int, [_P, _I64, _I64, _I64, _P, _P]),
    "accel_requant_i32_i8": (C.c_int, [_P, _P, _I64, _I64, _I64, _P, _P, _I32, _P, _P]),
    "accel_add_residual_i8": (C.c_int, [_P, _P, _P, _I64, _F, _F, _F, _P]),
    "accel_subsample2_i8": (C.c_int, [_P, _I64, C.c_int32, C.c_int32, C.c_int32, _P, C.c_int32, _P]),
    "accel_relu_i8": (C.c_int, [_P, _I64, _P]),
    "accel_relu6_i8": (C.c_int, [_P, _I64, _F, _P]),
    "accel_relu_i32": (C.c_int, [_P, _I64, _P]),
    "accel_gemm_bsr_int8_fp32": (C.c_int, [_P, _I64, _I64, _P, _P, _P, _P, _I32, _P, _I64, _I32, _I32, _I64, _I64, C.c_double, _I32,
                                           _P, _I32, _I32, _P, _P, _I64, _P]),
    "accel_maxpool_i8": (C.c_int, [_P, _P, _I64, _I32, _I32, _I32, _I32, _I32, _I32, _I32, _P]),
    "accel_avgpool_i8": (C.c_int, [_P, _P, _I64, _I32, _I32, _I32, _P]),
}

_lib = None


def lib() -> C.CDLL:
    """Load (once) and return the shared library; raise loudly if it is not there."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -m resnet_accel_b200.build` "
                "(nvcc, sm_100a).  resnet_accel_b200 has no CPU fallback.")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(handle, name)        # AttributeError here = ABI/header mismatch
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        raise AcceleratorError(rc, lib().accel_last_error_string().decode("utf-8", "replace"))
