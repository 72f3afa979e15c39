"""ctypes bindings for the two compiled CPU checkers.  TEST INFRASTRUCTURE ONLY.

* ``liborc``  - oracle/c/bsr_oracle.c, our plain-C restatement ("port").
* ``libref``  - oracle/_ref/libref_golden.so, the unmodified reference C++ golden
  model behind an extern "C" doorway (oracle/ref_wrap.cpp) ("reference").

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ORC_PATH = os.path.join(_HERE, "_build", "liborc.so")
_REF_PATH = os.path.join(_HERE, "_ref", "libref_golden.so")

_i8p = np.ctypeslib.ndpointer(dtype=np.int8, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")


def build(ref: bool = True) -> None:
    """Compile the checkers (gcc/g++ via oracle/Makefile).  Building is not using."""
    targets = ["oracle"] + (["ref"] if ref else [])
    subprocess.run(["make", "-s", "-C", _HERE] + targets, check=True)


def _opt(a, dtype):
    if a is None:
        return None
    return np.ascontiguousarray(a, dtype=dtype).ctypes.data_as(C.c_void_p)


_orc = None


def orc():
    global _orc
    if _orc is None:
        if not os.path.exists(_ORC_PATH):
            build(ref=False)
        lib = C.CDLL(_ORC_PATH)
        lib.orc_bsr_gemm_i32.argtypes = [_i8p, C.c_int64, C.c_int64, C.c_int64, _i32p, _i32p, C.c_void_p,
                                         C.c_int32, C.c_int32, _i32p, C.c_int64]
        lib.orc_bsr_gemm_i32.restype = None
        lib.orc_conv_bsr_layer.argtypes = [_i8p] + [C.c_int32] * 7 + [_i32p, _i32p, C.c_void_p, C.c_int32, C.c_int32,
                                                                      C.c_int32, C.c_void_p, C.c_int32, C.c_void_p,
                                                                      C.c_void_p, C.c_float, C.c_float, C.c_float,
                                                                      _i8p]
        lib.orc_conv_bsr_layer.restype = C.c_uint64
        lib.orc_requant_channel.argtypes = [_i32p, _i8p, C.c_int64, C.c_int64, C.c_int64, _f32p, C.c_void_p,
                                            C.c_int32]
        lib.orc_requant_channel.restype = C.c_uint64
        lib.orc_add_residual.argtypes = [_i8p, _i8p, _i8p, C.c_int64, C.c_float, C.c_float, C.c_float]
        lib.orc_add_residual.restype = None
        lib.orc_maxpool.argtypes = [_i8p, _i8p, C.c_int64] + [C.c_int32] * 5
        lib.orc_maxpool.restype = None
        lib.orc_avgpool.argtypes = [_i8p, _i8p, C.c_int64, C.c_int32, C.c_int32]
        lib.orc_avgpool.restype = None
        lib.orc_bsr_gemm_i32_conva.argtypes = [_i8p, C.c_int64, C.c_int64, C.c_int64, _i32p, _i32p, C.c_void_p,
                                               C.c_int32, C.c_int32, _i32p]
        lib.orc_bsr_gemm_i32_conva.restype = None
        _orc = lib
    return _orc


def have_ref() -> bool:
    return os.path.exists(_REF_PATH)


_ref = None


def ref():
    global _ref
    if _ref is None:
        lib = C.CDLL(_REF_PATH)
        sz = C.c_size_t
        lib.ref_matmul_int8.argtypes = [_i8p, _i8p, _i32p, sz, sz, sz]
        lib.ref_pack_and_bsr_matmul_int8.argtypes = [_i8p, _i8p, _i32p, sz, sz, sz]
        lib.ref_pack_and_bsr_matmul_int8.restype = C.c_long
        lib.ref_pack_to_bsr.argtypes = [_i8p, sz, sz, _i64p, _i64p, _i8p]
        lib.ref_pack_to_bsr.restype = C.c_long
        lib.ref_relu_int8.argtypes = [_i8p, sz]
        lib.ref_relu_int32.argtypes = [_i32p, sz]
        lib.ref_relu6_int8.argtypes = [_i8p, sz, C.c_float]
        lib.ref_requantize_int32_to_int8.argtypes = [_i32p, _i8p, sz, C.c_float, C.c_float]
        lib.ref_add_residual_int8.argtypes = [_i8p, _i8p, _i8p, sz, C.c_float, C.c_float, C.c_float]
        lib.ref_maxpool2d_int8.argtypes = [_i8p, _i8p, sz, sz, sz, sz, sz]
        lib.ref_avgpool_global_int8.argtypes = [_i8p, _i8p, sz, sz, sz]
        lib.ref_im2col_int8.argtypes = [_i8p, _i8p] + [sz] * 8
        lib.ref_conv2d_int8_simple.argtypes = [_i8p, _i8p, C.c_void_p, _i32p] + [sz] * 7
        lib.ref_conv2d_int8_im2col.argtypes = [_i8p, _i8p, C.c_void_p, _i32p] + [sz] * 7
        lib.ref_conv_layer_image.argtypes = [_i8p, _i8p, C.c_void_p] + [sz] * 7 + [C.c_int, _f32p, C.c_float,
                                                                                   C.c_void_p, C.c_float, C.c_float,
                                                                                   C.c_float, _i32p, _i8p]
        for name in dir(lib):
            pass
        _ref = lib
    return _ref


# ---------------------------------------------------------------------------- liborc wrappers
def bsr_gemm_i32(X: np.ndarray, row_ptr, col_idx, blocks) -> np.ndarray:
    X = np.ascontiguousarray(X, dtype=np.int8)
    blocks = np.ascontiguousarray(blocks, dtype=np.int8)
    b = blocks.shape[1] if blocks.ndim == 3 and blocks.shape[0] else 14
    nbr = len(row_ptr) - 1
    M, K = X.shape
    Y = np.zeros((M, nbr * b), dtype=np.int32)
    orc().orc_bsr_gemm_i32(X, M, K, K, np.ascontiguousarray(row_ptr, np.int32),
                           np.ascontiguousarray(col_idx, np.int32) if len(col_idx) else np.zeros(1, np.int32),
                           blocks.ctypes.data_as(C.c_void_p), nbr, b, Y, nbr * b)
    return Y


def conv_bsr_layer(x: np.ndarray, row_ptr, col_idx, blocks, cout: int, k: int, stride: int, pad: int,
                   bias=None, relu: bool = False, sf: Optional[np.ndarray] = None, residual=None,
                   res_scales: Tuple[float, float, float] = (1.0, 1.0, 1.0)) -> Tuple[np.ndarray, int]:
    x = np.ascontiguousarray(x, dtype=np.int8)
    B, Cc, H, W = x.shape
    Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    blocks = np.ascontiguousarray(blocks, dtype=np.int8)
    b = blocks.shape[1] if blocks.shape[0] else 14
    out = np.zeros((B, cout, Ho, Wo), dtype=np.int8)
    sf = np.ascontiguousarray(sf, dtype=np.float32)
    sat = orc().orc_conv_bsr_layer(
        x, B, Cc, H, W, k, stride, pad, np.ascontiguousarray(row_ptr, np.int32),
        np.ascontiguousarray(col_idx, np.int32) if len(col_idx) else np.zeros(1, np.int32),
        blocks.ctypes.data_as(C.c_void_p), len(row_ptr) - 1, b, cout, _opt(bias, np.int32), int(relu),
        sf.ctypes.data_as(C.c_void_p), _opt(residual, np.int8), *[float(s) for s in res_scales], out)
    return out, int(sat)


# ---------------------------------------------------------------------------- libref wrappers
def ref_conv_layer_image(x_chw, w_oihw, bias, stride, pad, relu, in_scale_c, out_scale, residual=None,
                         res_scales=(1.0, 1.0, 1.0)) -> np.ndarray:
    x = np.ascontiguousarray(x_chw, np.int8)
    w = np.ascontiguousarray(w_oihw, np.int8)
    Cout, Cin, k, _ = w.shape
    H, W = x.shape[1:]
    Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    scratch = np.empty(Cout * Ho * Wo, np.int32)
    out = np.empty((Cout, Ho, Wo), np.int8)
    ref().ref_conv_layer_image(x, w, _opt(bias, np.int32), Cin, H, W, Cout, k, stride, pad, int(relu),
                               np.ascontiguousarray(in_scale_c, np.float32), float(out_scale),
                               _opt(residual, np.int8), *[float(s) for s in res_scales], scratch, out)
    return out
