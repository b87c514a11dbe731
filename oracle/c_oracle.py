"""
ctypes binding of oracle/liboracle_c.so (the plain-C restatement, fp8_oracle.c).
TEST INFRASTRUCTURE ONLY -- see fp8_oracle.c's header for who may load it.
"""

from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle_c.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "fp8_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle_c.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = ctypes.CDLL(_LIB_PATH)
        vp, sz, u32, i32, f32 = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint32, ctypes.c_int, ctypes.c_float
        L.fp8o_decode.restype = f32
        L.fp8o_decode.argtypes = [ctypes.c_uint8]
        L.fp8o_encode.restype = ctypes.c_uint8
        L.fp8o_encode.argtypes = [f32]
        for name in ("fp8o_encode_f32", "fp8o_encode_bf16", "fp8o_encode_f16"):
            getattr(L, name).restype = None
            getattr(L, name).argtypes = [vp, vp, sz]
        L.fp8o_to_half.restype = None
        L.fp8o_to_half.argtypes = [vp, vp, sz, f32, i32]
        L.fp8o_vecmat.restype = None
        L.fp8o_vecmat.argtypes = [vp, vp, vp, vp, vp, i32, u32, u32]
        L.fp8o_matmul.restype = None
        L.fp8o_matmul.argtypes = [vp, vp, vp, vp, i32, vp, i32, u32, u32, u32]
        L.fp8o_num_threads.restype = i32
        L.fp8o_set_threads.argtypes = [i32]
        _lib = L
    return _lib


def _p(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


def num_threads() -> int:
    return int(lib().fp8o_num_threads())


def set_threads(n: int) -> None:
    lib().fp8o_set_threads(int(n))


def decode_table() -> np.ndarray:
    L = lib()
    return np.array([L.fp8o_decode(b) for b in range(256)], dtype=np.float32)


def decode_table_e5m2() -> np.ndarray:
    L = lib()
    L.fp8o_decode_e5m2.restype = ctypes.c_float
    L.fp8o_decode_e5m2.argtypes = [ctypes.c_uint8]
    return np.array([L.fp8o_decode_e5m2(b) for b in range(256)], dtype=np.float32)


def encode(x: np.ndarray) -> np.ndarray:
    """x: float32, float16, or uint16 holding bf16 bits (pass kind='bf16')."""
    x = np.ascontiguousarray(x)
    out = np.empty(x.shape, dtype=np.uint8)
    if x.dtype == np.float32:
        lib().fp8o_encode_f32(_p(x), _p(out), x.size)
    elif x.dtype == np.float16:
        lib().fp8o_encode_f16(_p(x), _p(out), x.size)
    else:
        raise TypeError(x.dtype)
    return out


def encode_bf16_bits(u16: np.ndarray) -> np.ndarray:
    u16 = np.ascontiguousarray(u16, dtype=np.uint16)
    out = np.empty(u16.shape, dtype=np.uint8)
    lib().fp8o_encode_bf16(_p(u16), _p(out), u16.size)
    return out


def to_half(u8: np.ndarray, scale=None) -> np.ndarray:
    u8 = np.ascontiguousarray(u8, dtype=np.uint8)
    out = np.empty(u8.shape, dtype=np.float16)
    lib().fp8o_to_half(_p(u8), _p(out), u8.size, float(scale) if scale is not None else 1.0,
                       0 if scale is None else 1)
    return out


def scaled_mm(A: np.ndarray, B: np.ndarray, sa: np.ndarray, sb: np.ndarray) -> np.ndarray:
    """fp32 (M,N) = the reference kernels' result: vecmat when M == 1, else matmul
    (dispatch rule of fp8_mps_native.py:78-93)."""
    A = np.ascontiguousarray(A, dtype=np.uint8)
    B = np.ascontiguousarray(B, dtype=np.uint8)
    sa = np.ascontiguousarray(sa, dtype=np.float32).reshape(-1)
    sb = np.ascontiguousarray(sb, dtype=np.float32).reshape(-1)
    M, K = A.shape
    N = B.shape[0]
    assert B.shape[1] == K and sa.size in (1, M) and sb.size in (1, N)
    C = np.empty((M, N), dtype=np.float32)
    if M == 1:
        lib().fp8o_vecmat(_p(A), _p(B), _p(C), _p(sa), _p(sb), sb.size, N, K)
    else:
        lib().fp8o_matmul(_p(A), _p(B), _p(C), _p(sa), sa.size, _p(sb), sb.size, M, N, K)
    return C
