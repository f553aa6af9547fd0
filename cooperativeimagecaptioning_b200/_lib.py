"""ctypes binding of libcoopcap.so (see include/coopcap.h).

There is no CPU fallback and no alternative backend: if the library is missing or a call fails
the error is raised to the caller.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcoopcap.so")

c_void_p = C.c_void_p
c_int = C.c_int
c_i64 = C.c_int64
c_float = C.c_float


class CoopcapError(RuntimeError):
    pass


class GemmArgs(C.Structure):
    _fields_ = [
        ("kind", c_int), ("a_major", c_int), ("b_major", c_int),
        ("A", c_void_p), ("lda", c_i64),
        ("B", c_void_p), ("ldb", c_i64),
        ("M", c_int), ("N", c_int), ("K", c_int),
        ("alpha", c_float),
        ("bias", c_void_p), ("row_scale", c_void_p),
        ("relu", c_int), ("mode", c_int),
        ("C", c_void_p), ("ldc", c_i64),
        ("C16", c_void_p), ("ldc16", c_i64),
        ("Ct16", c_void_p), ("ldct", c_i64),
        ("split_k", c_int), ("tile_n", c_int), ("backend", c_int),
    ]


_lib = None


def load() -> C.CDLL:
    """Load libcoopcap.so once; raise loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise CoopcapError(
            f"{LIB_PATH} not found: build it with `python -m cooperativeimagecaptioning_b200.build` "
            "(or __graft_entry__.build()). There is no fallback path.")
    lib = C.CDLL(LIB_PATH)
    lib.coopcap_last_error.restype = C.c_char_p
    lib.coopcap_version.restype = c_int
    _declare(lib)
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        msg = load().coopcap_last_error().decode("utf-8", "replace")
        raise CoopcapError(f"libcoopcap call failed (code {rc}): {msg}")


# name -> argtypes; every function returns int.  Kept in one table so tests can verify that the
# library exports exactly what include/coopcap.h declares.
SIGNATURES = {}


def _sig(name, *argtypes):
    SIGNATURES[name] = list(argtypes)


_sig("coopcap_device_info", C.POINTER(c_int), C.POINTER(c_int), C.POINTER(c_int))
_sig("coopcap_gemm", C.POINTER(GemmArgs), c_void_p)
_sig("coopcap_cast_bf16", c_void_p, c_i64, c_i64, c_i64, c_void_p, c_i64, c_void_p, c_i64, c_void_p)


def _declare(lib):
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = c_int
