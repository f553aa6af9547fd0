"""Thin torch-tensor -> C-ABI wrappers (pointers, sizes, current stream).

Nothing here computes: every function marshals arguments for a libcoopcap entry point and
launches it on torch's current CUDA stream.  Device memory, streams and autograd bookkeeping come
from PyTorch; the arithmetic is in csrc/*.cu.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import GemmArgs, check


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _req_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.CoopcapError("coopcap ops need CUDA tensors (there is no CPU path)")


def device_info():
    sm, ma, mi = C.c_int(), C.c_int(), C.c_int()
    check(_lib.load().coopcap_device_info(C.byref(sm), C.byref(ma), C.byref(mi)))
    return sm.value, ma.value, mi.value


def gemm(A, B, M, N, K, *, a_major=0, b_major=0, alpha=1.0, bias=None, row_scale=None, relu=False,
         mode=0, out=None, out16=None, out_t16=None, split_k=1, tile_n=0, backend=0):
    """C[M,N] = alpha * A·Bᵀ (+bias)(relu)(*row_scale). A/B bf16 (kind 0) or fp32 (kind 1, tf32).

    a_major/b_major 0: operand stored [rows, K]; 1: stored [K, rows]. Operands are 2-D views with
    unit inner stride; the row stride is passed as the leading dimension.
    """
    _req_cuda(A, B, bias, row_scale, out, out16, out_t16)
    if A.dtype != B.dtype or A.dtype not in (torch.bfloat16, torch.float32):
        raise _lib.CoopcapError(f"gemm: unsupported operand dtypes {A.dtype}/{B.dtype}")
    for t in (A, B, out, out16, out_t16):
        if t is not None and (t.dim() != 2 or t.stride(1) != 1):
            raise _lib.CoopcapError("gemm: operands must be 2-D with unit inner stride")
    a = GemmArgs()
    a.kind = 0 if A.dtype == torch.bfloat16 else 1
    a.a_major, a.b_major = a_major, b_major
    a.A, a.lda = A.data_ptr(), A.stride(0)
    a.B, a.ldb = B.data_ptr(), B.stride(0)
    a.M, a.N, a.K = M, N, K
    a.alpha = alpha
    a.bias = None if bias is None else bias.data_ptr()
    a.row_scale = None if row_scale is None else row_scale.data_ptr()
    a.relu = int(relu)
    a.mode = mode
    if out is not None:
        assert out.dtype == torch.float32
        a.C, a.ldc = out.data_ptr(), out.stride(0)
    if out16 is not None:
        assert out16.dtype == torch.bfloat16
        a.C16, a.ldc16 = out16.data_ptr(), out16.stride(0)
    if out_t16 is not None:
        assert out_t16.dtype == torch.bfloat16
        a.Ct16, a.ldct = out_t16.data_ptr(), out_t16.stride(0)
    a.split_k, a.tile_n, a.backend = split_k, tile_n, backend
    check(_lib.load().coopcap_gemm(C.byref(a), _stream()))


def cast_bf16(src, dst=None, dst_t=None):
    """fp32 [rows, cols] -> bf16 copy and/or transposed bf16 copy."""
    _req_cuda(src, dst, dst_t)
    assert src.dtype == torch.float32 and src.dim() == 2 and src.stride(1) == 1
    rows, cols = src.shape
    check(_lib.load().coopcap_cast_bf16(
        _ptr(src), rows, cols, src.stride(0),
        _ptr(dst), 0 if dst is None else dst.stride(0),
        _ptr(dst_t), 0 if dst_t is None else dst_t.stride(0), _stream()))
    return dst, dst_t
