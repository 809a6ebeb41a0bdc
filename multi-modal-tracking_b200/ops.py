"""Thin torch-tensor wrappers over the C ABI (include/mmt_b200.h).

torch is used here only as the owner of device memory and of the current CUDA stream; every
function forwards raw pointers to libmmt_b200.so and raises RuntimeError on a non-zero status.
Nothing in this module computes on the CPU or with torch ops.
"""
from __future__ import annotations

import ctypes
from ctypes import c_float, c_int, c_void_p

import torch

from . import _lib

ACT_NONE, ACT_GELU, ACT_RELU = 0, 1, 2


def _ptr(t):
    return c_void_p(0 if t is None else t.data_ptr())


def _stream():
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            # same convention as the reference's native ops (prroi_pool/functional.py:62-63)
            raise NotImplementedError("mmt_b200 ops are CUDA-only (no CPU fallback)")


_gemm_bf16 = _lib.fn("mmt_gemm_bf16")
_gemm_f32 = _lib.fn("mmt_gemm_f32")


def gemm(a, w, bias=None, act=ACT_NONE, resid=None, rowadd=None, out=None, out_dtype=None, max_ctas=0):
    """out = act(a @ w.T + bias) + rowadd[row % period] + resid.

    a: [M, K] (may be a row-strided view), w: [N, K]; both bf16 (tcgen05 kernel) or both fp32 (parity kernel).
    bias [N] / rowadd [period, N] / resid [M, N] are fp32.  `out` may be a column-slice view.
    """
    _need_cuda(a, w, bias, resid, rowadd, out)
    assert a.dim() == 2 and w.dim() == 2 and a.shape[1] == w.shape[1], (a.shape, w.shape)
    assert a.stride(1) == 1 and w.stride(1) == 1
    M, K = a.shape
    N = w.shape[0]
    is_bf16 = a.dtype == torch.bfloat16
    assert w.dtype == a.dtype
    if out is None:
        if out_dtype is None:
            out_dtype = torch.float32 if (resid is not None or not is_bf16) else torch.bfloat16
        out = torch.empty((M, N), device=a.device, dtype=out_dtype)
    assert out.shape == (M, N) and out.stride(1) == 1
    for t in (bias, resid, rowadd):
        assert t is None or t.dtype == torch.float32
    assert bias is None or (bias.numel() == N and bias.is_contiguous())
    assert rowadd is None or (rowadd.shape[1] == N and rowadd.is_contiguous())
    assert resid is None or (resid.shape == (M, N) and resid.stride(1) == 1)
    ldr = resid.stride(0) if resid is not None else 0
    period = rowadd.shape[0] if rowadd is not None else 0
    if is_bf16:
        st = _gemm_bf16(_ptr(a), c_int(a.stride(0)), _ptr(w), c_int(w.stride(0)), c_int(M), c_int(N), c_int(K),
                        _ptr(bias), c_int(act), _ptr(resid), c_int(ldr), _ptr(rowadd), c_int(period), _ptr(out),
                        c_int(out.stride(0)), c_int(1 if out.dtype == torch.float32 else 0), c_int(max_ctas),
                        _stream())
        _lib.check(st, "mmt_gemm_bf16")
    else:
        assert a.dtype == torch.float32 and out.dtype == torch.float32
        st = _gemm_f32(_ptr(a), c_int(a.stride(0)), _ptr(w), c_int(w.stride(0)), c_int(M), c_int(N), c_int(K),
                       _ptr(bias), c_int(act), _ptr(resid), c_int(ldr), _ptr(rowadd), c_int(period), _ptr(out),
                       c_int(out.stride(0)), _stream())
        _lib.check(st, "mmt_gemm_f32")
    return out
