"""Thin tensor-level wrappers over the C ABI (include/tt_b200.h).

Every function takes CUDA tensors, passes raw pointers + sizes to libtt_b200.so on torch's
current stream, and returns torch tensors allocated by torch's caching allocator (the library
never allocates).  No torch math happens here.
"""
import ctypes

import torch

from . import _lib
from ._lib import TtGemmParams, c_float, c_int, c_ll, c_void_p

ACT_NONE, ACT_RELU, ACT_GELU = 0, 1, 2


def _stream():
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return c_void_p(t.data_ptr()) if t is not None else c_void_p(0)


def _check_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.TtError('tell_b200 ops need CUDA tensors (no CPU fallback); got %s' % t.device)


def cast_bf16(x, transpose=False, split=0, out=None):
    """fp32 [R,C] -> bf16 GEMM operand ([R,C*rep] or [C,R*rep] when transposed); see tt_cast_bf16."""
    _check_cuda(x)
    assert x.dtype == torch.float32 and x.dim() == 2 and x.stride(1) == 1
    rows, cols = x.shape
    rep = 1 if split == 0 else 3
    shape = (cols, rows * rep) if transpose else (rows, cols * rep)
    if out is None:
        # pad the leading dimension to a multiple of 8 elements (TMA: 16-byte row pitch)
        ld = (shape[1] + 7) // 8 * 8
        buf = torch.empty((shape[0], ld), dtype=torch.bfloat16, device=x.device)
        if ld != shape[1]:
            buf.zero_()
        out = buf[:, :shape[1]]
    _lib.call('tt_cast_bf16', _ptr(x), c_ll(x.stride(0)), _ptr(out), c_ll(out.stride(0)),
              c_int(rows), c_int(cols), c_int(1 if transpose else 0), c_int(split), _stream())
    return out


def gemm_tn(a, b, out=None, out16=None, bias=None, residual=None, alpha=1.0, act=ACT_NONE,
            accumulate=False, m_limit=None, want32=True, want16=False):
    """C[M,N] = act(alpha * a[M,K] @ b[N,K]^T + bias) + residual, bf16 operands, fp32 accumulate."""
    _check_cuda(a, b, out, out16, bias, residual)
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16
    assert a.dim() == 2 and b.dim() == 2 and a.shape[1] == b.shape[1], (a.shape, b.shape)
    assert a.stride(1) == 1 and b.stride(1) == 1
    M, K = a.shape
    N = b.shape[0]
    if out is None and want32:
        out = torch.empty((M, N), dtype=torch.float32, device=a.device)
    if out16 is None and want16:
        out16 = torch.empty((M, N), dtype=torch.bfloat16, device=a.device)
    p = TtGemmParams()
    p.M, p.N, p.K = M, N, K
    p.A, p.lda = a.data_ptr(), a.stride(0)
    p.B, p.ldb = b.data_ptr(), b.stride(0)
    if out is not None:
        assert out.dtype == torch.float32 and out.shape == (M, N) and out.stride(1) == 1
        p.C, p.ldc = out.data_ptr(), out.stride(0)
    if out16 is not None:
        assert out16.dtype == torch.bfloat16 and out16.shape == (M, N) and out16.stride(1) == 1
        p.C16, p.ldc16 = out16.data_ptr(), out16.stride(0)
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() == N and bias.is_contiguous()
        p.bias = bias.data_ptr()
    if residual is not None:
        assert residual.dtype == torch.float32 and residual.shape == (M, N)
        assert residual.stride(1) == 1
        p.residual, p.ldr = residual.data_ptr(), residual.stride(0)
    p.alpha = alpha
    p.act = act
    p.accumulate = 1 if accumulate else 0
    if m_limit is not None:
        assert m_limit.dtype == torch.int32
        p.m_limit = m_limit.data_ptr()
    _lib.call('tt_gemm_bf16_tn', ctypes.byref(p), _stream())
    if out is not None and out16 is not None:
        return out, out16
    return out if out is not None else out16
