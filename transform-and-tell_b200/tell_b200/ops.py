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


def cast_bf16(x, transpose=False, split=0, out=None, seg_stride=0):
    """fp32 [R,C] -> bf16 GEMM operand ([R,C*rep] or [C,R*rep] when transposed); see tt_cast_bf16."""
    _check_cuda(x)
    assert x.dtype == torch.float32 and x.dim() == 2 and x.stride(1) == 1
    from . import weight_bank
    for bank in weight_bank.LIVE:
        if bank.is_unwritten_handle(x):
            raise _lib.TtError('cast of a weight-bank handle: the fp32 effective weights are not materialised '
                               '(WeightBank.write_w32 = False); the bf16 operand must come from the bank')
    rows, cols = x.shape
    rep = 1 if split == 0 else 3
    shape = (cols, rows * rep) if transpose else (rows, cols * rep)
    if out is None:
        # pad the leading dimension to a multiple of 8 elements (TMA: 16-byte row pitch)
        ld = (shape[1] + 7) // 8 * 8
        buf = torch.empty((shape[0], ld), dtype=torch.bfloat16, device=x.device)
        if ld != shape[1]:
            buf[:, shape[1]:].zero_()      # only the pad columns (outside every TMA extent anyway)
        out = buf[:, :shape[1]]
    _lib.call('tt_cast_bf16', _ptr(x), c_ll(x.stride(0)), _ptr(out), c_ll(out.stride(0)),
              c_int(rows), c_int(cols), c_int(1 if transpose else 0), c_int(split),
              c_ll(seg_stride), _stream())
    return out


def bf16_buffer(rows, cols, device):
    """Zero-initialised bf16 operand buffer whose row pitch is a multiple of 8 elements (TMA)."""
    ld = (cols + 7) // 8 * 8
    return torch.zeros((rows, ld), dtype=torch.bfloat16, device=device)[:, :cols]


def f32_padded(rows, cols, like, zero=False):
    """fp32 [rows, cols] view whose row pitch is a multiple of 8 floats: a GEMM output with an odd
    width (5002-way head, 30265-word tail) then takes the 16-byte vector epilogue instead of the
    scalar one (3-4x faster for those launches)."""
    ld = (cols + 7) // 8 * 8
    alloc = torch.zeros if zero else torch.empty
    return alloc((rows, ld), dtype=torch.float32, device=like.device)[:, :cols]


def scalar_mul(a, b):
    out = torch.empty_like(a)
    _lib.call('tt_scalar_mul', _ptr(a), _ptr(b), _ptr(out), _stream())
    return out


def gemm_tn(a, b, out=None, out16=None, bias=None, residual=None, alpha=1.0, act=ACT_NONE,
            accumulate=False, m_limit=None, want32=True, want16=False, residual16=None,
            trans_a=False, trans_b=False, m_hint=0, k_limit=None, col_stats=None):
    """C[M,N] = act(alpha * (a[M,K] @ b[N,K]^T + bias) + residual), bf16 operands, fp32 accumulate."""
    _check_cuda(a, b, out, out16, bias, residual)
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16
    assert a.dim() == 2 and b.dim() == 2 and a.stride(1) == 1 and b.stride(1) == 1
    # trans_a: a is stored [K, M]; trans_b: b is stored [K, N]
    K, M = (a.shape[0], a.shape[1]) if trans_a else (a.shape[1], a.shape[0])
    Kb, N = (b.shape[0], b.shape[1]) if trans_b else (b.shape[1], b.shape[0])
    assert K == Kb, (a.shape, b.shape, trans_a, trans_b)
    if out is None and want32:
        # with a dynamic row limit the untouched rows must be defined (they feed later GEMMs)
        alloc = torch.zeros if m_limit is not None else torch.empty
        out = alloc((M, N), dtype=torch.float32, device=a.device)
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
    if residual16 is not None:
        assert residual16.dtype == torch.bfloat16 and residual16.shape == (M, N)
        p.residual16, p.ldr16 = residual16.data_ptr(), residual16.stride(0)
    p.alpha = alpha
    p.act = act
    p.accumulate = 1 if accumulate else 0
    p.trans_a, p.trans_b = (1 if trans_a else 0), (1 if trans_b else 0)
    if m_limit is not None:
        assert m_limit.dtype == torch.int32
        p.m_limit = m_limit.data_ptr()
        p.m_hint = int(m_hint)
    if k_limit is not None:      # operands are zero for k >= *k_limit (device int32)
        assert k_limit.dtype == torch.int32
        p.k_limit = k_limit.data_ptr()
    if col_stats is not None:    # fp32 [2N], zeroed by the caller: += column sums / sums of squares
        assert col_stats.dtype == torch.float32 and col_stats.numel() == 2 * N and out is None
        p.col_stats = col_stats.data_ptr()
    if _lib.PROFILE is not None:
        _lib.GEMM_FLOPS.append(0.0 if m_limit is not None else 2.0 * M * N * K)
    _lib.call('tt_gemm_bf16_tn', ctypes.byref(p), _stream())
    if out is not None and out16 is not None:
        return out, out16
    return out if out is not None else out16


def gemm_tn_batched(a, b, nbatch, M, N, K, out, ldc_off, a_off=(0, 0), b_off=(0, 0), bias=None,
                    bias_off=0, alpha=1.0, act=ACT_NONE, trans_a=False, trans_b=False):
    """nbatch same-shape problems C_z = act(alpha * (A_z . B_z^T + bias_z)) in ONE launch.  a / b are
    the SHARED bf16 buffers the per-batch operands are blocks of (batch z starts a_off[0] * z
    elements along the contiguous axis and a_off[1] * z rows further; same for b); out is the shared
    fp32 or bf16 output buffer, batch z written at + z * ldc_off elements with row pitch
    out.stride(0); bias a shared fp32 vector read at + z * bias_off."""
    _check_cuda(a, b, out, bias)
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16 and a.stride(1) == 1 and b.stride(1) == 1
    p = TtGemmParams()
    p.M, p.N, p.K = M, N, K
    p.A, p.lda = a.data_ptr(), a.stride(0)
    p.B, p.ldb = b.data_ptr(), b.stride(0)
    if out.dtype == torch.float32:
        p.C, p.ldc = out.data_ptr(), out.stride(0)
    else:
        assert out.dtype == torch.bfloat16
        p.C16, p.ldc16 = out.data_ptr(), out.stride(0)
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.is_contiguous()
        p.bias = bias.data_ptr()
    p.alpha, p.act = alpha, act
    p.trans_a, p.trans_b = (1 if trans_a else 0), (1 if trans_b else 0)
    p.nbatch = nbatch
    p.a_off0, p.a_off1 = a_off
    p.b_off0, p.b_off1 = b_off
    p.bias_off, p.c_off = bias_off, ldc_off
    if _lib.PROFILE is not None:
        _lib.GEMM_FLOPS.append(2.0 * nbatch * M * N * K)
    _lib.call('tt_gemm_bf16_tn', ctypes.byref(p), _stream())
    return out


# --------------------------------------------------------------------------- row-wise kernels
def _ull(x):
    return ctypes.c_ulonglong(int(x) & 0xFFFFFFFFFFFFFFFF)


def _f32(*shape, like):
    return torch.empty(shape, dtype=torch.float32, device=like.device)


def ln_fwd(h, res, gamma, beta, eps=1e-5, p=0.0, seed=0, out=None):
    """h [N,E] is overwritten with x = res + dropout(h); returns (y, mean, rstd)."""
    _check_cuda(h, res, gamma, beta)
    N, E = h.shape
    assert h.is_contiguous() and (res is None or res.is_contiguous())
    y = out if out is not None else _f32(N, E, like=h)
    mean, rstd = _f32(N, like=h), _f32(N, like=h)
    _lib.call('tt_ln_fwd', _ptr(h), _ptr(res), _ptr(gamma), _ptr(beta), _ptr(y),
              c_ll(y.stride(0)), _ptr(mean), _ptr(rstd), c_int(N), c_int(E), c_float(eps),
              c_float(p), _ull(seed), _stream())
    return y, mean, rstd


def ln_bwd(dy, x, mean, rstd, gamma, p=0.0, seed=0, want_dx=True, want_dh=True, dgamma=None,
           dbeta=None):
    _check_cuda(dy, x)
    N, E = x.shape
    assert dy.stride(1) == 1
    dx = _f32(N, E, like=x) if want_dx else None
    dh = _f32(N, E, like=x) if want_dh else None
    _lib.call('tt_ln_bwd', _ptr(dy), c_ll(dy.stride(0)), _ptr(x), _ptr(mean), _ptr(rstd),
              _ptr(gamma), _ptr(dx), _ptr(dh), _ptr(dgamma), _ptr(dbeta), c_int(N), c_int(E),
              c_float(p), _ull(seed), _stream())
    return dx, dh


def glu_fwd(h):
    N, C2 = h.shape
    out = _f32(N, C2 // 2, like=h)
    _lib.call('tt_glu_fwd', _ptr(h), _ptr(out), c_ll(N), c_int(C2 // 2), _stream())
    return out


def glu_bwd(dout, h):
    N, C2 = h.shape
    dh = torch.empty_like(h)
    _lib.call('tt_glu_bwd', _ptr(dout), _ptr(h), _ptr(dh), c_ll(N), c_int(C2 // 2), _stream())
    return dh


def bf16_like(rows, cols, device):
    """Uninitialised bf16 [rows, cols] view whose row pitch is a multiple of 8 elements (TMA); the pad
    columns lie outside every TMA extent and are zeroed only when they exist."""
    ld = (cols + 7) // 8 * 8
    if ld == cols:
        return torch.empty((rows, cols), dtype=torch.bfloat16, device=device)
    return torch.zeros((rows, ld), dtype=torch.bfloat16, device=device)[:, :cols]


def dropout_tw(x, p, seed, want32=True):
    """dropout with the bf16 operand twin: x [N,C] contiguous -> (y fp32 or None, y16)."""
    assert x.is_contiguous() and x.dim() == 2 and x.shape[1] % 8 == 0
    y = torch.empty_like(x) if want32 else None
    y16 = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    _lib.call('tt_dropout_tw', _ptr(x), _ptr(y), _ptr(y16), c_ll(x.numel()), c_float(p), _ull(seed),
              _stream())
    return y, y16


def relu_bwd_tw(dy, y):
    assert dy.is_contiguous() and y.is_contiguous() and dy.shape == y.shape and dy.shape[-1] % 8 == 0
    dx = torch.empty_like(dy)
    dx16 = torch.empty(dy.shape, dtype=torch.bfloat16, device=dy.device)
    _lib.call('tt_relu_bwd_tw', _ptr(dy), _ptr(y), _ptr(dx), _ptr(dx16), c_ll(dy.numel()), _stream())
    return dx, dx16


def glu_fwd_tw(h):
    N, C2 = h.shape
    assert h.is_contiguous() and (C2 // 2) % 8 == 0
    out = _f32(N, C2 // 2, like=h)
    out16 = torch.empty((N, C2 // 2), dtype=torch.bfloat16, device=h.device)
    _lib.call('tt_glu_fwd_tw', _ptr(h), _ptr(out), _ptr(out16), c_ll(N), c_int(C2 // 2), _stream())
    return out, out16


def glu_bwd_tw(dout, h):
    N, C2 = h.shape
    assert dout.is_contiguous() and h.is_contiguous() and (C2 // 2) % 8 == 0
    dh = torch.empty_like(h)
    dh16 = torch.empty(h.shape, dtype=torch.bfloat16, device=h.device)
    _lib.call('tt_glu_bwd_tw', _ptr(dout), _ptr(h), _ptr(dh), _ptr(dh16), c_ll(N), c_int(C2 // 2),
              _stream())
    return dh, dh16


def ln_fwd_multi(hs, res, gammas, betas, eps=1e-5, p=0.0, seeds=None, want32=True, want16=True):
    """n <= 4 LayerNorms sharing the residual `res` in one launch: hs[c] [N,E] is overwritten with
    res + dropout_c(hs[c]); returns (Y fp32 [N,n*E] or None, Y16 bf16 [N,n*E] or None, means, rstds)."""
    n = len(hs)
    N, E = hs[0].shape
    _check_cuda(*hs, res)
    assert 1 <= n <= 4 and E % 8 == 0 and all(h.is_contiguous() and h.shape == (N, E) for h in hs)
    assert res is None or (res.is_contiguous() and res.shape == (N, E))
    dev = hs[0].device
    Y = torch.empty((N, n * E), dtype=torch.float32, device=dev) if want32 else None
    Y16 = torch.empty((N, n * E), dtype=torch.bfloat16, device=dev) if want16 else None
    stats = torch.empty((2 * n, N), dtype=torch.float32, device=dev)
    means, rstds = [stats[2 * c] for c in range(n)], [stats[2 * c + 1] for c in range(n)]
    a = _lib.TtLnFwdMulti()
    for c in range(n):
        a.h[c], a.gamma[c], a.beta[c] = hs[c].data_ptr(), gammas[c].data_ptr(), betas[c].data_ptr()
        a.mean[c], a.rstd[c] = means[c].data_ptr(), rstds[c].data_ptr()
        a.seed[c] = (int(seeds[c]) if seeds is not None else 0) & 0xFFFFFFFFFFFFFFFF
    a.res = _dp(res)
    a.y, a.ldy = _dp(Y), (n * E)
    a.y16, a.ldy16 = _dp(Y16), (n * E)
    a.n, a.N, a.E, a.eps, a.p_drop = n, N, E, eps, p
    _lib.call('tt_ln_fwd_multi', ctypes.byref(a), _stream())
    return Y, Y16, means, rstds


def ln_bwd_multi(dY, xs, means, rstds, gammas, p=0.0, seeds=None, want_dx=True, want_dh32=False,
                 want_dh16=True, dgammas=None, dbetas=None):
    """Backward of ln_fwd_multi in one launch.  dY [N, n*E] (row stride free).  Returns
    (dX [N,E] = sum_c dx_c or None, [dh_c fp32] or None, dh16 [N, n*E] bf16 or None)."""
    n = len(xs)
    N, E = xs[0].shape
    _check_cuda(dY, *xs)
    assert dY.stride(1) == 1 and dY.shape == (N, n * E)
    dev = dY.device
    dX = torch.empty((N, E), dtype=torch.float32, device=dev) if want_dx else None
    dhs = [torch.empty((N, E), dtype=torch.float32, device=dev) for _ in range(n)] if want_dh32 else None
    dh16 = torch.empty((N, n * E), dtype=torch.bfloat16, device=dev) if want_dh16 else None
    a = _lib.TtLnBwdMulti()
    for c in range(n):
        a.x[c], a.mean[c], a.rstd[c] = xs[c].data_ptr(), means[c].data_ptr(), rstds[c].data_ptr()
        a.gamma[c] = gammas[c].data_ptr()
        a.dgamma[c] = _dp(dgammas[c]) if dgammas is not None else None
        a.dbeta[c] = _dp(dbetas[c]) if dbetas is not None else None
        a.dh[c] = dhs[c].data_ptr() if dhs is not None else None
        a.seed[c] = (int(seeds[c]) if seeds is not None else 0) & 0xFFFFFFFFFFFFFFFF
    a.dy, a.lddy = dY.data_ptr(), dY.stride(0)
    a.dx = _dp(dX)
    a.dh16, a.lddh16 = _dp(dh16), n * E
    a.n, a.N, a.E, a.p_drop = n, N, E, p
    _lib.call('tt_ln_bwd_multi', ctypes.byref(a), _stream())
    return dX, dhs, dh16


def dropout(x, p, seed):
    assert x.is_contiguous()
    y = torch.empty_like(x)
    _lib.call('tt_dropout', _ptr(x), _ptr(y), c_ll(x.numel()), c_float(p), _ull(seed), _stream())
    return y


def axpby(x, y, a, b):
    assert x.is_contiguous() and y.is_contiguous() and x.numel() == y.numel()
    _lib.call('tt_axpby', _ptr(x), _ptr(y), c_ll(x.numel()), c_float(a), c_float(b), _stream())
    return y


def wnorm_fwd(v, g):
    O, I = v.shape
    w, norm = torch.empty_like(v), _f32(O, like=v)
    _lib.call('tt_wnorm_fwd', _ptr(v), _ptr(g), _ptr(w), _ptr(norm), c_int(O), c_int(I), _stream())
    return w, norm


def wnorm_bwd(dw, v, g, norm):
    O, I = v.shape
    dv, dg = torch.empty_like(v), torch.empty_like(g)
    _lib.call('tt_wnorm_bwd', _ptr(dw), _ptr(v), _ptr(g), _ptr(norm), _ptr(dv), _ptr(dg),
              c_int(O), c_int(I), _stream())
    return dv, dg


def nan_rows_(x):
    """In place: zero NaN rows of x [..., D]; returns the bool row mask."""
    D = x.shape[-1]
    if x.numel() == 0:
        # a batch without any face / object is [B,1,0] (np.array([[]]) in the reference reader):
        # isnan(x).any(-1) over an empty axis is all False (transformer_faces_objects.py:374-379)
        _check_cuda(x)
        return torch.zeros(x.shape[:-1], dtype=torch.bool, device=x.device)
    R = x.numel() // D
    mask = torch.empty(x.shape[:-1], dtype=torch.uint8, device=x.device)
    assert x.is_contiguous()
    _lib.call('tt_nan_rows', _ptr(x), _ptr(mask), c_int(R), c_int(D), _stream())
    return mask.bool()


def zeros_f32(shape, like):
    """Zero-filled fp32 tensor for accumulate-into gradients; served from the per-step zero arena
    (config.ZeroArena, one memset per step) when the trainer enabled it."""
    from . import config
    arena = config.zero_arena
    if arena is not None and arena.buf.device == like.device:
        n = 1
        for d in shape:
            n *= d
        v = arena.take(n)
        if v is not None:
            return v.view(shape)
    return torch.zeros(shape, dtype=torch.float32, device=like.device)


def colsum(x, scale=1.0, out=None, accumulate=False):
    M, N = x.shape
    if out is None:
        from . import config
        if config.zero_arena is not None:
            out, accumulate = zeros_f32((N,), x), True      # pre-zeroed: skip the memset node
        else:
            out = _f32(N, like=x)
    _lib.call('tt_colsum_bf16' if x.dtype == torch.bfloat16 else 'tt_colsum', _ptr(x),
              c_ll(x.stride(0)), c_int(M), c_int(N), _ptr(out), c_float(scale),
              c_int(1 if accumulate else 0), _stream())
    return out


def relu_bwd(dy, y):
    dx = torch.empty_like(y)
    _lib.call('tt_relu_bwd', _ptr(dy), _ptr(y), _ptr(dx), c_ll(y.numel()), _stream())
    return dx


def layer_mix_fwd(hiddens, w):
    """hiddens bf16 [L, R, E]; w fp32 [L] -> fp32 [R, E]."""
    L = hiddens.shape[0]
    n = hiddens[0].numel()
    out = torch.empty(hiddens.shape[1:], dtype=torch.float32, device=hiddens.device)
    _lib.call('tt_layer_mix_fwd', _ptr(hiddens), c_ll(hiddens.stride(0)), _ptr(w), c_int(L),
              c_ll(n), _ptr(out), _stream())
    return out


def layer_mix_bwd(hiddens, w, dout):
    L = hiddens.shape[0]
    dots, dw = _f32(L, like=w), _f32(L, like=w)
    _lib.call('tt_layer_mix_bwd', _ptr(hiddens), c_ll(hiddens.stride(0)), _ptr(w), _ptr(dout),
              c_int(L), c_ll(hiddens[0].numel()), _ptr(dots), _ptr(dw), _stream())
    return dw


# --------------------------------------------------------------------------- dynamic conv
def dynconv_fwd(x, z, H, K, softmax=True, p=0.0, seed=0, broadcast=False, twin=False):
    """twin=True: also returns out16 (bf16 [T*B, C]), the operand of the following linear2."""
    T, B, C = x.shape
    out = torch.empty_like(x)
    probs = _f32(T, B, H, K, like=x)
    out16 = torch.empty((T * B, C), dtype=torch.bfloat16, device=x.device) if twin and C % 8 == 0 else None
    _lib.call('tt_dynconv_fwd_tw', _ptr(x), _ptr(z), c_ll(0 if broadcast else H * K), _ptr(out),
              _ptr(probs), c_int(T), c_int(B), c_int(C), c_int(H), c_int(K),
              c_int(1 if softmax else 0), c_float(p), _ull(seed), _ptr(out16), _stream())
    if twin:
        return out, probs, out16
    return out, probs


def dynconv_step(window, x_new, z, H, K, softmax=True, broadcast=False, twin=False):
    """One incremental step: window [K-1,B,C] (updated in place), x_new [B,C], z [B,H*K] (or [H,K]
    with broadcast) -> out [B,C] (twin=True: (out, out16))."""
    _check_cuda(window, x_new, z)
    B, C = x_new.shape
    assert x_new.is_contiguous() and z.is_contiguous()
    assert K == 1 or (window.is_contiguous() and window.shape == (K - 1, B, C))
    out = torch.empty_like(x_new)
    out16 = torch.empty((B, C), dtype=torch.bfloat16, device=x_new.device) if twin and C % 8 == 0 else None
    _lib.call('tt_dynconv_step_tw', _ptr(window if K > 1 else None), _ptr(x_new), _ptr(z),
              c_ll(0 if broadcast else H * K), _ptr(out), c_int(B), c_int(C), c_int(H), c_int(K),
              c_int(1 if softmax else 0), _ptr(out16), _stream())
    if twin:
        return out, out16
    return out


def dynconv_bwd(dout, x, probs, H, K, softmax=True, p=0.0, seed=0, twin=False):
    """twin=True: also returns dz16 (bf16 [T*B, H*K] view), the operand of the filter projection's
    backward GEMMs."""
    T, B, C = x.shape
    dx = torch.empty_like(x)
    dz = _f32(T, B, H, K, like=x)
    dz16 = bf16_like(T * B, H * K, x.device) if twin else None
    _lib.call('tt_dynconv_bwd_tw', _ptr(dout), _ptr(x), _ptr(probs), _ptr(dx), _ptr(dz), c_int(T),
              c_int(B), c_int(C), c_int(H), c_int(K), c_int(1 if softmax else 0), c_float(p),
              _ull(seed), _ptr(dz16), c_ll(dz16.stride(0) if dz16 is not None else 0), _stream())
    if twin:
        return dx, dz, dz16
    return dx, dz


# --------------------------------------------------------------------------- attention
def attn_fwd(q, k, v, bias_k, bias_v, mask, T, B, S, H, D, zero_row=True, p=0.0, seed=0, tc=False,
             out=None):
    """q [T*B, >=E] view, k/v [S*B, >=E] views (row strides honoured); returns (out [T*B,E], lse).
    `out` may be a strided view (e.g. a column block of a [T*B, n*E] buffer)."""
    E = H * D
    if out is None:
        out = _f32(T * B, E, like=q)
    lse = _f32(B, H, T, like=q)
    kv16 = S > 0 and k.dtype == torch.bfloat16
    assert not kv16 or tc, 'bf16 keys|values need the tensor-core attention kernels'
    _lib.call(('tt_attn_fwd_tc_kv16' if kv16 else 'tt_attn_fwd_tc') if tc else 'tt_attn_fwd',
              _ptr(q), _ptr(k if S > 0 else None),
              _ptr(v if S > 0 else None),
              _ptr(bias_k), _ptr(bias_v), _ptr(mask), _ptr(out), _ptr(lse), c_int(T), c_int(B),
              c_int(S), c_int(H), c_int(D), c_ll(q.stride(0)),
              c_ll(k.stride(0) if S > 0 else E), c_ll(out.stride(0)), c_int(1 if zero_row else 0),
              c_float(p), _ull(seed), _stream())
    return out, lse


def attn_bwd(dout, q, k, v, bias_k, bias_v, mask, out, lse, dq, dk, dv, dbias_k, dbias_v, T, B, S,
             H, D, zero_row=True, p=0.0, seed=0, tc=False):
    """dq/dk/dv are caller-allocated views with the same row strides as q/k/v."""
    E = H * D
    assert dout.stride(0) == out.stride(0) and dq.stride(0) == q.stride(0)
    if S > 0:
        assert dk.stride(0) == k.stride(0) and dv.stride(0) == k.stride(0)
    kv16 = S > 0 and k.dtype == torch.bfloat16
    assert not kv16 or (tc and dk.dtype == torch.bfloat16 and dv.dtype == torch.bfloat16)
    _lib.call(('tt_attn_bwd_tc_kv16' if kv16 else 'tt_attn_bwd_tc') if tc else 'tt_attn_bwd',
              _ptr(dout), _ptr(q),
              _ptr(k if S > 0 else None),
              _ptr(v if S > 0 else None), _ptr(bias_k), _ptr(bias_v), _ptr(mask), _ptr(out),
              _ptr(lse), _ptr(dq), _ptr(dk if S > 0 else None), _ptr(dv if S > 0 else None),
              _ptr(dbias_k), _ptr(dbias_v), c_int(T), c_int(B), c_int(S), c_int(H), c_int(D),
              c_ll(q.stride(0)), c_ll(k.stride(0) if S > 0 else E), c_ll(out.stride(0)),
              c_int(1 if zero_row else 0), c_float(p), _ull(seed), _stream())


def _dp(t):
    return t.data_ptr() if t is not None else None


def _attn_ctx_array(items, E):
    """items: dicts with q, k, v, bias_k, bias_v, mask, out, lse, S, seed (+ dout, dq, dk, dv, dbias_k,
    dbias_v for the backward) -> ctypes array of TtAttnCtx."""
    arr = (_lib.TtAttnCtx * len(items))()
    for c, it in zip(arr, items):
        S = it['S']
        c.q, c.out, c.lse = _dp(it['q']), _dp(it['out']), _dp(it.get('lse'))
        c.k, c.v = (_dp(it['k']), _dp(it['v'])) if S > 0 else (None, None)
        c.bias_k, c.bias_v = _dp(it.get('bias_k')), _dp(it.get('bias_v'))
        c.mask = _dp(it.get('mask')) if S > 0 else None
        c.S = S
        c.ldq, c.ldo = it['q'].stride(0), it['out'].stride(0)
        c.ldkv = it['k'].stride(0) if (S > 0 and it['k'].dim() == 2) else E
        c.seed = int(it.get('seed', 0)) & 0xFFFFFFFFFFFFFFFF
        c.kv_len = _dp(it.get('kv_len')) if S > 0 else None
        if it.get('out16') is not None:
            c.out16, c.ldo16 = _dp(it['out16']), it['out16'].stride(0)
        if it.get('dq16') is not None:
            c.dq16, c.ldq16 = _dp(it['dq16']), it['dq16'].stride(0)
        if it.get('dsum') is not None:
            c.dsum = _dp(it['dsum'])
        if it.get('dout') is not None:
            c.dout, c.dq = _dp(it['dout']), _dp(it['dq'])
            c.dk, c.dv = (_dp(it['dk']), _dp(it['dv'])) if S > 0 else (None, None)
            c.dbias_k, c.dbias_v = _dp(it.get('dbias_k')), _dp(it.get('dbias_v'))
    return arr


def _attn_kv16(items):
    kinds = {it['k'].dtype for it in items if it['S'] > 0}
    assert len(kinds) <= 1, 'contexts of one launch must store k/v in one dtype'
    return bool(kinds) and kinds.pop() == torch.bfloat16


def attn_fwd_tc_multi(items, T, B, H, D, zero_row=True, p=0.0):
    """One launch for the cross-attention of up to 4 contexts (tensor-core kernels); `items` as in
    _attn_ctx_array; out / lse are caller-allocated (out may be a column block of a wider buffer)."""
    arr = _attn_ctx_array(items, H * D)
    _lib.call('tt_attn_fwd_tc_multi', arr, c_int(len(items)), c_int(T), c_int(B), c_int(H), c_int(D),
              c_int(1 if zero_row else 0), c_float(p), c_int(1 if _attn_kv16(items) else 0), _stream())


def attn_bwd_tc_multi(items, T, B, H, D, zero_row=True, p=0.0):
    for it in items:
        assert it['dout'].stride(0) == it['out'].stride(0) and it['dq'].stride(0) == it['q'].stride(0)
        if it['S'] > 0:
            assert it['dk'].stride(0) == it['k'].stride(0) and it['dv'].stride(0) == it['k'].stride(0)
            assert it['dk'].dtype == it['k'].dtype
    arr = _attn_ctx_array(items, H * D)
    _lib.call('tt_attn_bwd_tc_multi', arr, c_int(len(items)), c_int(T), c_int(B), c_int(H), c_int(D),
              c_int(1 if zero_row else 0), c_float(p), c_int(1 if _attn_kv16(items) else 0), _stream())


def attn_decode_hm_multi(items, B, H, D, zero_row=True):
    """T = 1 step over head-major bf16 caches (k, v = [B,H,S,64]) for up to 4 contexts in one launch;
    a context with S = 0 attends to its bias / zero rows only."""
    arr = _attn_ctx_array(items, H * D)
    _lib.call('tt_attn_decode_hm_multi', arr, c_int(len(items)), c_int(B), c_int(H), c_int(D),
              c_int(1 if zero_row else 0), _stream())


def kv_repack_heads(k, v, S, B, H, D):
    """Token-major bf16 k, v views ([S*B, H*D], shared row stride) -> head-major K, V [B,H,S,D]."""
    _check_cuda(k, v)
    assert k.dtype == torch.bfloat16 and v.dtype == torch.bfloat16 and k.stride(0) == v.stride(0)
    ko = torch.empty((B, H, S, D), dtype=torch.bfloat16, device=k.device)
    vo = torch.empty_like(ko)
    _lib.call('tt_kv_repack_heads', _ptr(k), _ptr(v), c_ll(k.stride(0)), _ptr(ko), _ptr(vo), c_int(S),
              c_int(B), c_int(H), c_int(D), _stream())
    return ko, vo


def attn_decode_hm(q, k_hm, v_hm, bias_k, bias_v, mask, out, zero_row=True):
    """T = 1 attention over the head-major decode cache; q/out are [B, H*D] views (row-strided)."""
    B, H, S, D = k_hm.shape
    _lib.call('tt_attn_decode_hm', _ptr(q), _ptr(k_hm), _ptr(v_hm), _ptr(bias_k), _ptr(bias_v),
              _ptr(mask), _ptr(out), _ptr(None), c_int(B), c_int(S), c_int(H), c_int(D),
              c_ll(q.stride(0)), c_ll(out.stride(0)), c_int(1 if zero_row else 0), _stream())
    return out


def attn_avg_weights(q, k, bias_k, mask, lse, T, B, S, H, D, zero_row=True):
    L = S + (1 if bias_k is not None else 0) + (1 if zero_row else 0)
    w = torch.zeros((B, T, L), dtype=torch.float32, device=q.device)
    E = H * D
    _lib.call('tt_attn_avg_weights', _ptr(q), _ptr(k if S > 0 else None), _ptr(bias_k), _ptr(mask),
              _ptr(lse), _ptr(w), c_int(T), c_int(B), c_int(S), c_int(H), c_int(D),
              c_ll(q.stride(0)), c_ll(k.stride(0) if S > 0 else E), c_int(1 if zero_row else 0),
              _stream())
    return w


# --------------------------------------------------------------------------- adaptive softmax
def _int_array(vals):
    return (c_int * len(vals))(*vals)


def adaptive_prepare(target, cutoffs, pad_idx=1):
    """target int64 [N] -> head_target [N], tail_idx [n_tails,N], tail_local, tail_count, ntokens."""
    N = target.numel()
    nt = len(cutoffs) - 1
    dev = target.device
    head_t = torch.empty(N, dtype=torch.int32, device=dev)
    tail_idx = torch.zeros((max(nt, 1), N), dtype=torch.int32, device=dev)
    tail_local = torch.zeros((max(nt, 1), N), dtype=torch.int32, device=dev)
    tail_count = torch.zeros(max(nt, 1), dtype=torch.int32, device=dev)
    ntok = torch.zeros(1, dtype=torch.int32, device=dev)
    _lib.call('tt_adaptive_prepare', _ptr(target), c_int(N), _int_array(cutoffs),
              c_int(len(cutoffs)), c_int(pad_idx), _ptr(head_t), _ptr(tail_idx), _ptr(tail_local),
              _ptr(tail_count), _ptr(ntok), _stream())
    return head_t, tail_idx, tail_local, tail_count, ntok


def gather_rows(src, idx, count=None, cap=None):
    E = src.shape[1]
    cap = cap if cap is not None else idx.numel()
    dst = _f32(cap, E, like=src)
    _lib.call('tt_gather_rows', _ptr(src), _ptr(idx), _ptr(count), _ptr(dst), c_int(cap), c_int(E),
              _stream())
    return dst


def scatter_add_rows(src, idx, dst, count=None):
    cap, E = src.shape
    _lib.call('tt_scatter_add_rows', _ptr(src), _ptr(idx), _ptr(count), _ptr(dst), c_int(cap),
              c_int(E), _stream())
    return dst


def ce_fwd(logits, target, count=None, ignore_index=1, row_loss=None):
    M, V = logits.shape
    lse = _f32(M, like=logits)
    if row_loss is None:
        row_loss = _f32(M, like=logits)
    _lib.call('tt_ce_fwd', _ptr(logits), c_ll(logits.stride(0)), _ptr(target), _ptr(count),
              c_int(M), c_int(V), c_int(ignore_index), _ptr(lse), _ptr(row_loss), _stream())
    return lse, row_loss


def ce_bwd_(logits, target, lse, scale, count=None, ignore_index=1):
    M, V = logits.shape
    _lib.call('tt_ce_bwd', _ptr(logits), c_ll(logits.stride(0)), _ptr(target), _ptr(count),
              c_int(M), c_int(V), c_int(ignore_index), _ptr(lse), _ptr(scale), _stream())
    return logits


def ce_bwd16(logits, target, lse, scale, count=None, ignore_index=1, zero_round=0):
    """(softmax - onehot) * scale as a bf16 GEMM operand [M, V] (row pitch padded to 8 elements);
    logits stay intact.  zero_round > 0: rows >= round_up(count, zero_round) are left unwritten --
    only for consumers limited to `count` rows (m_limit / k_limit)."""
    M, V = logits.shape
    ld = (V + 7) // 8 * 8
    out = torch.empty((M, ld), dtype=torch.bfloat16, device=logits.device)[:, :V]
    _lib.call('tt_ce_bwd_bf16', _ptr(logits), c_ll(logits.stride(0)), _ptr(target), _ptr(count),
              c_int(M), c_int(V), c_int(ignore_index), _ptr(lse), _ptr(scale), _ptr(out),
              c_ll(ld), c_int(zero_round), _stream())
    return out


def loss_finalize(row_loss, ntokens):
    loss, scale = _f32(1, like=row_loss), _f32(1, like=row_loss)
    _lib.call('tt_loss_finalize', _ptr(row_loss), c_ll(row_loss.numel()), _ptr(ntokens),
              _ptr(loss), _ptr(scale), _stream())
    return loss, scale


def adaptive_logprob(head, tails, cutoffs, want_logprobs=True, want_argmax=True):
    M = head.shape[0]
    nt = len(cutoffs) - 1
    lp = _f32(M, cutoffs[-1], like=head) if want_logprobs else None
    am = torch.empty(M, dtype=torch.int64, device=head.device) if want_argmax else None
    amlp = _f32(M, like=head) if want_argmax else None
    tp = (c_void_p * max(nt, 1))(*[t.data_ptr() for t in tails])
    tl = (c_ll * max(nt, 1))(*[t.stride(0) for t in tails])
    _lib.call('tt_adaptive_logprob', _ptr(head), c_ll(head.stride(0)), tp, tl,
              _int_array(cutoffs), c_int(len(cutoffs)), c_int(M), _ptr(lp), _ptr(am), _ptr(amlp),
              _stream())
    return lp, am, amlp


# --------------------------------------------------------------------------- embedding
def embed_gather(ids, cutoffs, tables, E, tbc=True):
    B, T = ids.shape
    nb = len(cutoffs)
    out = _f32(B * T, nb * E, like=tables[0])
    tp = (c_void_p * nb)(*[t.data_ptr() for t in tables])
    _lib.call('tt_embed_gather', _ptr(ids), c_int(B), c_int(T), c_int(1 if tbc else 0),
              _int_array(cutoffs), c_int(nb), tp, c_int(E), _ptr(out), _stream())
    return out


def embed_scatter_grad(ids, cutoffs, grads, E, dA, padding_idx=0, tbc=True):
    B, T = ids.shape
    nb = len(cutoffs)
    gp = (c_void_p * nb)(*[(g.data_ptr() if g is not None else 0) for g in grads])
    _lib.call('tt_embed_scatter_grad', _ptr(ids), c_int(B), c_int(T), c_int(1 if tbc else 0),
              _int_array(cutoffs), c_int(nb), gp, c_int(E), c_int(padding_idx), _ptr(dA), _stream())


def make_positions(ids, pad=1, left_pad=False, start_pos=0, tbc=False):
    """start_pos: host int, or an int32 device tensor [1] (running position of a captured decode
    step)."""
    B, T = ids.shape
    pos = torch.empty((T, B) if tbc else (B, T), dtype=torch.int32, device=ids.device)
    if torch.is_tensor(start_pos):
        assert start_pos.dtype == torch.int32 and start_pos.is_cuda
        _lib.call('tt_make_positions_at', _ptr(ids), c_int(B), c_int(T), c_int(pad),
                  c_int(1 if left_pad else 0), c_int(0), _ptr(start_pos), c_int(1 if tbc else 0),
                  _ptr(pos), _stream())
        return pos
    _lib.call('tt_make_positions', _ptr(ids), c_int(B), c_int(T), c_int(pad),
              c_int(1 if left_pad else 0), c_int(start_pos), c_int(1 if tbc else 0), _ptr(pos),
              _stream())
    return pos


def transpose01(x):
    A, B, C = x.shape
    assert x.is_contiguous()
    out = torch.empty((B, A, C), dtype=torch.float32, device=x.device)
    _lib.call('tt_transpose01', _ptr(x), _ptr(out), c_int(A), c_int(B), c_int(C), _stream())
    return out


# --------------------------------------------------------------------------- frozen encoders
def im2col_nhwc(x, KH, KW, stride, pad):
    """x [B,H,W,C] bf16 -> ([B*Ho*Wo, Kp] bf16, Ho, Wo)."""
    B, H, W, C = x.shape
    Ho, Wo = (H + 2 * pad - KH) // stride + 1, (W + 2 * pad - KW) // stride + 1
    Kp = (KH * KW * C + 7) // 8 * 8
    out = torch.empty((B * Ho * Wo, Kp), dtype=torch.bfloat16, device=x.device)
    _lib.call('tt_im2col_nhwc', _ptr(x), _ptr(out), c_int(B), c_int(H), c_int(W), c_int(C),
              c_int(KH), c_int(KW), c_int(stride), c_int(pad), c_int(Kp), _stream())
    return out, Ho, Wo


def im2col_nhwc_bn(x, KH, KW, stride, pad, stats, n_stat, gamma, beta, eps, running_mean=None,
                   running_var=None, momentum=0.1, num_batches_tracked=None):
    """im2col of relu(batchnorm_train(x)) for the RAW convolution output x [B,H,W,C] bf16 whose
    per-channel sums over n_stat rows are in stats [2C]; -> ([B*Ho*Wo, Kp] bf16, Ho, Wo)."""
    _check_cuda(x, stats, gamma, beta)
    B, H, W, C = x.shape
    assert x.dtype == torch.bfloat16 and x.is_contiguous() and stats.numel() == 2 * C
    Ho, Wo = (H + 2 * pad - KH) // stride + 1, (W + 2 * pad - KW) // stride + 1
    Kp = (KH * KW * C + 7) // 8 * 8
    out = torch.empty((B * Ho * Wo, Kp), dtype=torch.bfloat16, device=x.device)
    _lib.call('tt_im2col_nhwc_bn', _ptr(x), _ptr(out), c_int(B), c_int(H), c_int(W), c_int(C),
              c_int(KH), c_int(KW), c_int(stride), c_int(pad), c_int(Kp), _ptr(stats), c_ll(n_stat),
              _ptr(gamma), _ptr(beta), c_float(eps),
              _ptr(running_mean) if running_mean is not None else None,
              _ptr(running_var) if running_var is not None else None, c_float(momentum),
              _ptr(num_batches_tracked) if num_batches_tracked is not None else None, _stream())
    return out, Ho, Wo


def im2col_nchw_f32(x, KH, KW, stride, pad, Kp):
    B, C, H, W = x.shape
    Ho, Wo = (H + 2 * pad - KH) // stride + 1, (W + 2 * pad - KW) // stride + 1
    out = torch.empty((B * Ho * Wo, Kp), dtype=torch.bfloat16, device=x.device)
    _lib.call('tt_im2col_nchw_f32', _ptr(x), _ptr(out), c_int(B), c_int(H), c_int(W), c_int(C),
              c_int(KH), c_int(KW), c_int(stride), c_int(pad), c_int(Kp), _stream())
    return out, Ho, Wo


def maxpool3x3s2_nhwc(x):
    B, H, W, C = x.shape
    Ho, Wo = (H + 2 - 3) // 2 + 1, (W + 2 - 3) // 2 + 1
    out = torch.empty((B, Ho, Wo, C), dtype=torch.bfloat16, device=x.device)
    _lib.call('tt_maxpool3x3s2_nhwc', _ptr(x), _ptr(out), c_int(B), c_int(H), c_int(W), c_int(C),
              _stream())
    return out


def bf16_to_f32(x):
    out = torch.empty(x.shape, dtype=torch.float32, device=x.device)
    _lib.call('tt_bf16_to_f32', _ptr(x), _ptr(out), c_ll(x.numel()), _stream())
    return out


def roberta_embed(ids, tok, pos, pad=1):
    B, S = ids.shape
    E = tok.shape[1]
    x = _f32(B * S, E, like=tok)
    is_pad = torch.empty(B * S, dtype=torch.uint8, device=ids.device)
    _lib.call('tt_roberta_embed', _ptr(ids), _ptr(tok), _ptr(pos), _ptr(x), _ptr(is_pad), c_int(B),
              c_int(S), c_int(E), c_int(pad), _stream())
    return x, is_pad


def ln_fwd16(x, gamma, beta, out16, row_zero=None, eps=1e-5):
    N, E = x.shape
    _lib.call('tt_ln_fwd16', _ptr(x), _ptr(gamma), _ptr(beta), _ptr(out16), _ptr(row_zero),
              c_int(N), c_int(E), c_float(eps), _stream())
    return out16


def flash_self_attn(qkv, mask, B, S, H, D):
    out = torch.empty((B * S, H * D), dtype=torch.bfloat16, device=qkv.device)
    _lib.call('tt_flash_self_attn', _ptr(qkv), _ptr(mask), _ptr(out), c_int(B), c_int(S), c_int(H),
              c_int(D), _stream())
    return out


def varlen_prepare(ids, pad=1):
    """ids [B,S] -> (inv_map int32 [B*S]: padded row -> packed row or -1, cu_seqlens int32 [B+1])."""
    B, S = ids.shape
    inv_map = torch.empty(B * S, dtype=torch.int32, device=ids.device)
    cu = torch.empty(B + 1, dtype=torch.int32, device=ids.device)
    _lib.call('tt_varlen_prepare', _ptr(ids), c_int(B), c_int(S), c_int(pad), _ptr(inv_map), _ptr(cu),
              _stream())
    return inv_map, cu


def flash_self_attn_varlen(qkv, cu, B, S_max, H, D, tc5=None):
    """Packed self-attention.  tc5 (default config.flash_tc5): the tcgen05/TMEM kernel, which needs
    finite values in every allocated row of qkv; otherwise the mma.sync kernel."""
    from . import config
    out = torch.empty((qkv.shape[0], H * D), dtype=torch.bfloat16, device=qkv.device)
    use_tc5 = config.flash_tc5 if tc5 is None else tc5
    if use_tc5 and qkv.shape[0] >= 128 and D == 64:     # TMA boxes are 128 rows tall
        assert qkv.is_contiguous() and qkv.shape[1] == 3 * H * D
        _lib.call('tt_flash_self_attn_varlen_tc5', _ptr(qkv), _ptr(cu), _ptr(out), c_int(B),
                  c_int(S_max), c_int(H), c_int(D), c_ll(qkv.shape[0]), _stream())
        return out
    _lib.call('tt_flash_self_attn_varlen', _ptr(qkv), _ptr(cu), _ptr(out), c_int(B), c_int(S_max),
              c_int(H), c_int(D), _stream())
    return out


def ln_fwd16_varlen(x, gamma, beta, y_packed=None, y_padded=None, inv_map=None, count=None,
                    x_packed=True, eps=1e-5):
    """LayerNorm fp32 -> bf16 between the packed and padded layouts (tt_ln_fwd16_varlen)."""
    E = x.shape[1]
    R = inv_map.numel() if inv_map is not None else x.shape[0]
    _lib.call('tt_ln_fwd16_varlen', _ptr(x), c_int(1 if x.dtype == torch.bfloat16 else 0),
              c_int(1 if x_packed else 0), _ptr(gamma), _ptr(beta),
              _ptr(y_packed), _ptr(y_padded), _ptr(inv_map), _ptr(count), c_int(R), c_int(E),
              c_float(eps), _stream())



# --------------------------------------------------------------------------- face encoders (convnets.cu)
def conv_out_size(n, k, stride, pad, ceil_mode=False):
    return int(_lib.lib().tt_conv_out_size(c_int(n), c_int(k), c_int(stride), c_int(pad),
                                           c_int(1 if ceil_mode else 0)))


def _nhwc_view(x):
    """x: [B,H,W,C] bf16 view whose last dim is contiguous and whose pixels are evenly pitched
    (a channel slice of a contiguous [B,H,W,Ctot] buffer).  Returns the pixel pitch."""
    B, H, W, C = x.shape
    pitch = x.stride(2)
    assert x.dtype == torch.bfloat16 and x.stride(3) == 1
    assert x.stride(1) == W * pitch and (B == 1 or x.stride(0) == H * W * pitch), x.stride()
    return pitch


def im2col_nhwc_hw(x, KH, KW, stride, pad_h, pad_w):
    """x [B,H,W,C] bf16 (channel-slice views allowed) -> ([B*Ho*Wo, Kp] bf16, Ho, Wo)."""
    _check_cuda(x)
    B, H, W, C = x.shape
    pitch = _nhwc_view(x)
    Ho, Wo = (H + 2 * pad_h - KH) // stride + 1, (W + 2 * pad_w - KW) // stride + 1
    Kp = (KH * KW * C + 7) // 8 * 8
    out = torch.empty((B * Ho * Wo, Kp), dtype=torch.bfloat16, device=x.device)
    _lib.call('tt_im2col_nhwc_hw', _ptr(x), c_ll(pitch), _ptr(out), c_int(B), c_int(H), c_int(W),
              c_int(C), c_int(KH), c_int(KW), c_int(stride), c_int(pad_h), c_int(pad_w), c_int(Kp),
              _stream())
    return out, Ho, Wo


def maxpool_nhwc(x, k, stride, pad=0, ceil_mode=False, out=None):
    """nn.MaxPool2d on NHWC bf16; `out` may be a channel slice of a concatenation buffer."""
    _check_cuda(x, out)
    B, H, W, C = x.shape
    Ho, Wo = conv_out_size(H, k, stride, pad, ceil_mode), conv_out_size(W, k, stride, pad, ceil_mode)
    if out is None:
        out = torch.empty((B, Ho, Wo, C), dtype=torch.bfloat16, device=x.device)
    assert out.shape == (B, Ho, Wo, C)
    _lib.call('tt_maxpool_nhwc', _ptr(x), c_ll(_nhwc_view(x)), _ptr(out), c_ll(_nhwc_view(out)),
              c_int(B), c_int(H), c_int(W), c_int(C), c_int(k), c_int(stride), c_int(pad),
              c_int(1 if ceil_mode else 0), _stream())
    return out


def avgpool_nhwc(x):
    """[B,H,W,C] bf16 contiguous -> [B,C] fp32."""
    _check_cuda(x)
    B, H, W, C = x.shape
    assert x.is_contiguous() and x.dtype == torch.bfloat16
    out = torch.empty((B, C), dtype=torch.float32, device=x.device)
    _lib.call('tt_avgpool_nhwc', _ptr(x), _ptr(out), c_int(B), c_int(H * W), c_int(C), _stream())
    return out


def prelu_bf16_(x2d, slope):
    """In-place PReLU on bf16 rows [R, C] (row pitch = x2d.stride(0))."""
    _check_cuda(x2d, slope)
    assert x2d.dtype == torch.bfloat16 and x2d.dim() == 2 and x2d.stride(1) == 1
    _lib.call('tt_prelu_bf16', _ptr(x2d), c_ll(x2d.stride(0)), _ptr(slope), c_ll(x2d.shape[0]),
              c_int(x2d.shape[1]), _stream())
    return x2d


def l2norm_rows(x, eps=1e-12):
    _check_cuda(x)
    assert x.dtype == torch.float32 and x.dim() == 2 and x.is_contiguous()
    y = torch.empty_like(x)
    _lib.call('tt_l2norm_rows', _ptr(x), _ptr(y), c_int(x.shape[0]), c_int(x.shape[1]), c_float(eps),
              _stream())
    return y


def softmax2_(x, c0=0):
    _check_cuda(x)
    assert x.dtype == torch.float32 and x.dim() == 2 and x.stride(1) == 1
    _lib.call('tt_softmax2', _ptr(x), c_ll(x.stride(0)), c_ll(x.shape[0]), c_int(c0), _stream())
    return x


def bn_stats(x2d, stats):
    """Per-channel sum / sum of squares of bf16 rows [M, C] accumulated into fp32 stats [2*C]."""
    _check_cuda(x2d, stats)
    assert x2d.dtype == torch.bfloat16 and x2d.dim() == 2 and x2d.stride(1) == 1
    assert stats.dtype == torch.float32 and stats.numel() == 2 * x2d.shape[1]
    _lib.call('tt_bn_stats_bf16', _ptr(x2d), c_ll(x2d.stride(0)), c_ll(x2d.shape[0]),
              c_int(x2d.shape[1]), _ptr(stats), _stream())
    return stats


def bn_apply_(x2d, stats, gamma, beta, eps, residual=None, relu=True, running_mean=None,
              running_var=None, momentum=0.1, num_batches_tracked=None):
    """Train-mode BatchNorm (+ identity, + ReLU) in place on bf16 rows [M, C] from bn_stats' sums."""
    _check_cuda(x2d, stats, gamma, beta)
    assert x2d.dtype == torch.bfloat16 and x2d.dim() == 2 and x2d.stride(1) == 1
    if residual is not None:
        assert residual.dtype == torch.bfloat16 and residual.shape == x2d.shape and residual.stride(1) == 1
    _lib.call('tt_bn_apply_bf16', _ptr(x2d), c_ll(x2d.stride(0)), c_ll(x2d.shape[0]), c_int(x2d.shape[1]),
              _ptr(stats), _ptr(gamma), _ptr(beta), c_float(eps),
              _ptr(residual) if residual is not None else None,
              c_ll(residual.stride(0) if residual is not None else 0), c_int(1 if relu else 0),
              _ptr(running_mean) if running_mean is not None else None,
              _ptr(running_var) if running_var is not None else None, c_float(momentum),
              _ptr(num_batches_tracked) if num_batches_tracked is not None else None, _stream())
    return x2d
