"""Parity of the tcgen05 GEMM (tt_gemm_bf16_tn) and the bf16 operand cast against torch fp32."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref(a16, b16, bias, residual, alpha, act):
    # act(alpha * (A.B^T + bias) + residual)
    c = a16.float() @ b16.float().t()
    if bias is not None:
        c = c + bias
    c = alpha * c
    if residual is not None:
        c = c + residual
    if act == 1:
        c = torch.relu(c)
    elif act == 2:
        c = torch.nn.functional.gelu(c)
    return c


@pytest.mark.parametrize('M,N,K', [
    (128, 128, 64), (128, 256, 128), (800, 1024, 1024), (8192, 1024, 1024),
    (800, 2048, 1024), (800, 1024, 4096), (100, 48, 1024), (77, 5002, 1024),
    (8192, 3072, 1024), (1, 1024, 1024), (300, 496, 1000), (4096, 4096, 512),
    (33, 40, 8), (2048, 30265, 1024)])
def test_gemm_plain(M, N, K):
    from tell_b200 import ops
    torch.manual_seed(M * 31 + N * 7 + K)
    a = torch.randn(M, K, device='cuda').bfloat16()
    b = torch.randn(N, K, device='cuda').bfloat16()
    c = ops.gemm_tn(a, b)
    torch.cuda.synchronize()
    ref = _ref(a, b, None, None, 1.0, 0)
    err = (c - ref).abs().max().item()
    assert err <= 2e-3 * max(1.0, ref.abs().max().item()), (M, N, K, err)


@pytest.mark.parametrize('act', [0, 1, 2])
def test_gemm_epilogue(act):
    from tell_b200 import ops
    torch.manual_seed(act)
    M, N, K = 800, 1024, 1024
    a = (torch.randn(M, K, device='cuda') / 32).bfloat16()
    b = torch.randn(N, K, device='cuda').bfloat16()
    bias = torch.randn(N, device='cuda')
    res = torch.randn(M, N, device='cuda')
    c, c16 = ops.gemm_tn(a, b, bias=bias, residual=res, alpha=0.5, act=act, want16=True)
    ref = _ref(a, b, bias, res, 0.5, act)
    assert (c - ref).abs().max().item() < 2e-3
    assert (c16.float() - ref).abs().max().item() < 5e-2
    # accumulate into C
    c2 = c.clone()
    ops.gemm_tn(a, b, out=c2, accumulate=True)
    ref2 = ref + _ref(a, b, None, None, 1.0, 0)
    assert (c2 - ref2).abs().max().item() < 4e-3
    # bf16 residual (ResNet identity branch)
    c3 = ops.gemm_tn(a, b, bias=bias, residual16=res.bfloat16(), act=act)
    ref3 = _ref(a, b, bias, res.bfloat16().float(), 1.0, act)
    assert (c3 - ref3).abs().max().item() < 2e-3


def test_gemm_unaligned_output_and_m_limit():
    from tell_b200 import ops
    torch.manual_seed(5)
    M, N, K = 300, 50, 72
    a = torch.randn(M, K, device='cuda').bfloat16()
    b = torch.randn(N, K, device='cuda').bfloat16()
    out = torch.full((M, N), 7.0, device='cuda')
    lim = torch.tensor([130], dtype=torch.int32, device='cuda')
    ops.gemm_tn(a, b, out=out, m_limit=lim)
    ref = _ref(a, b, None, None, 1.0, 0)
    assert (out[:130] - ref[:130]).abs().max().item() < 2e-3
    assert (out[130:] == 7.0).all()


@pytest.mark.parametrize('transpose', [False, True])
@pytest.mark.parametrize('split', [0, 1, 2])
def test_cast(transpose, split):
    from tell_b200 import ops
    torch.manual_seed(3)
    x = torch.randn(70, 100, device='cuda')
    y = ops.cast_bf16(x, transpose=transpose, split=split)
    src = x.t() if transpose else x
    hi = src.bfloat16()
    lo = (src - hi.float()).bfloat16()
    if split == 0:
        exp = hi
    elif split == 1:
        exp = torch.cat([hi, lo, hi], 1)
    else:
        exp = torch.cat([hi, hi, lo], 1)
    assert torch.equal(y, exp)


@pytest.mark.parametrize('cols', [102, 101, 7, 3])
def test_cast_odd_width_padded_pitch(cols):
    """fp32 rows with a 16-byte-aligned pitch but a width that is not a multiple of 4 (the 5002-way
    adaptive-softmax head): vector body + per-row leftover columns, exact bf16 rounding."""
    from tell_b200 import ops
    torch.manual_seed(cols)
    x = ops.f32_padded(70, cols, torch.zeros(1, device='cuda'))
    x.copy_(torch.randn(70, cols, device='cuda'))
    y = ops.cast_bf16(x)
    assert torch.equal(y, x.bfloat16())
    assert (y.as_strided((70, y.stride(0)), (y.stride(0), 1))[:, cols:] == 0).all()


def test_gemm_split_precision():
    """bf16x3 operands recover ~fp32 accuracy (parity mode)."""
    from tell_b200 import ops
    torch.manual_seed(11)
    M, N, K = 256, 512, 1024
    x = torch.randn(M, K, device='cuda')
    w = torch.randn(N, K, device='cuda') / 32
    a = ops.cast_bf16(x, split=1)
    b = ops.cast_bf16(w, split=2)
    c = ops.gemm_tn(a, b)
    ref = (x.double() @ w.double().t()).float()
    err = (c - ref).abs().max().item()
    assert err < 2e-4, err
    c1 = ops.gemm_tn(ops.cast_bf16(x), ops.cast_bf16(w))
    assert (c1 - ref).abs().max().item() > err


@pytest.mark.parametrize('M,N,K', [(128, 128, 64), (800, 1024, 1024), (1024, 4096, 800),
                                   (300, 200, 136), (64, 64, 8192), (2048, 30265, 800), (50, 40, 24)])
@pytest.mark.parametrize('ta,tb', [(True, False), (False, True), (True, True)])
def test_gemm_mn_major_operands(M, N, K, ta, tb):
    """trans_a / trans_b: the kernel reads [K,M] / [K,N] storage through MN-major UMMA descriptors."""
    from tell_b200 import ops
    torch.manual_seed(M + N + K)

    def pad8(t):
        r, c = t.shape
        buf = torch.zeros(r, (c + 7) // 8 * 8, device='cuda', dtype=torch.bfloat16)
        buf[:, :c] = t
        return buf[:, :c]
    a = torch.randn(M, K, device='cuda').bfloat16()
    b = torch.randn(N, K, device='cuda').bfloat16()
    a_st = pad8(a.t().contiguous()) if ta else pad8(a)
    b_st = pad8(b.t().contiguous()) if tb else pad8(b)
    c = ops.gemm_tn(a_st, b_st, trans_a=ta, trans_b=tb)
    ref = a.float() @ b.float().t()
    err = (c - ref).abs().max().item()
    assert err <= 2e-3 * max(1.0, ref.abs().max().item()), (M, N, K, ta, tb, err)


@pytest.mark.parametrize('M,N,K,lim', [(256, 1024, 4096, 256), (800, 1024, 30265, 23), (130, 256, 2104, 97),
                                       (800, 1024, 15000, 300)])
def test_gemm_split_k(M, N, K, lim):
    """Row-limited, few-tile / long-K problems that carry a row-count hint run as split-K (partials
    added atomically into a zeroed C): checked against a float64 reference, with bias + residual
    entering exactly once, and with a device row limit smaller AND larger than the hint."""
    from tell_b200 import ops
    torch.manual_seed(K)
    a32 = (torch.randn(M, K) * 0.1).to(torch.bfloat16).float()
    a16 = ops.cast_bf16(a32.cuda())                          # row pitch padded to a multiple of 8
    a = a32
    w = (torch.randn(K, N) * 0.1).to(torch.bfloat16)        # stored [K, N]: MN-major B (trans_b)
    bias = torch.randn(N)
    res = torch.randn(M, N)
    want = (a.double() @ w.double() + bias.double()) * 0.5 + res.double()
    kw = {}
    if lim is not None:
        kw = dict(m_limit=torch.tensor([lim], dtype=torch.int32, device='cuda'), m_hint=128)
    out = ops.gemm_tn(a16, w.cuda(), bias=bias.cuda(), residual=res.cuda(), alpha=0.5, trans_b=True, **kw)
    rows = M if lim is None else lim
    err = (out[:rows].cpu().double() - want[:rows]).abs().max().item()
    assert err < 2e-3 * max(1.0, want.abs().max().item()), err
    if lim is not None:
        assert torch.equal(out[lim:].cpu(), torch.zeros(M - lim, N))


def _staged(on):
    from tell_b200 import _lib
    _lib.lib().tt_gemm_set_staged_epilogue(1 if on else 0)


@pytest.mark.parametrize('M,N,K,bias,res,act,alpha', [
    (8192, 3072, 1024, True, False, 1, 1.0),     # CTA-pair kernel, 256-wide tiles, bias + ReLU
    (8192, 4096, 1024, True, False, 2, 1.0),     # RoBERTa fc1: GELU
    (8192, 1024, 4096, True, True, 0, 1.0),      # RoBERTa fc2: bf16 residual
    (3136, 1024, 256, True, True, 1, 1.0),       # ResNet stage-3 expanding 1x1 conv + identity + ReLU
    (12544, 512, 128, True, True, 1, 1.0),
    (50176, 64, 64, True, False, 1, 1.0),        # N < every tile width except 64
    (800, 1024, 1024, False, False, 0, 0.5),     # decoder-sized, alpha
    (777, 96, 200, True, True, 1, 1.0),          # ragged M (last warp rows take the register path), N = 3 chunks
    (40, 32, 64, False, True, 0, 1.0),           # one chunk, one staged warp + one ragged warp
    (5000, 2048, 512, True, True, 2, 1.0)])
def test_gemm_staged_epilogue_bit_identical_to_register_epilogue(M, N, K, bias, res, act, alpha):
    """The staged epilogue (shared-memory tile + TMA store, bf16 residual by TMA load) must produce
    exactly the bits of the row-per-thread register epilogue, and both must match torch fp32."""
    from tell_b200 import ops
    torch.manual_seed(M + N + K)
    a = (torch.randn(M, K, device='cuda') / 8).bfloat16()
    b = (torch.randn(N, K, device='cuda') / 4).bfloat16()
    bv = torch.randn(N, device='cuda') if bias else None
    r16 = torch.randn(M, N, device='cuda').bfloat16() if res else None
    outs = []
    try:
        # the staged path runs several times: an ordering bug between tcgen05.ld and wait::ld once
        # showed up as stale 16-byte pieces in ~1 of 8 launches of the epilogue-bound shapes
        for on in (False, True, True, True, True):
            _staged(on)
            c16 = torch.full((M, N), 7.0, dtype=torch.bfloat16, device='cuda')
            ops.gemm_tn(a, b, out16=c16, bias=bv, residual16=r16, act=act, alpha=alpha, want32=False)
            torch.cuda.synchronize()
            outs.append(c16)
    finally:
        _staged(True)
    for o in outs[1:]:
        assert torch.equal(outs[0], o)
    ref = _ref(a, b, bv, r16.float() if res else None, alpha, act)
    err = (outs[1].float() - ref).abs().max().item()
    assert err <= 1.2e-2 * max(1.0, ref.abs().max().item()), err


def test_gemm_staged_epilogue_row_limit_and_strided_output():
    """Device-side row limit: rows >= m_limit keep their previous contents (the ragged 32-row group
    takes the register path); the output may be a column slice of a wider buffer (ldc16 > N)."""
    from tell_b200 import ops
    torch.manual_seed(5)
    M, N, K = 4096, 1024, 512
    a = (torch.randn(M, K, device='cuda') / 8).bfloat16()
    b = (torch.randn(N, K, device='cuda') / 4).bfloat16()
    bias = torch.randn(N, device='cuda')
    r16 = torch.randn(M, N, device='cuda').bfloat16()
    for lim in (4096, 2977, 1, 64):
        m_limit = torch.tensor([lim], dtype=torch.int32, device='cuda')
        wide = torch.full((M, 2 * N + 64), 3.0, dtype=torch.bfloat16, device='cuda')
        out = wide[:, 64:64 + N]
        ops.gemm_tn(a, b, out16=out, bias=bias, residual16=r16, act=1, want32=False, m_limit=m_limit,
                    m_hint=lim)
        torch.cuda.synchronize()
        ref = _ref(a[:lim], b, bias, r16[:lim].float(), 1.0, 1)
        assert (out[:lim].float() - ref).abs().max().item() <= 1.2e-2 * max(1.0, ref.abs().max().item())
        assert (out[lim:] == 3.0).all()
        assert (wide[:, :64] == 3.0).all() and (wide[:, 64 + N:] == 3.0).all()


@pytest.mark.parametrize('M,N,K', [(3136, 256, 1024), (3136, 1024, 256), (98, 2048, 512), (12544, 128, 512),
                                   (784, 512, 2048), (50176, 64, 64), (8192, 3072, 1024), (45, 64, 72)])
def test_gemm_column_statistics_in_the_epilogue(M, N, K):
    """col_stats: per-column sum and sum of squares of the bf16-ROUNDED outputs, accumulated by the
    GEMM epilogue (train-mode BatchNorm behind a convolution); full 32-row groups take the staged
    path (shared-memory tile), ragged last rows the register path."""
    from tell_b200 import ops
    torch.manual_seed(M + N)
    a = (torch.randn(M, K, device='cuda') / 8).bfloat16()
    b = (torch.randn(N, K, device='cuda') / 4).bfloat16()
    st = torch.zeros(2 * N, device='cuda')
    y = ops.gemm_tn(a, b, want32=False, want16=True, col_stats=st)
    y_plain = ops.gemm_tn(a, b, want32=False, want16=True)
    torch.cuda.synchronize()
    assert torch.equal(y, y_plain)
    yd = y.double()
    s_ref, q_ref = yd.sum(0), (yd * yd).sum(0)
    assert (st[:N].double() - s_ref).abs().max().item() <= 1e-4 * max(1.0, yd.abs().sum(0).max().item())
    assert ((st[N:].double() - q_ref).abs() / q_ref.clamp_min(1e-6)).max().item() < 1e-4
    # accumulates: a second launch doubles the sums
    ops.gemm_tn(a, b, want32=False, want16=True, col_stats=st)
    assert ((st[N:].double() - 2 * q_ref).abs() / q_ref.clamp_min(1e-6)).max().item() < 1e-4


@pytest.mark.parametrize('nb,M,E', [(4, 800, 1024), (4, 256, 1024), (3, 77, 64), (2, 130, 256)])
def test_gemm_batched_launch_equals_separate_launches(nb, M, E):
    """nbatch same-shape problems in ONE launch (the four out-projections of a decoder layer, their
    dX and dW GEMMs): forward layout (A column blocks, stacked weights), dX (trans_b) and dW
    (trans_a, trans_b) each bit-identical to nb separate launches."""
    from tell_b200 import ops
    torch.manual_seed(nb * 1000 + M)
    a = (torch.randn(M, nb * E, device='cuda') / 8).bfloat16()
    w = (torch.randn(nb * E, E, device='cuda') / 8).bfloat16()
    bias = torch.randn(nb * E, device='cuda')
    d = (torch.randn(M, nb * E, device='cuda') / 8).bfloat16()
    # forward: out[z] = a[:, zE:(z+1)E] @ w[zE:(z+1)E]^T + bias[zE:(z+1)E]
    out = torch.empty(nb, M, E, device='cuda')
    ops.gemm_tn_batched(a, w, nb, M, E, E, out.view(nb * M, E), M * E, a_off=(E, 0), b_off=(0, E), bias=bias,
                        bias_off=E)
    for z in range(nb):
        ref = ops.gemm_tn(a[:, z * E:(z + 1) * E], w[z * E:(z + 1) * E], bias=bias[z * E:(z + 1) * E].contiguous())
        assert torch.equal(out[z], ref), ('fwd', z)
    # dX: dx[:, zE:(z+1)E] = d[:, zE:(z+1)E] @ w[zE:(z+1)E]      (B stored [K, N])
    dx = torch.empty(M, nb * E, device='cuda')
    ops.gemm_tn_batched(d, w, nb, M, E, E, dx, E, a_off=(E, 0), b_off=(0, E), trans_b=True)
    for z in range(nb):
        ref = ops.gemm_tn(d[:, z * E:(z + 1) * E], w[z * E:(z + 1) * E], trans_b=True)
        assert torch.equal(dx[:, z * E:(z + 1) * E], ref), ('dx', z)
    # dW: dw[zE:(z+1)E] = d[:, zE:(z+1)E]^T @ a[:, zE:(z+1)E]      (both stored [K = rows, .])
    dw = torch.empty(nb * E, E, device='cuda')
    ops.gemm_tn_batched(d, a, nb, E, E, M, dw, E * E, a_off=(E, 0), b_off=(E, 0), trans_a=True, trans_b=True)
    for z in range(nb):
        ref = ops.gemm_tn(d[:, z * E:(z + 1) * E], a[:, z * E:(z + 1) * E], trans_a=True, trans_b=True)
        assert torch.equal(dw[z * E:(z + 1) * E], ref), ('dw', z)
