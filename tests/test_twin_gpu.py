"""bf16 operand twins (csrc/twin.cu, tell_b200/twin.py): every twin kernel against the single-output
kernel it extends (fp32 outputs bit-identical, bf16 outputs = round-to-nearest of them), the
multi-context LayerNorm launches against per-context launches, and the decoder with twins on against
twins off (same bf16 operands reach every GEMM, so the forward is bit-identical)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _r(*shape, seed=0, scale=1.0):
    g = torch.Generator(device='cuda').manual_seed(seed)
    return torch.randn(*shape, generator=g, device='cuda') * scale


def _bf(x):
    return x.to(torch.bfloat16)


def test_elementwise_twins_match_single_output_kernels():
    from tell_b200 import ops
    x = _r(800, 4096, seed=1)
    y0 = ops.dropout(x, 0.3, 1234)
    y, y16 = ops.dropout_tw(x, 0.3, 1234)
    assert torch.equal(y, y0) and torch.equal(y16, _bf(y0))
    _, y16b = ops.dropout_tw(x, 0.3, 1234, want32=False)
    assert torch.equal(y16b, y16)
    act = _r(800, 4096, seed=2).relu()
    d0 = ops.relu_bwd(x, act)
    d, d16 = ops.relu_bwd_tw(x, act)
    assert torch.equal(d, d0) and torch.equal(d16, _bf(d0))
    h = _r(800, 2048, seed=3)
    g0 = ops.glu_fwd(h)
    g, g16 = ops.glu_fwd_tw(h)
    assert torch.equal(g, g0) and torch.equal(g16, _bf(g0))
    dout = _r(800, 1024, seed=4)
    dh0 = ops.glu_bwd(dout, h)
    dh, dh16 = ops.glu_bwd_tw(dout, h)
    assert torch.equal(dh, dh0) and torch.equal(dh16, _bf(dh0))


@pytest.mark.parametrize('E,n,p', [(1024, 1, 0.0), (1024, 4, 0.1), (1024, 3, 0.0), (64, 4, 0.1), (256, 2, 0.0)])
def test_layernorm_multi_matches_per_context_launches(E, n, p):
    from tell_b200 import ops
    N = 800 if E == 1024 else 27
    res = _r(N, E, seed=10)
    hs = [_r(N, E, seed=11 + c) for c in range(n)]
    gam = [1.0 + 0.1 * _r(E, seed=20 + c) for c in range(n)]
    bet = [0.1 * _r(E, seed=30 + c) for c in range(n)]
    seeds = [100 + c for c in range(n)]
    # reference: one launch per context into column blocks
    Y0 = torch.empty(N, n * E, device='cuda')
    h0 = [h.clone() for h in hs]
    st0 = []
    for c in range(n):
        _, m, r = ops.ln_fwd(h0[c], res, gam[c], bet[c], 1e-5, p, seeds[c], out=Y0[:, c * E:(c + 1) * E])
        st0.append((m, r))
    h1 = [h.clone() for h in hs]
    Y, Y16, means, rstds = ops.ln_fwd_multi(h1, res, gam, bet, 1e-5, p, seeds)
    exact = E == 1024            # same kernel arithmetic only at the production width
    for c in range(n):
        if exact:
            assert torch.equal(h1[c], h0[c])
            assert torch.equal(means[c], st0[c][0]) and torch.equal(rstds[c], st0[c][1])
        else:
            assert torch.allclose(h1[c], h0[c], atol=1e-6)
            assert torch.allclose(means[c], st0[c][0], atol=1e-5) and torch.allclose(rstds[c], st0[c][1], rtol=1e-4)
    if exact:
        assert torch.equal(Y, Y0)
    else:
        assert torch.allclose(Y, Y0, atol=2e-5)
    assert torch.equal(Y16, _bf(Y))
    # ---- backward
    dY = _r(N, n * E, seed=40)
    dX0 = None
    dh0, dg0, db0 = [], [], []
    for c in range(n):
        dg, db = torch.zeros(E, device='cuda'), torch.zeros(E, device='cuda')
        dx, dh = ops.ln_bwd(dY[:, c * E:(c + 1) * E], h1[c], means[c], rstds[c], gam[c], p, seeds[c],
                            dgamma=dg, dbeta=db)
        dX0 = dx.clone() if dX0 is None else dX0 + dx
        dh0.append(dh)
        dg0.append(dg)
        db0.append(db)
    dgs = [torch.zeros(E, device='cuda') for _ in range(n)]
    dbs = [torch.zeros(E, device='cuda') for _ in range(n)]
    dX, dhs, dh16 = ops.ln_bwd_multi(dY, h1, means, rstds, gam, p, seeds, want_dx=True, want_dh32=True,
                                     want_dh16=True, dgammas=dgs, dbetas=dbs)
    assert torch.equal(dX, dX0)
    for c in range(n):
        assert torch.equal(dhs[c], dh0[c])
        assert torch.equal(dh16[:, c * E:(c + 1) * E], _bf(dh0[c]))
        assert torch.allclose(dgs[c], dg0[c], rtol=1e-4, atol=1e-4)      # atomics: order differs
        assert torch.allclose(dbs[c], db0[c], rtol=1e-4, atol=1e-4)
    # outputs are optional
    dX2, none32, none16 = ops.ln_bwd_multi(dY, h1, means, rstds, gam, p, seeds, want_dx=True,
                                           want_dh32=False, want_dh16=False)
    assert none32 is None and none16 is None and torch.equal(dX2, dX0)


@pytest.mark.parametrize('K', [3, 31])
def test_dynconv_twins(K):
    from tell_b200 import ops
    T, B, C, H = 50, 16, 1024, 16
    x = _r(T, B, C, seed=1)
    z = _r(T, B, H * K, seed=2)
    o0, p0 = ops.dynconv_fwd(x, z, H, K, True, 0.1, 77)
    o, p, o16 = ops.dynconv_fwd(x, z, H, K, True, 0.1, 77, twin=True)
    assert torch.equal(o, o0) and torch.equal(p, p0)
    assert o16.shape == (T * B, C) and torch.equal(o16, _bf(o0).view(T * B, C))
    dout = _r(T, B, C, seed=3)
    dx0, dz0 = ops.dynconv_bwd(dout, x, p0, H, K, True, 0.1, 77)
    dx, dz, dz16 = ops.dynconv_bwd(dout, x, p0, H, K, True, 0.1, 77, twin=True)
    assert torch.equal(dx, dx0) and torch.equal(dz, dz0)
    assert torch.equal(dz16, _bf(dz0).view(T * B, H * K))
    w0 = _r(K - 1, B, C, seed=4)
    w1 = w0.clone()
    xn = _r(B, C, seed=5)
    zs = _r(B, H * K, seed=6)
    s0 = ops.dynconv_step(w0, xn, zs, H, K, True)
    s1, s16 = ops.dynconv_step(w1, xn, zs, H, K, True, twin=True)
    assert torch.equal(s1, s0) and torch.equal(w1, w0) and torch.equal(s16, _bf(s0))


def test_twin_registry_rejects_stale_or_reshaped_tensors():
    from tell_b200 import twin
    twin.clear()
    a = _r(8, 16)
    a16 = _bf(a)
    twin.put(a, a16)
    assert twin.get(a) is a16
    assert twin.get(a.view(8, 16)) is a16            # another tensor object on the same memory
    assert twin.get(a.view(16, 8)) is None           # same pointer, other shape
    assert twin.get(a[:, :8]) is None                # same pointer, other strides / shape
    a.add_(1.0)                                      # modified after the twin was written
    assert twin.get(a) is None
    twin.clear()
    assert twin.get(a) is None


def _decoder_run(twins_on, train):
    from tell_b200 import config, synth, twin
    from tell_b200 import models as M
    from tell_b200.testing import build_decoder
    config.set_precision('bf16')
    old = config.twins
    config.twins = twins_on
    try:
        cfg = dict(synth.CFG_FULL, vocab=6000, cutoffs=(1000, 3000), kernels=(3, 15))
        sd = synth.decoder_state_dict(cfg, seed=0, logit_gain=4.0)
        dec = build_decoder(cfg, M.DynamicConvFacesObjectsDecoder, sd).cuda()
        dec.train(train)
        cap, ctx = synth.decoder_inputs(cfg, 4, 12, 40, 3, 5, 9, seed=7)
        inp, tgt = cap[:, :-1].contiguous().cuda(), cap[:, 1:].contiguous().cuda()
        cctx = {k: v.cuda() for k, v in ctx.items()}
        cctx['article'].requires_grad_(True)
        twin.hits = twin.misses = 0
        out, _ = dec({'roberta': inp}, cctx)
        loss, _ = dec.adaptive_softmax.fused_loss(out, tgt)
        loss.backward()
        grads = {k: p.grad.detach().clone() for k, p in dec.named_parameters() if p.grad is not None}
        return out.detach(), loss.detach(), cctx['article'].grad.detach(), grads, (twin.hits, twin.misses)
    finally:
        config.twins = old


def test_decoder_with_twins_equals_decoder_with_casts():
    """Eval mode (no dropout: the seeds of the two runs would otherwise have to line up): forward
    bit-identical, gradients equal up to the order of atomic column sums."""
    o1, l1, da1, g1, st = _decoder_run(True, False)
    o0, l0, da0, g0, _ = _decoder_run(False, False)
    assert st[0] >= 2 * 12, st          # the twins were actually used (12.5 hits per layer, fwd + bwd)
    assert torch.equal(o1, o0) and torch.equal(l1, l0)
    assert torch.allclose(da1, da0, rtol=1e-4, atol=1e-6 * da0.abs().max().item() + 1e-9)
    for k in g0:
        ref = g0[k]
        tol = 2e-4 * max(ref.abs().max().item(), 1e-6)
        assert (g1[k] - ref).abs().max().item() <= tol, k


def test_decoder_train_mode_with_twins_runs_and_uses_them():
    o1, l1, _, g1, st = _decoder_run(True, True)
    assert torch.isfinite(l1) and all(torch.isfinite(v).all() for v in g1.values())
    assert st[0] >= 2 * 12, st
