"""The weight bank (one tt_weight_prep launch per step, one tt_wnorm_bwd_multi per backward) must
reproduce the per-module path: same outputs, same loss, same gradients -- on the recording step
(weight norms banked, other operands cast per call) and on later steps (everything from the table).
Tolerance: the two paths round the same fp32 values to bf16; only the row-norm reduction order
differs (float4 vs scalar strides), so 1e-5 relative on outputs and 1e-4 on gradients."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'oracle'))
SHAPES = dict(B=3, T=9, S=11, F=3, O=4, P=5)


def _build(use_bank):
    from tell_b200 import config, synth
    from tell_b200.models import DynamicConvFacesObjectsDecoder
    from tell_b200.testing import build_decoder
    config.set_precision('bf16')
    cfg = synth.CFG_TINY
    sd = synth.decoder_state_dict(cfg, seed=0, logit_gain=4.0)
    dec = build_decoder(cfg, DynamicConvFacesObjectsDecoder, sd).cuda().eval()
    dec.use_weight_bank = use_bank
    cap, ctx = synth.decoder_inputs(cfg, **SHAPES, seed=1234)
    return dec, cap.cuda(), {k: v.cuda() for k, v in ctx.items()}


def _step(dec, cap, ctx):
    for p in dec.parameters():
        p.grad = None
    inp, tgt = cap[:, :-1].contiguous(), cap[:, 1:].contiguous()
    with dec.weight_scope():
        X, _ = dec.forward_tbc({'roberta': inp}, ctx)
        T, B, E = X.shape
        loss, _ = dec.adaptive_softmax.fused_loss(X.view(T * B, E), tgt.t().contiguous())
    loss.backward()
    return X.detach().clone(), loss.detach().clone(), \
        {n: p.grad.detach().clone() for n, p in dec.named_parameters() if p.grad is not None}


def _close(a, b, rel):
    return (a - b).abs().max().item() <= rel * max(1e-6, b.abs().max().item())


def test_bank_matches_per_module_path():
    ref, cap, ctx = _build(False)
    X0, l0, g0 = _step(ref, cap, ctx)
    dec, cap, ctx = _build(True)
    for step in range(3):      # step 0 records, steps 1-2 are served entirely from the table
        X, l, g = _step(dec, cap, ctx)
        assert _close(X, X0, 1e-5), step
        assert abs(l.item() - l0.item()) < 1e-5 * max(1.0, abs(l0.item())), step
        assert set(g) == set(g0)
        for n in g0:
            assert _close(g[n], g0[n], 1e-4), (step, n)
    bank = dec._bank
    assert bank.n_segs > len(bank.wn) and not bank.log          # recorded forms were finalised


def test_bank_sees_weight_updates():
    """An in-place parameter update between steps (what an optimizer does) must reach the GEMMs."""
    dec, cap, ctx = _build(True)
    _step(dec, cap, ctx)
    X1, _, _ = _step(dec, cap, ctx)
    with torch.no_grad():
        for p in dec.parameters():
            p.mul_(1.25)
    X2, _, _ = _step(dec, cap, ctx)
    ref, cap, ctx = _build(False)
    with torch.no_grad():
        for p in ref.parameters():
            p.mul_(1.25)
    X2r, _, _ = _step(ref, cap, ctx)
    assert not _close(X2, X1, 1e-3)
    assert _close(X2, X2r, 1e-5)


def test_zero_arena_gradients_match():
    """Small accumulate-into gradients served from the per-step zero arena (one memset per step)
    equal the torch.zeros path; colsum through the vector kernel equals a torch column sum."""
    from tell_b200 import config, ops
    ref, cap, ctx = _build(True)
    _, _, g0 = _step(ref, cap, ctx)
    dec, cap, ctx = _build(True)
    arena = config.enable_zero_arena('cuda')
    try:
        for step in range(2):
            arena.reset()
            _, _, g = _step(dec, cap, ctx)
            assert arena.off > 0
            for n in g0:
                assert _close(g[n], g0[n], 1e-4), (step, n)
    finally:
        config.disable_zero_arena()
    for (M, N) in [(800, 1024), (8192, 4096), (37, 20), (800, 30265), (5, 4)]:
        x = torch.randn(M, N, device='cuda')
        got = ops.colsum(x, scale=0.5)
        want = x.double().sum(0).float() * 0.5
        assert (got - want).abs().max().item() < 2e-3 * max(1.0, want.abs().max().item()), (M, N)


@pytest.mark.parametrize('level', [1, 2])
def test_wgrad_stream_gradients_match(level):
    """Weight gradients forked onto the second stream (functional.wgrad) equal the single-stream
    backward -- eagerly and when the step is captured in a CUDA graph (the fork becomes a parallel
    branch that must rejoin before the capture ends)."""
    from tell_b200 import config
    ref, cap, ctx = _build(True)
    _step(ref, cap, ctx)
    _, l0, g0 = _step(ref, cap, ctx)
    dec, cap, ctx = _build(True)
    config.enable_wgrad_stream(level)
    try:
        for step in range(3):
            _, l, g = _step(dec, cap, ctx)
            torch.cuda.synchronize()
            assert set(g) == set(g0)
            for n in g0:
                assert _close(g[n], g0[n], 1e-4), (step, n)
        # captured: replay twice, gradients land in the tensors produced during capture
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.stream(s):
            _step(dec, cap, ctx)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        held = {}
        with torch.cuda.graph(graph):
            for p in dec.parameters():
                p.grad = None
            inp, tgt = cap[:, :-1].contiguous(), cap[:, 1:].contiguous()
            with dec.weight_scope():
                X, _ = dec.forward_tbc({'roberta': inp}, ctx)
                T, B, E = X.shape
                loss, _ = dec.adaptive_softmax.fused_loss(X.view(T * B, E), tgt.t().contiguous())
            loss.backward()
            held = {n: p.grad for n, p in dec.named_parameters() if p.grad is not None}
        for _ in range(2):
            graph.replay()
        torch.cuda.synchronize()
        assert abs(loss.item() - l0.item()) < 1e-5 * max(1.0, abs(l0.item()))
        for n in g0:
            assert _close(held[n], g0[n], 1e-4), ('graph', n)
    finally:
        config.enable_wgrad_stream(0)


def test_bank_gradient_accumulation_over_two_backwards():
    """p.grad adopted from the bank's persistent dv / dg buffers must survive the next backward:
    two backwards without zeroing give g1 + g2 (== 2 g for the same batch), not 2 * g2."""
    dec, cap, ctx = _build(True)
    _step(dec, cap, ctx)                       # recording step
    _, _, g1 = _step(dec, cap, ctx)
    inp, tgt = cap[:, :-1].contiguous(), cap[:, 1:].contiguous()
    # second batch: another caption, so that g2 != g1 and "2 * g2" is distinguishable from g1 + g2
    cap2 = cap.clone()
    cap2[:, 1:4] = cap.flip(0)[:, 1:4]
    _, _, g2 = _step(dec, cap2, ctx)
    for p in dec.parameters():
        p.grad = None
    for c in (cap, cap2):                      # accumulate: no zeroing in between
        i, t = c[:, :-1].contiguous(), c[:, 1:].contiguous()
        with dec.weight_scope():
            X, _ = dec.forward_tbc({'roberta': i}, ctx)
            T, B, E = X.shape
            loss, _ = dec.adaptive_softmax.fused_loss(X.view(T * B, E), t.t().contiguous())
        loss.backward()
    checked = 0
    for n, p in dec.named_parameters():
        if n.endswith('weight_v') or n.endswith('weight_g'):
            want = g1[n] + g2[n]
            assert _close(p.grad, want, 1e-4), n
            assert not _close(p.grad, 2 * g2[n], 1e-3) or _close(g1[n], g2[n], 1e-3), n
            checked += 1
    assert checked >= 10
