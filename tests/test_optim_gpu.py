"""BertAdam (tt_bertadam_step) against the oracle restatement of
pytorch-pretrained-bert's BertAdam.step (oracle/restate.py: bert_adam_step).
Tolerance: fp32 elementwise math in a different association order (fused multiply-adds, the
gradient-norm reduction tree) -- 2e-6 absolute on parameters of O(1) after 14 steps; 1e-4 relative
on the moments, which inherit the relative error of the fp32 gradient-norm reduction over up to
1.6 M elements (torch CPU and the kernel add in different orders; the kernel's clip coefficient is
checked against a float64 norm to 2e-6 in the next test)."""
import math
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'oracle'))

SHAPES = [(1024, 1024), (4096,), (3, 5), (1,), (8192 + 3,), (50265 // 8, 257), (16, 1, 31)]
HYPER = dict(lr=1e-2, warmup=0.25, t_total=12, schedule='warmup_linear', b1=0.9, b2=0.98, e=1e-6,
             weight_decay=1e-5, max_grad_norm=0.1)       # config.yaml:126-136 with a short t_total


def _make(seed=0):
    g = torch.Generator().manual_seed(seed)
    return [torch.randn(*s, generator=g) for s in SHAPES]


def _grads(step, seed=100):
    g = torch.Generator().manual_seed(seed + step)
    # mixed magnitudes: some tensors clipped (norm >> 0.1), some not (norm << 0.1)
    return [torch.randn(*s, generator=g) * (1e-4 if i % 2 else 1.0) for i, s in enumerate(SHAPES)]


def test_bertadam_matches_oracle_over_schedule():
    import restate
    from tell_b200.optim import BertAdam
    ref = _make()
    state = [dict() for _ in ref]
    params = [torch.nn.Parameter(p.clone().cuda()) for p in _make()]
    opt = BertAdam(params, **HYPER)
    for step in range(14):          # crosses warm-up (3 steps), decay and the clamp at t_total
        gs = _grads(step)
        restate.bert_adam_step(ref, gs, state, **HYPER)
        for p, g in zip(params, gs):
            p.grad = g.cuda()
        opt.step()
        want_lr = HYPER['lr'] * restate.bert_adam_schedule(step, HYPER['t_total'], HYPER['warmup'])
        assert abs(opt.get_lr()[0] - want_lr) <= 1e-7 * max(1.0, want_lr), step
    st = opt._groups[0]
    assert int(st.step.item()) == 14
    off = 0
    for p, r, s in zip(params, ref, state):
        assert (p.detach().cpu() - r).abs().max().item() < 2e-6
        n = r.numel()
        m, v = st.m[off:off + n].cpu().view_as(r), st.v[off:off + n].cpu().view_as(r)
        assert (m - s['next_m']).abs().max().item() <= 1e-4 * s['next_m'].abs().max().item() + 1e-12
        assert (v - s['next_v']).abs().max().item() <= 1e-4 * s['next_v'].abs().max().item() + 1e-12
        off += n


def test_bertadam_clip_coefficients_and_nan_skip():
    from tell_b200.optim import BertAdam
    params = [torch.nn.Parameter(p.cuda()) for p in _make()]
    before = [p.detach().clone() for p in params]
    opt = BertAdam(params, **dict(HYPER, t_total=-1))
    gs = _grads(0)
    for p, g in zip(params, gs):
        p.grad = g.cuda()
    # NaN loss: nothing moves, the step counter stays
    opt.step(loss=torch.full((1,), float('nan'), device='cuda'))
    st = opt._groups[0]
    assert int(st.step.item()) == 0 and float(st.scratch[1].item()) == 1.0
    for p, b in zip(params, before):
        assert torch.equal(p.detach(), b)
    assert float(st.m.abs().max().item()) == 0.0
    # finite loss: the update happens; clip coefficient = min(1, 0.1 / (||g|| + 1e-6)) per tensor
    opt.step(loss=torch.ones(1, device='cuda'))
    assert int(st.step.item()) == 1 and float(st.scratch[1].item()) == 0.0
    coef = st.scratch[2:].cpu()
    for i, g in enumerate(gs):
        want = min(1.0, 0.1 / (g.double().norm().item() + 1e-6))
        assert abs(coef[i].item() - want) <= 2e-6 * want, i
    assert any(not torch.equal(p.detach(), b) for p, b in zip(params, before))


def test_bertadam_groups_tied_weights_and_graph_replay():
    """parameter_groups regexes (config.yaml:137-149), a tied weight stepping once, and a step
    captured in a CUDA graph advancing the device-side schedule on every replay."""
    import restate
    from tell_b200.optim import BertAdam
    torch.manual_seed(0)
    lin = torch.nn.Linear(64, 32).cuda()
    named = [('decoder.layers.0.w', lin.weight), ('decoder.embedder.b', lin.bias),
             ('decoder.adaptive_softmax.tied', lin.weight)]
    opt = BertAdam.from_config(named, lr=1e-3, parameter_groups=[[['^decoder.embedder'], {}],
                                                                  [['^decoder.layers.0'], {'lr': 1e-2}]],
                               warmup=0.5, t_total=4, b2=0.98, weight_decay=0.0, max_grad_norm=0.0)
    assert [len(g['params']) for g in opt.param_groups] == [1, 1, 0]
    assert opt.param_groups[1]['lr'] == 1e-2
    ref = [lin.bias.detach().cpu().clone(), lin.weight.detach().cpu().clone()]
    state = [dict(), dict()]
    gb, gw = torch.randn(32), torch.randn(32, 64)
    lin.bias.grad, lin.weight.grad = gb.cuda(), gw.cuda()
    opt.step()                                    # eager step uploads the segment tables
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        opt.step()
    for _ in range(3):
        graph.replay()
    torch.cuda.synchronize()
    kw = dict(warmup=0.5, t_total=4, b1=0.9, b2=0.98, e=1e-6, weight_decay=0.0, max_grad_norm=0.0)
    for _ in range(4):
        restate.bert_adam_step(ref[:1], [gb], state[:1], lr=1e-3, **kw)
        restate.bert_adam_step(ref[1:], [gw], state[1:], lr=1e-2, **kw)
    assert (lin.bias.detach().cpu() - ref[0]).abs().max().item() < 1e-6
    assert (lin.weight.detach().cpu() - ref[1]).abs().max().item() < 1e-6
    assert not math.isnan(lin.weight.sum().item())


def test_bertadam_rejects_cpu_and_missing_grads():
    from tell_b200 import _lib
    from tell_b200.optim import BertAdam
    with pytest.raises(_lib.TtError):
        BertAdam([torch.nn.Parameter(torch.zeros(4))], lr=1e-3)
    p = torch.nn.Parameter(torch.zeros(4, device='cuda'))
    with pytest.raises(_lib.TtError):
        BertAdam([p], lr=1e-3).step()
    with pytest.raises(ValueError):
        BertAdam([p], lr=-1.0)
