"""CPU suite: pins the oracle (oracle/restate.py) to golden vectors produced by the reference's
own modules (oracle/gen_golden.py) and to the reference's known-answer test."""
import math
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'oracle'))
GOLD = os.path.join(ROOT, 'tests', 'golden')

import restate  # noqa: E402
from tell_b200 import synth  # noqa: E402

SHAPES = dict(B=3, T=9, S=11, F=3, O=4, P=5)


def T_(x):
    return torch.from_numpy(np.asarray(x))


def test_make_positions_reference_vectors():
    """tell/modules/token_embedders/tests/test_positional.py:12-32."""
    left_in = torch.tensor([[9, 9, 9, 9, 9], [1, 9, 9, 9, 9], [1, 1, 1, 9, 9]])
    left_out = torch.tensor([[2, 3, 4, 5, 6], [1, 2, 3, 4, 5], [1, 1, 1, 2, 3]])
    right_in = torch.tensor([[9, 9, 9, 9, 9], [9, 9, 9, 9, 1], [9, 9, 1, 1, 1]])
    right_out = torch.tensor([[2, 3, 4, 5, 6], [2, 3, 4, 5, 1], [2, 3, 1, 1, 1]])
    assert torch.equal(restate.make_positions(left_in, 1, True), left_out)
    assert torch.equal(restate.make_positions(right_in, 1, False), right_out)


def test_op_goldens():
    g = np.load(os.path.join(GOLD, 'ops.npz'))
    for (T, K) in [(5, 15), (12, 7)]:
        tag = 'dynconv_T%d_K%d/' % (T, K)
        y = restate.dynamic_conv(T_(g[tag + 'x']), T_(g[tag + 'w']), K, 4)
        assert (y - T_(g[tag + 'y'])).abs().max() < 1e-5
    for (T, K, H) in [(5, 7, 4), (12, 3, 8), (4, 15, 2)]:       # lightweight.py:88-240, incl. K > T
        tag = 'lightconv_T%d_K%d/' % (T, K)
        y = restate.lightweight_conv(T_(g[tag + 'x']), T_(g[tag + 'w']), K, H)
        assert (y - T_(g[tag + 'y'])).abs().max() < 1e-5
    sd = {k[len('mha_empty/'):]: T_(g[k]) for k in g.files if k.startswith('mha_empty/a.')}
    q = T_(g['mha_empty/q'])
    y, w = restate.multi_head_attention(q, torch.zeros(1, 2, 0), torch.zeros(2, 1, dtype=torch.bool),
                                        sd, 'a.', 4, True)
    assert (y - T_(g['mha_empty/y'])).abs().max() < 1e-5
    assert (w - T_(g['mha_empty/w'])).abs().max() < 1e-6
    assert w.shape == (2, 4, 2)      # bias row + zero row only


def test_forward_glue_goldens():
    g = np.load(os.path.join(GOLD, 'forward_glue.npz'))
    hid = [T_(h) for h in g['hid']]
    ctx = restate.build_contexts(T_(g['feats']), hid, T_(g['bert_weight']), T_(g['art']),
                                 T_(g['faces']).clone(), T_(g['objs']).clone())
    for k, v in ctx.items():
        assert torch.allclose(v.float(), T_(g['ctx/' + k]).float(), atol=1e-6), k
    inp, tgt = restate.shift_caption(T_(g['cap']))
    assert torch.equal(inp, T_(g['caption_ids'])) and torch.equal(tgt, T_(g['target_ids']))


@pytest.mark.parametrize('tag,cfg', [('tiny_faces_objects', synth.CFG_TINY),
                                     ('tiny_no_image', synth.CFG_TINY_NO_IMAGE),
                                     ('tiny_flattened', synth.CFG_TINY_FLATTENED),
                                     ('tiny_faces_parallel', synth.CFG_TINY_FACES)])
def test_decoder_goldens(tag, cfg):
    g = np.load(os.path.join(GOLD, 'decoder_%s.npz' % tag))
    sd = synth.decoder_state_dict(cfg, seed=0, logit_gain=4.0)
    for k in sd:
        if sd[k].is_floating_point():
            sd[k] = sd[k].clone().requires_grad_(True)
    # re-tie after cloning
    pre = 'embedder.token_embedder_adaptive.embeddings.'
    sd['adaptive_softmax.head.word_proj.weight'] = sd[pre + '0.0.weight']
    for i in range(2):
        sd['adaptive_softmax.tail.%d.2.weight' % i] = sd[pre + '%d.0.weight' % (i + 1)]
    cap, ctx = synth.decoder_inputs(cfg, **SHAPES, seed=1234)
    inp, tgt = cap[:, :-1].contiguous(), cap[:, 1:].contiguous()
    ocfg = synth.oracle_cfg(cfg)
    ctx['article'].requires_grad_(True)
    out, attns = restate.decoder_forward(inp, ctx, sd, ocfg)
    assert (out.detach() - T_(g['dec_out'])).abs().max() < 2e-5
    loss_sum, n, loss = restate.adaptive_loss(out, tgt, sd, ocfg['cutoffs'])
    assert n == int(g['ntokens'][0])
    assert abs(loss_sum.item() - float(g['loss_sum'][0])) < 1e-3
    loss.backward()
    assert (ctx['article'].grad - T_(g['d_article'])).abs().max() < 1e-5
    for k in g.files:
        if k.startswith('gfull/'):
            name = k[len('gfull/'):]
            assert (sd[name].grad - T_(g[k])).abs().max() < 2e-5, name
        if k.startswith('gsum/'):
            name = k[len('gsum/'):]
            gr = sd[name].grad
            norm = gr.double().norm().item() if gr is not None else 0.0
            assert abs(norm - g[k][0]) < 1e-4 * max(1.0, g[k][0]), name
    if 'attn0/article' in g.files:
        for nm in ocfg['ctx_names']:
            assert (attns[0][nm].detach() - T_(g['attn0/' + nm])).abs().max() < 1e-5
    with torch.no_grad():
        sdd = {k: v.detach() for k, v in sd.items()}
        ctxd = {k: v.detach() for k, v in ctx.items()}
        lp = restate.adaptive_log_prob(out[:, -1:].detach(), sdd, ocfg['cutoffs'])
        assert (lp - T_(g['log_probs_last'])).abs().max() < 1e-4
        state, steps = {}, []
        for t in range(inp.shape[1]):
            o, _ = restate.decoder_forward(inp[:, t:t + 1], ctxd, sdd, ocfg, state)
            steps.append(o)
        assert (torch.cat(steps, 1) - T_(g['dec_out_incremental'])).abs().max() < 2e-5
        if 'greedy_ids' in g.files:
            ids, lps = restate.greedy_generate(cap[:, 0:1], ctxd, sdd, ocfg, gen_len=20)
            assert torch.equal(ids, T_(g['greedy_ids'])[:, :21])
            assert (lps - T_(g['greedy_lp'])[:, :20]).abs().max() < 1e-4


def test_tail_local_index_one_quirk():
    """ignore_index=1 is applied to cluster-local targets (adaptive_loss.py:57-60)."""
    torch.manual_seed(0)
    E = 8
    sd = {'adaptive_softmax.head.word_proj.weight': torch.randn(10, E),
          'adaptive_softmax.head.class_proj.weight': torch.randn(2, E),
          'adaptive_softmax.tail.0.0.weight': torch.randn(E, E),
          'adaptive_softmax.tail.0.2.weight': torch.randn(10, E),
          'adaptive_softmax.tail.1.0.weight': torch.randn(E, E),
          'adaptive_softmax.tail.1.2.weight': torch.randn(10, E)}
    X = torch.randn(1, 4, E)
    full = -restate.adaptive_log_prob(X, sd, [10, 20, 30])[0]
    t_ok = torch.tensor([[5, 12, 22, 1]])
    t_quirk = torch.tensor([[5, 11, 21, 1]])
    l_ok, n, _ = restate.adaptive_loss(X, t_ok, sd, [10, 20, 30])
    true_ok = sum(full[i, t_ok[0, i]] for i in range(3))
    assert n == 3 and abs(l_ok.item() - true_ok.item()) < 1e-4
    l_q, _, _ = restate.adaptive_loss(X, t_quirk, sd, [10, 20, 30])
    true_q = sum(full[i, t_quirk[0, i]] for i in range(3))
    assert l_q.item() < true_q.item() - 1e-3      # tail terms of local index 1 are dropped


def test_bert_adam_restatement_known_answers():
    """Hand-computed known answers for the BertAdam restatement (the optimizer behind
    `type: bert_adam`, config.yaml:126-136; pytorch-pretrained-bert is absent here, so these pin
    the arithmetic of the published algorithm rather than the package itself)."""
    import restate
    # schedule: warm-up is linear in progress, then linear decay to 0 at t_total, clamped after
    assert restate.bert_adam_schedule(0, 1000, 0.05) == 0.0
    assert abs(restate.bert_adam_schedule(25, 1000, 0.05) - 0.5) < 1e-12
    assert abs(restate.bert_adam_schedule(500, 1000, 0.05) - 0.5 / 0.95) < 1e-12
    assert restate.bert_adam_schedule(1500, 1000, 0.05) == 0.0
    assert restate.bert_adam_schedule(7, -1, 0.05) == 1.0
    assert restate.bert_adam_schedule(500, 1000, 0.05, 'warmup_constant') == 1.0
    # one step on a scalar: clip 0.5 -> 0.1/(0.5+1e-6), m = 0.1 g, v = 0.001 g^2
    p, g, st = [torch.tensor([1.0])], [torch.tensor([0.5])], [dict()]
    restate.bert_adam_step(p, g, st, lr=1e-3, b1=0.9, b2=0.999, e=1e-6, weight_decay=0.01,
                           max_grad_norm=0.1)
    gc = 0.5 * (0.1 / (0.5 + 1e-6))
    m, v = 0.1 * gc, 0.001 * gc * gc
    want = 1.0 - 1e-3 * (m / (v ** 0.5 + 1e-6) + 0.01 * 1.0)
    assert abs(p[0].item() - want) < 1e-7 and st[0]['step'] == 1
    assert abs(st[0]['next_m'].item() - m) < 1e-9 and abs(st[0]['next_v'].item() - v) < 1e-11
    # unclipped small gradient, no decay: update is m / (sqrt(v) + e)
    p, g, st = [torch.tensor([2.0])], [torch.tensor([1e-3])], [dict()]
    restate.bert_adam_step(p, g, st, lr=1.0, b1=0.9, b2=0.98, e=1e-6, weight_decay=0.0,
                           max_grad_norm=0.1)
    want = 2.0 - (1e-4 / ((0.02 * 1e-6) ** 0.5 + 1e-6))
    assert abs(p[0].item() - want) < 1e-5


def test_face_encoder_restatements_match_reference_goldens():
    """oracle/restate.py face encoders vs tests/golden/facenet.npz (outputs of the reference's own
    InceptionResnetV1 / PNet / RNet / ONet modules, the latter also with the vendored checkpoints)."""
    import restate
    from tell_b200 import synth
    g = np.load(os.path.join(ROOT, 'tests', 'golden', 'facenet.npz'))
    fwd = {'pnet': restate.pnet_forward, 'rnet': restate.rnet_forward, 'onet': restate.onet_forward}
    with torch.no_grad():
        for name in ('pnet', 'rnet', 'onet'):
            keys = [k[len(name) + 8:] for k in g.files if k.startswith(name + '_real_w.')]
            sd_real = {k: torch.from_numpy(g['%s_real_w.%s' % (name, k)]) for k in keys}
            sd_synth = synth.shaped_state_dict({k: v.shape for k, v in sd_real.items()}, seed=5)
            x = torch.from_numpy(g[name + '_x'])
            for kind, sd in (('synth', sd_synth), ('real', sd_real)):
                for i, o in enumerate(fwd[name](x, sd)):
                    want = torch.from_numpy(g['%s_%s_out%d' % (name, kind, i)])
                    assert (o - want).abs().max().item() < 1e-5, (name, kind, i)
        from tell_b200.facenet import InceptionResnetV1
        shapes = {k: v.shape for k, v in InceptionResnetV1(num_classes=10).state_dict().items()}
        sd = synth.shaped_state_dict(shapes, seed=3)
        emb, logits = restate.inception_resnet_v1_forward(torch.from_numpy(g['irv1_x']), sd)
        assert (emb - torch.from_numpy(g['irv1_emb'])).abs().max().item() < 1e-5
        assert (logits - torch.from_numpy(g['irv1_logits'])).abs().max().item() < 1e-4


# ------------------------------------------------------------------ full-size goldens (round 2)
def test_resnet152_restatement_vs_reference_golden_both_bn_modes():
    """restate.resnet152_forward at [3,8,36,3] / 224x224 against the output of the reference's own
    ResNetFeatureExtractor (tests/golden/resnet152.npz) in eval() and train() BatchNorm modes, and
    the running statistics the train-mode forward leaves behind."""
    g = np.load(os.path.join(GOLD, 'resnet152.npz'))
    sd = synth.resnet_state_dict((3, 8, 36, 3), seed=3, bn3_gain=0.25)
    img = T_(np.random.RandomState(11).standard_normal((2, 3, 224, 224)).astype(np.float32))
    with torch.no_grad():
        y = restate.resnet152_forward(img, sd, prefix='')
        ref = T_(g['y_eval'])
        assert (y - ref).abs().max().item() < 1e-4 * ref.abs().max().item()
        y, stats = restate.resnet152_forward(img, sd, prefix='', bn_mode='batch', return_stats=True)
        ref = T_(g['y_train'])
        # train-mode BN on random weights amplifies fp32 summation-order noise (fp32 vs fp64: 3e-4)
        assert (y - ref).abs().max().item() < 2e-3 * ref.abs().max().item()
    for k in g.files:
        if k.startswith('after_train/'):
            assert torch.allclose(stats[k[len('after_train/'):]], T_(g[k]), rtol=1e-3, atol=1e-5), k


def test_roberta_restatement_vs_hf_golden():
    """restate.roberta_forward at roberta.large size against HF RobertaModel's hidden states
    (tests/golden/roberta_large.npz): pins the stand-in oracle of SURVEY 8c."""
    g = np.load(os.path.join(GOLD, 'roberta_large.npz'))
    L, E, H = 24, 1024, 16
    sd = synth.roberta_state_dict(L, E, 4096, 50265, 514, seed=5)
    rs = np.random.RandomState(21)
    ids = synth.article_batch(3, 200, 50265, rs, min_len=40)
    ids[1, 7:] = 1
    ids[1, 6] = 2
    with torch.no_grad():
        hs = torch.stack(restate.roberta_forward(ids, sd, L, H, prefix=''))     # [25,B,S,E]
    pos = g['sample_pos']
    got = torch.stack([hs[:, b, pos[b]] for b in range(3)], 1)
    assert (got - T_(g['hidden_at_pos'])).abs().max().item() < 2e-4
    real = (ids != 1)
    norms = ((hs * real.unsqueeze(0).unsqueeze(-1)) ** 2).sum(dim=(2, 3)).sqrt()
    assert ((norms - T_(g['layer_norms'])).abs() / T_(g['layer_norms'])).max().item() < 1e-4
    w = torch.softmax(T_(g['bert_weight']), 0)
    mix = (hs * w.view(-1, 1, 1, 1)).sum(0) * real.unsqueeze(-1)
    assert (mix - T_(g['mix']).float()).abs().max().item() < 2e-3       # golden stored as fp16


def test_decoder_restatement_vs_reference_full_size_golden():
    """restate.decoder_forward / adaptive_loss / adaptive_log_prob at the cfg-2 architecture and
    shapes (B=4, T=50, S=512) against the reference decoder's own output."""
    g = np.load(os.path.join(GOLD, 'decoder_full.npz'))
    cfg = synth.CFG_FULL
    sd = synth.decoder_state_dict(cfg, seed=1, logit_gain=2.0)
    cap, ctx = synth.decoder_inputs(cfg, B=4, T=50, S=512, F=4, O=16, P=49, seed=4247)
    inp, tgt = cap[:, :-1].contiguous(), cap[:, 1:].contiguous()
    ocfg = synth.oracle_cfg(cfg)
    with torch.no_grad():
        out, extra = restate.decoder_forward(inp, ctx, sd, ocfg)
        assert (out - T_(g['dec_out'])).abs().max().item() < 5e-5
        loss_sum, n, loss = restate.adaptive_loss(out, tgt, sd, ocfg['cutoffs'])
        assert n == int(g['ntokens'][0])
        assert abs(loss.item() - float(g['loss'][0])) < 1e-4 * float(g['loss'][0])
        lp = restate.adaptive_log_prob(out[:, -1:], sd, ocfg['cutoffs'])
        assert (lp - T_(g['log_probs_last'])).abs().max().item() < 1e-4
    assert float(g['greedy_margin_min'][0]) > 2e-3      # the stored greedy path is decidable at 1e-3
    assert g['greedy_ids'].shape == (4, 101)


def test_bert_adam_restatement_vs_torch_adam_trajectory():
    """Independent cross-check of restate.bert_adam_step against torch's own optimizer arithmetic
    over a 40-step trajectory (pytorch-pretrained-bert is absent, so its BertAdam cannot be run).
    BertAdam is Adam WITHOUT bias correction: p -= lr_t * (m / (sqrt(v) + e) + wd * p).  torch's Adam
    computes p -= (lr / bc1) * m / (sqrt(v) / sqrt(bc2) + eps) with bc1 = 1 - b1^t, bc2 = 1 - b2^t, so
    feeding it lr = lr_t * bc1 / sqrt(bc2) and eps = e / sqrt(bc2) at every step reproduces the
    uncorrected update exactly; the decoupled decay term and the per-tensor clip
    (torch.nn.utils.clip_grad_norm_, what BertAdam itself calls) are applied around it."""
    torch.manual_seed(0)
    shapes = [(7, 5), (11,), (3, 4, 2)]
    hyper = dict(lr=1e-3, warmup=0.1, t_total=40, b1=0.9, b2=0.999, e=1e-6, weight_decay=1e-2,
                 max_grad_norm=0.1)
    p_ref = [torch.randn(s) for s in shapes]
    p_tch = [torch.nn.Parameter(p.clone()) for p in p_ref]
    state = [dict() for _ in shapes]
    opts = [torch.optim.Adam([p], lr=1.0, betas=(hyper['b1'], hyper['b2']), eps=1.0) for p in p_tch]
    for step in range(40):
        grads = [torch.randn(s) * (3.0 if step % 3 == 0 else 0.05) for s in shapes]   # clipped and unclipped steps
        restate.bert_adam_step(p_ref, grads, state, **hyper)
        lr_t = hyper['lr'] * restate.bert_adam_schedule(step, hyper['t_total'], hyper['warmup'])
        t = step + 1
        bc1, bc2 = 1 - hyper['b1'] ** t, 1 - hyper['b2'] ** t
        for p, g, opt in zip(p_tch, grads, opts):
            p.grad = g.clone()
            torch.nn.utils.clip_grad_norm_([p], hyper['max_grad_norm'])
            decay = lr_t * hyper['weight_decay'] * p.detach().clone()
            for grp in opt.param_groups:
                grp['lr'] = lr_t * bc1 / math.sqrt(bc2)
                grp['eps'] = hyper['e'] / math.sqrt(bc2)
            opt.step()
            with torch.no_grad():
                p.sub_(decay)
        for a, b in zip(p_ref, p_tch):
            assert (a - b.detach()).abs().max().item() < 2e-6, step
    # the trajectory is not trivial: parameters moved
    assert all((a - torch.zeros_like(a)).abs().max() > 0 for a in p_ref)
