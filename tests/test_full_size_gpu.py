"""Parity at the REAL model size against vectors produced by the reference's own code
(oracle/gen_golden_full.py): the cfg-2 decoder (E 1024, d 64, K 3/7/15/31, vocab 50265) on
B=4, T=50, S=512, F=4, O=16 incl. the reference's 100-step `_generate`; ResNet-152 at 224x224 from
tell/models/resnet.py in both BatchNorm modes; RoBERTa-large from HF transformers.

Every test prints and records its MEASURED error (gpurun_out/parity_measured.jsonl; the copy of a
full run is committed as profiles/r2_parity_measured.jsonl); the asserted tolerances are at most 2x
what was measured there (bf16 modes) or the north-star gate (bf16x3: 1e-3 / token-exact)."""
import json
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'oracle'))
GOLD = os.path.join(ROOT, 'tests', 'golden')
FULL_SHAPES = dict(B=4, T=50, S=512, F=4, O=16, P=49)
FULL_SEED, FULL_GAIN, FULL_INPUT_SEED = 1, 2.0, 4247


def T_(x):
    return torch.from_numpy(np.asarray(x))


def record(name, **vals):
    vals = {k: (float(v) if isinstance(v, (float, np.floating)) else v) for k, v in vals.items()}
    line = json.dumps(dict(test=name, **vals))
    print('\nMEASURED ' + line)
    try:
        os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
        with open(os.path.join(ROOT, 'gpurun_out', 'parity_measured.jsonl'), 'a') as f:
            f.write(line + '\n')
    except OSError:
        pass


def _full_decoder(precision):
    from tell_b200 import config, synth
    from tell_b200.models import DynamicConvFacesObjectsDecoder
    from tell_b200.testing import build_decoder
    config.set_precision(precision)
    cfg = synth.CFG_FULL
    sd = synth.decoder_state_dict(cfg, seed=FULL_SEED, logit_gain=FULL_GAIN)
    dec = build_decoder(cfg, DynamicConvFacesObjectsDecoder, sd).cuda().eval()
    cap, ctx = synth.decoder_inputs(cfg, **FULL_SHAPES, seed=FULL_INPUT_SEED)
    g = np.load(os.path.join(GOLD, 'decoder_full.npz'))
    return cfg, dec, cap, ctx, g


def _rel(got, ref):
    return (got - ref).abs().max().item() / max(1e-6, ref.abs().max().item())


# tolerances: (decoder output abs, loss abs, gradient rel-to-max, attention weights abs, log-probs abs)
# bf16x3: the north-star gate (1e-3) on outputs / loss / attention / log-probs; gradients 5e-3 of
# the largest element (2.6e-3 measured on the deepest parameter of the backward chain).
# bf16: at most 2x the measured error (out 0.030, loss 0.0027, gradients 0.066, attention 4.6e-4,
# log-probs 0.088 on the first full run).
TOL = {'bf16x3': dict(out=1e-3, loss=1e-3, grad=5e-3, attn=1e-3, lp=1e-3),
       'bf16': dict(out=0.06, loss=0.006, grad=0.13, attn=1e-3, lp=0.18)}


@pytest.mark.parametrize('precision', ['bf16x3', 'bf16'])
def test_decoder_full_size_vs_reference(precision):
    """decoder_faces_objects.py:95-142 + adaptive_loss.py:27-73 + autograd at cfg-2 width and
    sequence shapes, against the reference's own output (not the restatement)."""
    cfg, dec, cap, ctx, g = _full_decoder(precision)
    tol = TOL[precision]
    inp, tgt = cap[:, :-1].contiguous().cuda(), cap[:, 1:].contiguous().cuda()
    cctx = {k: v.cuda() for k, v in ctx.items()}
    cctx['article'].requires_grad_(True)
    out, extra = dec({'roberta': inp}, cctx)
    ref_out = T_(g['dec_out'])
    e_out = (out.detach().cpu() - ref_out).abs()
    loss, ntok = dec.adaptive_softmax.fused_loss(out, tgt)
    assert int(ntok) == int(g['ntokens'][0])
    e_loss = abs(loss.item() - float(g['loss'][0]))
    loss.backward()
    d_art = cctx['article'].grad.cpu()
    e_dart = _rel(d_art[::16], T_(g['d_article_sub']))
    e_dart_norm = abs(d_art.double().norm().item() - g['d_article_stats'][0]) / g['d_article_stats'][0]
    params = dict(dec.named_parameters())
    e_gfull, e_gnorm = {}, {}
    for k in g.files:
        if k.startswith('gfull/'):
            name = k[len('gfull/'):]
            e_gfull[name] = _rel(params[name].grad.cpu(), T_(g[k]))
        if k.startswith('gsum/'):
            name = k[len('gsum/'):]
            if name not in params:       # tied duplicates are one parameter here
                continue
            got = params[name].grad
            norm = got.double().norm().item() if got is not None else 0.0
            e_gnorm[name] = abs(norm - g[k][0]) / max(1e-6, g[k][0])
    e_attn = max((extra['attn'][0][nm].cpu() - T_(g['attn0/' + nm])).abs().max().item()
                 for nm, _ in cfg['contexts'])
    with torch.no_grad():
        lp = dec.get_normalized_probs((out[:, -1:].detach(), None), True).cpu()
    e_lp = (lp - T_(g['log_probs_last'])).abs().max().item()
    worst_g = max(e_gfull, key=e_gfull.get)
    worst_n = max(e_gnorm, key=e_gnorm.get)
    record('decoder_full_size', precision=precision, out_max_abs=e_out.max().item(),
           out_mean_abs=e_out.mean().item(), ref_out_absmax=ref_out.abs().max().item(),
           loss_abs=e_loss, loss_ref=float(g['loss'][0]), d_article_rel=e_dart,
           d_article_norm_rel=e_dart_norm, grad_full_rel_worst=e_gfull[worst_g],
           grad_full_worst=worst_g, grad_norm_rel_worst=e_gnorm[worst_n], grad_norm_worst=worst_n,
           attn_abs=e_attn, log_probs_abs=e_lp)
    assert e_out.max().item() < tol['out']
    assert e_loss < tol['loss']
    assert e_dart < tol['grad'] and e_dart_norm < tol['grad']
    assert e_gfull[worst_g] < tol['grad'], worst_g
    assert e_gnorm[worst_n] < tol['grad'] * 2.5, worst_n
    assert e_attn < tol['attn']
    assert e_lp < tol['lp']
    del dec
    torch.cuda.empty_cache()


class _Stub(torch.nn.Module):
    n_layers = 24


@pytest.mark.parametrize('precision', ['bf16x3', 'bf16'])
def test_greedy_full_size_vs_reference_generate(precision):
    """The reference's own `_generate` (transformer_faces_objects.py:399-494) at full width, B=4,
    S=512: bf16x3 must emit exactly its 101 columns with log-probs within 1e-3; the throughput mode
    reports its token-agreement rate (rows may legitimately diverge after a near-tie: the smallest
    top-1/top-2 margin on the reference path is stored in the golden)."""
    from tell_b200.models import TransformerFacesObjectModel
    from tell_b200.modules import AdaptiveLoss
    cfg, dec, cap, ctx, g = _full_decoder(precision)
    model = TransformerFacesObjectModel(None, dec, AdaptiveLoss(1), weigh_bert=True, resnet=_Stub(),
                                        roberta=_Stub(), padding_value=1, vocab_size=cfg['vocab'])
    model = model.cuda().eval()
    cctx = {k: v.cuda() for k, v in ctx.items()}
    ref_ids, ref_lp = T_(g['greedy_ids']), T_(g['greedy_lp'])
    lp, ids, _ = model._generate(cap[:, 0:1].cuda(), cctx)
    ids, lp = ids.cpu(), lp.cpu()
    n = min(ids.shape[1], ref_ids.shape[1])
    same = (ids[:, :n] == ref_ids[:, :n])
    # first divergence per row
    first_div = [int((~same[b]).nonzero()[0]) if (~same[b]).any() else n for b in range(ids.shape[0])]
    prefix_rate = float(np.mean([f / n for f in first_div]))
    agree = same.float().mean().item()
    lp_err = (lp[:, :n - 1] - ref_lp[:, :n - 1]).abs()
    lp_err_on_path = max((lp_err[b, :max(0, first_div[b] - 1)].max().item() if first_div[b] > 1 else 0.0)
                         for b in range(ids.shape[0]))
    record('greedy_full_size', precision=precision, columns=int(ref_ids.shape[1]),
           token_agreement=agree, agreed_prefix_fraction=prefix_rate, first_divergence=first_div,
           logprob_abs_on_agreed_path=lp_err_on_path, ref_margin_min=float(g['greedy_margin_min'][0]))
    if precision == 'bf16x3':
        assert ids.shape == ref_ids.shape
        assert torch.equal(ids, ref_ids)
        assert lp_err.max().item() < 1e-3
    else:
        assert prefix_rate >= GREEDY_BF16_MIN_PREFIX
        assert lp_err_on_path < 0.07          # 0.033 measured
    del model, dec
    torch.cuda.empty_cache()


# measured: rows 0-1 reproduce all 101 columns, rows 2-3 leave the reference path at columns 1 / 21
# (agreed-prefix fraction 0.55, token agreement 0.63) -- the reference path has top-1/top-2 margins
# down to 3e-3, far below bf16 log-prob error (0.03); bf16x3 above is the token-exact gate.
GREEDY_BF16_MIN_PREFIX = 0.4


RESNET_SEED, RESNET_IMG_SEED, RESNET_BN3_GAIN = 3, 11, 0.25
# End to end through 152 layers, relative to the largest output.
#  * vs the reference's fp32 output: bf16 storage of weights and activations is the whole error.
#    Running statistics: 1.4e-2 measured.  Batch statistics on RANDOM weights are chaotic (a BatchNorm
#    network at initialisation amplifies perturbations layer by layer: fp32 vs fp64 already differ by
#    3e-4 here, and rounding ONLY the weights to bf16 moves the output by 12 %): 0.26 measured, cosine
#    0.97 -- a property of the synthetic weights, not of the kernels, which is what the next two
#    comparisons establish.
#  * vs the SAME algorithm with bf16 storage modelled on the CPU (restate.resnet152_forward(storage=
#    bf16 round), pinned in fp32 mode to the reference golden): 1.6e-2 in running mode; in batch mode
#    the same chaos amplifies the fp32 summation-order differences between the CPU and the tensor
#    cores through flipped bf16 roundings (0.19 measured), so this comparison is loose there too.
#  * per Bottleneck, teacher-forced with the oracle's own block input: no accumulation, every block
#    of the network checked against the fp32 algorithm at bf16-storage tolerance -- THE tight check
#    in both modes (worst block measured: 6.3e-3 running, 1.0e-2 batch).
RESNET_TOL = {'running': dict(ref=0.03, cos=0.9995, model=0.03, block=0.015),
              'batch': dict(ref=0.45, cos=0.95, model=0.40, block=0.03)}


def _resnet(bn_mode):
    from tell_b200 import synth
    from tell_b200.models import ResNetFeatureExtractor
    sd = synth.resnet_state_dict((3, 8, 36, 3), seed=RESNET_SEED, bn3_gain=RESNET_BN3_GAIN)
    net = ResNetFeatureExtractor((3, 8, 36, 3))
    net.load_state_dict(sd, strict=True)
    net = net.cuda()
    net.train(bn_mode == 'batch')           # bn_mode 'auto': follows train()/eval() like nn.BatchNorm2d
    rs = np.random.RandomState(RESNET_IMG_SEED)
    img = torch.from_numpy(rs.standard_normal((2, 3, 224, 224)).astype(np.float32))
    return sd, net, img


def _bf16(t):
    return t.bfloat16().float()


@pytest.mark.parametrize('bn_mode', ['running', 'batch'])
def test_resnet152_full_depth_vs_reference(bn_mode):
    """tell/models/resnet.py:92-117 at [3,8,36,3] / 224x224 in eval() and train() BatchNorm modes."""
    import restate
    g = np.load(os.path.join(GOLD, 'resnet152.npz'))
    sd, net, img = _resnet(bn_mode)
    tol = RESNET_TOL[bn_mode]
    ref = T_(g['y_eval' if bn_mode == 'running' else 'y_train'])
    out = net(img.cuda()).cpu()
    assert out.shape == ref.shape == (2, 2048, 7, 7)
    with torch.no_grad():
        model = restate.resnet152_forward(img, sd, prefix='', bn_mode=bn_mode, storage=_bf16)
    err = (out - ref).abs()
    cos = torch.nn.functional.cosine_similarity(out.flatten(), ref.flatten(), dim=0).item()
    rel = err.max().item() / ref.abs().max().item()
    rel_rms = (err.pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()).item()
    rel_model = _rel(out, model)
    rms_model = ((out - model).pow(2).mean().sqrt() / model.pow(2).mean().sqrt()).item()
    model_vs_ref = _rel(model, ref)
    stats = {}
    if bn_mode == 'batch':
        after = net.state_dict()
        for k in g.files:
            if k.startswith('after_train/'):
                name = k[len('after_train/'):]
                stats[name] = _rel(after[name].cpu(), T_(g[k]))
        assert int(after['bn1.num_batches_tracked']) == 1
        assert int(after['layer4.2.bn3.num_batches_tracked']) == 1
    record('resnet152_full_depth', bn_mode=bn_mode, vs_reference_rel_max=rel, vs_reference_rel_rms=rel_rms,
           vs_reference_cosine=cos, ref_absmax=ref.abs().max().item(),
           vs_bf16_storage_model_rel_max=rel_model, vs_bf16_storage_model_rel_rms=rms_model,
           bf16_storage_model_vs_reference_rel_max=model_vs_ref,
           running_stats_rel_worst=max(stats.values()) if stats else None)
    assert rel < tol['ref'], rel
    assert cos > tol['cos'], cos
    assert rel_model < tol['model'], rel_model
    for name, e in stats.items():
        assert e < 3e-2, (name, e)


@pytest.mark.parametrize('bn_mode', ['running', 'batch'])
def test_resnet152_every_block_teacher_forced(bn_mode):
    """Each of the 50 Bottlenecks (and the stem) on the ORACLE's input to that block, against the
    oracle's fp32 output of that block: per-block error without accumulation through the depth."""
    import restate
    sd, net, img = _resnet(bn_mode)
    tol = RESNET_TOL[bn_mode]['block']
    collect = []
    with torch.no_grad():
        restate.resnet152_forward(img, sd, prefix='', bn_mode=bn_mode, collect=collect)
    worst = (0.0, None)
    for name, x, y in collect:
        if name == 'stem':
            got = net.stem_nhwc(x.cuda())
        else:
            li, bi = int(name[5]), int(name.split('.')[1])
            got = net.block_nhwc(x.permute(0, 2, 3, 1).contiguous().bfloat16().cuda(), li, bi)
        got = got.float().permute(0, 3, 1, 2).cpu()
        assert got.shape == y.shape, name
        e = _rel(got, y)
        if e > worst[0]:
            worst = (e, name)
        assert e < tol, (name, e)
    record('resnet152_per_block', bn_mode=bn_mode, blocks=len(collect), rel_max_worst=worst[0],
           worst_block=worst[1])


ROBERTA_SEED, ROBERTA_IDS_SEED = 5, 21


def test_roberta_large_vs_hf_reference():
    """The `roberta.extract_features(ids, return_all_hiddens=True)` call of
    transformer_faces_objects.py:352-353 at roberta.large size (24 layers) against HF transformers'
    RobertaModel on the same seeded weights, and the 25-layer mix of :355-364."""
    from tell_b200 import functional as Fn
    from tell_b200 import synth
    from tell_b200.models import RobertaEncoder
    g = np.load(os.path.join(GOLD, 'roberta_large.npz'))
    L, E, H, FFN, V, P = 24, 1024, 16, 4096, 50265, 514
    sd = synth.roberta_state_dict(L, E, FFN, V, P, seed=ROBERTA_SEED)
    enc = RobertaEncoder(L, E, H, FFN, V, P)
    enc.load_state_dict(sd, strict=True)
    del sd
    enc = enc.cuda().eval()
    rs = np.random.RandomState(ROBERTA_IDS_SEED)
    ids = synth.article_batch(3, 200, 50265, rs, min_len=40)
    ids[1, 7:] = 1
    ids[1, 6] = 2
    real = ids != 1
    B, S = ids.shape
    pos = g['sample_pos']
    ref_at = T_(g['hidden_at_pos'])                              # [25,B,8,E]
    res = {}
    for varlen in (True, False):
        enc.varlen = varlen
        hid, _ = enc.all_hiddens(ids.cuda())
        Hh = hid.float().view(L + 1, B, S, E).cpu()
        got_at = torch.stack([Hh[:, b, pos[b]] for b in range(B)], 1)
        e_layer = (got_at - ref_at).abs().amax(dim=(1, 2, 3))    # per layer
        m = real.unsqueeze(0).unsqueeze(-1).float()
        norms = ((Hh * m) ** 2).sum(dim=(2, 3)).sqrt()
        e_norm = ((norms - T_(g['layer_norms'])).abs() / T_(g['layer_norms'])).max().item()
        bw = T_(g['bert_weight']).cuda()
        mix = Fn.LayerMixFn.apply(hid, bw).view(B, S, E).cpu() * real.unsqueeze(-1)
        ref_mix = T_(g['mix']).float()
        e_mix = (mix - ref_mix).abs().max().item()
        res[varlen] = (e_layer, e_norm, e_mix)
        record('roberta_large', varlen=varlen, hidden_abs_max_last=e_layer[-1].item(),
               hidden_abs_max_worst_layer=e_layer.max().item(), layer_norm_rel_worst=e_norm,
               mix_abs_max=e_mix, ref_absmax=float(g['layer_absmax'].max()),
               ref_mix_absmax=ref_mix.abs().max().item())
    for varlen, (e_layer, e_norm, e_mix) in res.items():
        assert e_layer.max().item() < ROBERTA_TOL['hidden'], (varlen, e_layer)
        assert e_norm < ROBERTA_TOL['norm']
        assert e_mix < ROBERTA_TOL['mix']
    del enc
    torch.cuda.empty_cache()


# bf16 activations / bf16 residual stream through 24 post-LN layers, hidden states of O(1..4)
# measured: hidden 0.071 (of |h| <= 6.3), per-layer norm 3e-4 relative, layer mix 0.020 (of 1.58)
ROBERTA_TOL = dict(hidden=0.15, norm=1e-3, mix=0.045)
