"""TEST INFRASTRUCTURE -- golden vectors at the REAL model size (BASELINE configs[1] shapes), made by
the reference's own code running in the build container:

  decoder_full.npz   tell/models/decoder_faces_objects.py (unmodified, via oracle/ref_loader.py) at
                     E 1024 / 16 heads (d 64) / 4 layers K 3,7,15,31 / vocab 50265 / cutoffs 5000,20000
                     on B=4, T=50, S=512, F=4, O=16, P=49: output, loss, gradients (full for a few
                     parameters, norm+sum for all), layer-0 attention weights, last-position
                     log-probs, and the reference's own `_generate` (transformer_faces_objects.py:
                     399-494) greedy ids / log-probs for 100 steps with top-1/top-2 margins.
  resnet152.npz      tell/models/resnet.py:ResNetFeatureExtractor(Bottleneck,[3,8,36,3]) (unmodified)
                     at 224x224, B=2, in eval() (running statistics) AND train() (batch statistics --
                     what the reference's training step runs, callback_apex_trainer.py:259), plus the
                     running statistics the train-mode forward leaves behind.
  roberta_large.npz  HF transformers RobertaModel (24 layers, E 1024, 16 heads, FFN 4096, vocab 50265,
                     514 positions) standing in for fairseq roberta.large (un-vendored), with
                     the fairseq-keyed seeded weights of synth.roberta_state_dict mapped onto the HF
                     key layout: all 25 hidden states at sampled token positions, per-layer norms over
                     the real tokens, and the 25-layer mix of transformer_faces_objects.py:355-364.

Weights and inputs come from tell_b200/synth.py (numpy RandomState, machine independent), so the
vectors are small.  Each section first asserts that oracle/restate.py agrees with the reference
output it stores.  Run:  python oracle/gen_golden_full.py [decoder] [resnet] [roberta]
"""
import os
import sys
import time
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'oracle'))
sys.path.insert(0, os.path.join(ROOT, 'transform-and-tell_b200'))

import ref_loader  # noqa: E402
import restate  # noqa: E402
from tell_b200 import synth  # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden')
FULL_SHAPES = dict(B=4, T=50, S=512, F=4, O=16, P=49)
FULL_SEED, FULL_GAIN, FULL_INPUT_SEED = 1, 2.0, 4247


def decoder_full():
    from gen_golden import grads_summary, ref_decoder
    cfg = synth.CFG_FULL
    t0 = time.time()
    sd = synth.decoder_state_dict(cfg, seed=FULL_SEED, logit_gain=FULL_GAIN)
    cap, ctx = synth.decoder_inputs(cfg, **FULL_SHAPES, seed=FULL_INPUT_SEED)
    inp, tgt = cap[:, :-1].contiguous(), cap[:, 1:].contiguous()
    dec = ref_decoder(cfg, 'faces_objects', sd).eval()
    crit = ref_loader.adaptive_loss()
    ctx['article'].requires_grad_(True)
    out, extra = dec({'roberta': inp}, ctx)
    loss, n = crit(dec.adaptive_softmax, (out, None), tgt)
    final = loss / np.log(2) / n
    final.backward()
    g = dict(dec_out=out.detach().numpy(), loss_sum=loss.detach().numpy(), ntokens=np.array([n]),
             loss=final.detach().numpy())
    da = ctx['article'].grad
    g['d_article_sub'] = da[::16].numpy()                       # every 16th article position
    g['d_article_stats'] = np.array([da.double().norm().item(), da.double().sum().item()])
    for k, v in grads_summary(dec.named_parameters()).items():
        g['gsum/' + k] = v
    full = ['layers.0.conv.weight_linear.weight', 'layers.3.conv.weight_linear.weight',
            'layers.0.linear1.weight_g', 'layers.2.fc1.weight_g',
            'layers.1.context_attns.article.bias_k', 'layers.3.context_attns.faces.bias_v',
            'layers.0.conv_layer_norm.weight', 'layers.3.final_layer_norm.bias',
            'adaptive_softmax.head.class_proj.weight', 'layers.1.fc2.bias',
            'layers.2.context_attns.obj.out_proj.bias']
    params = dict(dec.named_parameters())
    for k in full:
        g['gfull/' + k] = params[k].grad.numpy()
    for nm, _ in cfg['contexts']:
        g['attn0/' + nm] = extra['attn'][0][nm]
    with torch.no_grad():
        g['log_probs_last'] = dec.get_normalized_probs((out[:, -1:].detach(), None), True).numpy()
    ctx_d = {k: v.detach() for k, v in ctx.items()}
    ocfg = synth.oracle_cfg(cfg)
    with torch.no_grad():
        ro, _ = restate.decoder_forward(inp, ctx_d, sd, ocfg)
        d = (ro - out.detach()).abs().max().item()
        assert d < 5e-5, d
        rl, rn, rf = restate.adaptive_loss(ro, tgt, sd, ocfg['cutoffs'])
        assert abs(rl.item() - loss.item()) < 1e-3 * abs(loss.item()) and rn == n
    print('decoder_full fwd/bwd %.1fs  loss %.5f ntokens %d  restate diff %.2e'
          % (time.time() - t0, final.item(), n, d))
    t0 = time.time()
    M = ref_loader.load_model_module()
    stub = types.SimpleNamespace(decoder=dec, index='roberta', sampling_topk=1, sampling_temp=1.0,
                                 padding_idx=1)
    with torch.no_grad():
        lp, ids, _ = M.TransformerFacesObjectModel._generate(stub, cap[:, 0:1].clone(), ctx_d)
    g['greedy_ids'], g['greedy_lp'] = ids.numpy(), lp.numpy()
    print('reference _generate: %d columns in %.1fs' % (ids.shape[1], time.time() - t0))
    with torch.no_grad():
        rids, rlp = restate.greedy_generate(cap[:, 0:1], ctx_d, sd, ocfg, gen_len=100)
    assert torch.equal(rids, ids), 'restatement greedy mismatch'
    assert (rlp - lp).abs().max() < 1e-3
    # top-1 / top-2 margin at every emitted position (teacher-forced along the reference's path)
    margins = []
    with torch.no_grad():
        state, prev = {}, cap[:, 0:1]
        for t in range(ids.shape[1] - 1):
            o, _ = dec({'roberta': prev}, ctx_d, incremental_state=state)
            l = dec.get_normalized_probs((o[:, -1:], None), True)[:, 0]
            top2 = l.topk(2).values
            margins.append((top2[:, 0] - top2[:, 1]).numpy())
            prev = ids[:, t + 1:t + 2]
    m = np.stack(margins, 1)                       # [B, steps]
    m[ids[:, 1:].numpy() == 1] = np.inf            # retired rows
    g['greedy_margins'] = m.astype(np.float32)
    g['greedy_margin_min'] = np.array([m.min()])
    np.savez_compressed(os.path.join(OUT, 'decoder_full.npz'), **g)
    print('decoder_full written; min greedy margin %.3e' % m.min())


RESNET_SEED, RESNET_IMG_SEED, RESNET_BN3_GAIN = 3, 11, 0.25


def resnet_image(B=2, hw=224, seed=RESNET_IMG_SEED):
    rs = np.random.RandomState(seed)
    return torch.from_numpy(rs.standard_normal((B, 3, hw, hw)).astype(np.float32))


def resnet152():
    ref_loader.load()
    from torchvision.models.resnet import Bottleneck
    from tell.models.resnet import ResNetFeatureExtractor
    sd = synth.resnet_state_dict((3, 8, 36, 3), seed=RESNET_SEED, bn3_gain=RESNET_BN3_GAIN)
    net = ResNetFeatureExtractor(Bottleneck, [3, 8, 36, 3])
    net.load_state_dict(sd, strict=True)
    img = resnet_image()
    g = {}
    t0 = time.time()
    with torch.no_grad():
        net.eval()
        y_eval = net(img)
        r = restate.resnet152_forward(img, sd, prefix='')
        d = (r - y_eval).abs().max().item()
        assert d < 1e-4 * max(1.0, y_eval.abs().max().item()), d
        net.train()
        y_train = net(img)
        r2, stats = restate.resnet152_forward(img, sd, prefix='', bn_mode='batch', return_stats=True)
        d2 = (r2 - y_train).abs().max().item()
        assert d2 < 1e-3 * max(1.0, y_train.abs().max().item()), d2   # fp32 vs fp64 differ by 3e-4 here
    after = net.state_dict()
    for k in ('bn1', 'layer1.0.bn3', 'layer3.17.bn2', 'layer4.2.bn3', 'layer2.0.downsample.1'):
        for leaf in ('running_mean', 'running_var'):
            g['after_train/%s.%s' % (k, leaf)] = after['%s.%s' % (k, leaf)].numpy()
            assert torch.allclose(stats['%s.%s' % (k, leaf)], after['%s.%s' % (k, leaf)],
                                  rtol=1e-4, atol=1e-5), (k, leaf)
    assert int(after['bn1.num_batches_tracked']) == 1
    g['y_eval'], g['y_train'] = y_eval.numpy(), y_train.numpy()
    np.savez_compressed(os.path.join(OUT, 'resnet152.npz'), **g)
    print('resnet152 %.1fs: eval |y|max %.3f (restate diff %.1e), train |y|max %.3f (restate diff %.1e)'
          % (time.time() - t0, y_eval.abs().max().item(), d, y_train.abs().max().item(), d2))


ROBERTA_SEED, ROBERTA_IDS_SEED = 5, 21
ROBERTA_SHAPE = dict(B=3, S=200)


def roberta_ids():
    rs = np.random.RandomState(ROBERTA_IDS_SEED)
    ids = synth.article_batch(ROBERTA_SHAPE['B'], ROBERTA_SHAPE['S'], 50265, rs, min_len=40)
    ids[1, 7:] = 1                                   # a 7-token article
    ids[1, 6] = 2
    return ids


def hf_from_fairseq(sd, n_layers, E):
    """fairseq `decoder.sentence_encoder.*` keys -> HF RobertaModel keys (same tensors)."""
    p = 'decoder.sentence_encoder.'
    out = {'embeddings.word_embeddings.weight': sd[p + 'embed_tokens.weight'],
           'embeddings.position_embeddings.weight': sd[p + 'embed_positions.weight'],
           'embeddings.token_type_embeddings.weight': torch.zeros(1, E),
           'embeddings.LayerNorm.weight': sd[p + 'emb_layer_norm.weight'],
           'embeddings.LayerNorm.bias': sd[p + 'emb_layer_norm.bias']}
    for i in range(n_layers):
        lp, hp = p + 'layers.%d.' % i, 'encoder.layer.%d.' % i
        w, b = sd[lp + 'self_attn.in_proj_weight'], sd[lp + 'self_attn.in_proj_bias']
        for j, nm in enumerate(('query', 'key', 'value')):
            out[hp + 'attention.self.%s.weight' % nm] = w[j * E:(j + 1) * E]
            out[hp + 'attention.self.%s.bias' % nm] = b[j * E:(j + 1) * E]
        out[hp + 'attention.output.dense.weight'] = sd[lp + 'self_attn.out_proj.weight']
        out[hp + 'attention.output.dense.bias'] = sd[lp + 'self_attn.out_proj.bias']
        out[hp + 'attention.output.LayerNorm.weight'] = sd[lp + 'self_attn_layer_norm.weight']
        out[hp + 'attention.output.LayerNorm.bias'] = sd[lp + 'self_attn_layer_norm.bias']
        out[hp + 'intermediate.dense.weight'] = sd[lp + 'fc1.weight']
        out[hp + 'intermediate.dense.bias'] = sd[lp + 'fc1.bias']
        out[hp + 'output.dense.weight'] = sd[lp + 'fc2.weight']
        out[hp + 'output.dense.bias'] = sd[lp + 'fc2.bias']
        out[hp + 'output.LayerNorm.weight'] = sd[lp + 'final_layer_norm.weight']
        out[hp + 'output.LayerNorm.bias'] = sd[lp + 'final_layer_norm.bias']
    return out


def roberta_large():
    from transformers import RobertaConfig, RobertaModel
    L, E, H, FFN, V, P = 24, 1024, 16, 4096, 50265, 514
    t0 = time.time()
    sd = synth.roberta_state_dict(L, E, FFN, V, P, seed=ROBERTA_SEED)
    cfg = RobertaConfig(vocab_size=V, hidden_size=E, num_hidden_layers=L, num_attention_heads=H,
                        intermediate_size=FFN, max_position_embeddings=P, type_vocab_size=1,
                        layer_norm_eps=1e-5, pad_token_id=1, bos_token_id=0, eos_token_id=2,
                        hidden_act='gelu', hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    hf = RobertaModel(cfg, add_pooling_layer=False).eval()
    missing, unexpected = hf.load_state_dict(hf_from_fairseq(sd, L, E), strict=False)
    assert not unexpected and all('position_ids' in m or 'token_type_ids' in m for m in missing), \
        (missing, unexpected)
    ids = roberta_ids()
    real = ids != 1
    with torch.no_grad():
        hs = hf(input_ids=ids, attention_mask=real.long(), output_hidden_states=True).hidden_states
        rr = restate.roberta_forward(ids, sd, L, H, prefix='')
    assert len(hs) == L + 1
    worst = 0.0
    for a, b in zip(hs, rr):
        worst = max(worst, (a - b).abs()[real].max().item())
    assert worst < 2e-4, worst
    B, S = ids.shape
    rs = np.random.RandomState(ROBERTA_IDS_SEED + 1)
    pos = np.zeros((B, 8), dtype=np.int64)
    for b in range(B):
        n = int(real[b].sum())
        pos[b] = np.concatenate([[0, n - 1], rs.randint(0, n, size=6)])
    g = dict(sample_pos=pos)
    H_ = torch.stack(hs)                                              # [25,B,S,E]
    g['hidden_at_pos'] = np.stack([H_[:, b, pos[b]].numpy() for b in range(B)], 1)   # [25,B,8,E]
    m = real.unsqueeze(0).unsqueeze(-1).float()
    g['layer_norms'] = ((H_ * m) ** 2).sum(dim=(2, 3)).sqrt().numpy()  # [25,B]
    g['layer_absmax'] = (H_ * m).abs().amax(dim=(2, 3)).numpy()
    bw = torch.from_numpy(np.random.RandomState(9).random_sample(L + 1).astype(np.float32))
    g['bert_weight'] = bw.numpy()
    w = torch.softmax(bw, 0)                                           # :355-364 layer mix
    mix = (H_ * w.view(-1, 1, 1, 1)).sum(0)
    g['mix'] = (mix * real.unsqueeze(-1)).numpy().astype(np.float16)
    np.savez_compressed(os.path.join(OUT, 'roberta_large.npz'), **g)
    print('roberta_large %.1fs: HF vs restatement %.2e; |h24|max %.2f'
          % (time.time() - t0, worst, hs[-1].abs().max().item()))


if __name__ == '__main__':
    torch.manual_seed(0)
    which = sys.argv[1:] or ['decoder', 'resnet', 'roberta']
    if 'resnet' in which:
        resnet152()
    if 'roberta' in which:
        roberta_large()
    if 'decoder' in which:
        decoder_full()
