"""TEST INFRASTRUCTURE -- generates tests/golden/*.npz by running the reference's own, unmodified
modules (imported from /root/reference through oracle/ref_loader.py) on seeded synthetic weights
and inputs (transform-and-tell_b200/tell_b200/synth.py).  Run in the build container only:

    python oracle/gen_golden.py

The committed vectors pin oracle/restate.py (tests/test_oracle_cpu.py) and, through it and
directly, the CUDA path (tests/test_decoder_gpu.py).  /root/reference is never read at test time.
"""
import os
import sys
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
SHAPES = dict(B=3, T=9, S=11, F=3, O=4, P=5)


def ref_decoder(cfg, kind, sd):
    dec = ref_loader.build_decoder(kind, vocab_size=cfg['vocab'], embed_dim=cfg['embed_dim'],
                                   heads=cfg['heads'], ffn=cfg['ffn'], kernels=cfg['kernels'],
                                   cutoff=cfg['cutoffs'])
    missing, unexpected = dec.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all('_float_tensor' in m or 'version' in m for m in missing), missing
    return dec


def grads_summary(named_params):
    out = {}
    for n, p in named_params:
        g = p.grad if p.grad is not None else torch.zeros_like(p)
        out[n] = np.array([g.double().norm().item(), g.double().sum().item()], dtype=np.float64)
    return out


def decoder_goldens(cfg, kind, tag, gain):
    sd = synth.decoder_state_dict(cfg, seed=0, logit_gain=gain)
    cap, ctx = synth.decoder_inputs(cfg, SHAPES['B'], SHAPES['T'], SHAPES['S'], SHAPES['F'],
                                    SHAPES['O'], SHAPES['P'], seed=1234)
    inp, tgt = cap[:, :-1].contiguous(), cap[:, 1:].contiguous()
    dec = ref_decoder(cfg, kind, sd).eval()
    crit = ref_loader.adaptive_loss()
    ctx['article'].requires_grad_(True)
    out, extra = dec({'roberta': inp}, ctx)
    loss, n = crit(dec.adaptive_softmax, (out, None), tgt)
    final = loss / np.log(2) / n
    final.backward()
    g = {}
    g['dec_out'] = out.detach().numpy()
    g['loss_sum'] = loss.detach().numpy()
    g['ntokens'] = np.array([n])
    g['loss'] = final.detach().numpy()
    g['d_article'] = ctx['article'].grad.numpy()
    for k, v in grads_summary(dec.named_parameters()).items():
        g['gsum/' + k] = v
    full = ['layers.0.conv.weight_linear.weight', 'layers.0.linear1.weight_g',
            'layers.1.context_attns.article.bias_k', 'layers.0.conv_layer_norm.weight',
            'adaptive_softmax.head.class_proj.weight', 'layers.1.fc2.bias']
    params = dict(dec.named_parameters())
    for k in full:
        g['gfull/' + k] = params[k].grad.numpy()
    names = [nm for nm, _ in cfg['contexts']]
    if kind == 'faces_objects':
        for nm in names:
            g['attn0/' + nm] = extra['attn'][0][nm]
    # log-probs of the last position
    with torch.no_grad():
        g['log_probs_last'] = dec.get_normalized_probs((out[:, -1:].detach(), None), True).numpy()
    # incremental decoding, one token at a time (teacher forced)
    with torch.no_grad():
        state = {}
        steps = []
        ctx_d = {k: v.detach() for k, v in ctx.items()}
        for t in range(inp.shape[1]):
            o, _ = dec({'roberta': inp[:, t:t + 1]}, ctx_d, incremental_state=state)
            steps.append(o)
        g['dec_out_incremental'] = torch.cat(steps, 1).numpy()
    # restatement must agree with the reference before anything is written
    ocfg = synth.oracle_cfg(cfg)
    with torch.no_grad():
        ro, _ = restate.decoder_forward(inp, ctx_d, sd, ocfg)
        assert (ro - out.detach()).abs().max() < 2e-5, (ro - out.detach()).abs().max()
        rl, rn, rf = restate.adaptive_loss(ro, tgt, sd, ocfg['cutoffs'])
        assert abs(rl.item() - loss.item()) < 1e-3 and rn == n
    if kind == 'faces_objects':
        M = ref_loader.load_model_module()
        stub = types.SimpleNamespace(decoder=dec, index='roberta', sampling_topk=1,
                                     sampling_temp=1.0, padding_idx=1)
        with torch.no_grad():
            lp, ids, _ = M.TransformerFacesObjectModel._generate(stub, cap[:, 0:1].clone(), ctx_d)
        g['greedy_ids'] = ids.numpy()
        g['greedy_lp'] = lp.numpy()
        with torch.no_grad():
            rids, rlp = restate.greedy_generate(cap[:, 0:1], ctx_d, sd, ocfg, gen_len=100)
        assert torch.equal(rids, ids), 'restatement greedy mismatch'
        assert (rlp - lp).abs().max() < 1e-4
        # top-1 / top-2 margin along the greedy path of the reference (documents robustness)
        margins = []
        with torch.no_grad():
            state = {}
            prev = cap[:, 0:1]
            for t in range(ids.shape[1] - 1):
                o, _ = dec({'roberta': prev}, ctx_d, incremental_state=state)
                l = dec.get_normalized_probs((o[:, -1:], None), True)[:, 0]
                top2 = l.topk(2).values
                margins.append((top2[:, 0] - top2[:, 1]).numpy())
                prev = ids[:, t + 1:t + 2]
        g['greedy_margin_min'] = np.array([np.min(np.stack(margins)[ids[:, 1:].t().numpy() != 1])])
    np.savez_compressed(os.path.join(OUT, 'decoder_%s.npz' % tag), **g)
    print(tag, 'loss', final.item(), 'ntokens', n, 'greedy margin',
          g.get('greedy_margin_min'), 'gen steps', g.get('greedy_ids', np.zeros((1, 1))).shape)


def forward_glue_goldens():
    """_forward (transformer_faces_objects.py:311-397) with injected encoders: pins the caption
    shift, the RoBERTa layer mix, the NaN masking and the context layouts."""
    M = ref_loader.load_model_module()
    rs = np.random.RandomState(7)
    B, S, L, P = 2, 4, 25, 2
    cfg = synth.CFG_TINY
    cap = synth.caption_batch(B, 8, cfg['vocab'], rs, cutoffs=cfg['cutoffs'])
    art = synth.article_batch(B, S, cfg['vocab'], rs)
    image = torch.from_numpy(rs.standard_normal((B, 3, 8, 8)).astype(np.float32))
    feats = torch.from_numpy(rs.standard_normal((B, 2048, P, P)).astype(np.float32))
    hid = [torch.from_numpy(rs.standard_normal((B, S, 1024)).astype(np.float32)) for _ in range(L)]
    faces = synth.nan_padded(B, 3, 512, rs, 'faces')
    objs = synth.nan_padded(B, 4, 2048, rs, 'obj')
    bert_weight = torch.from_numpy(rs.random_sample(L).astype(np.float32))
    roberta = types.SimpleNamespace(extract_features=lambda ids, return_all_hiddens: hid)
    stub = types.SimpleNamespace(index='roberta', padding_idx=1, resnet=lambda im: feats,
                                 roberta=roberta, weigh_bert=True, bert_weight=bert_weight)
    caption = {'roberta': cap.clone()}
    f_in, o_in = faces.clone(), objs.clone()
    with torch.no_grad():
        cap_ids, tgt_ids, contexts = M.TransformerFacesObjectModel._forward(
            stub, {'roberta': art}, image, caption, f_in, o_in)
    g = dict(cap=cap.numpy(), art=art.numpy(), feats=feats.numpy(), hid=torch.stack(hid).numpy(),
             faces=faces.numpy(), objs=objs.numpy(), bert_weight=bert_weight.numpy(),
             caption_ids=cap_ids.numpy(), target_ids=tgt_ids.numpy())
    for k, v in contexts.items():
        if v is not None:
            g['ctx/' + k] = v.numpy()
    rc = restate.build_contexts(feats, hid, bert_weight, art, faces.clone(), objs.clone())
    for k in rc:
        assert torch.allclose(rc[k].float(), contexts[k].float(), atol=1e-6), k
    ri, rt = restate.shift_caption(cap)
    assert torch.equal(ri, cap_ids) and torch.equal(rt, tgt_ids)
    np.savez_compressed(os.path.join(OUT, 'forward_glue.npz'), **g)
    print('forward_glue ok', {k: v.shape for k, v in contexts.items() if v is not None})


def op_goldens():
    ref_loader.load()
    from tell.modules import DynamicConv1dTBC, MultiHeadAttention
    rs = np.random.RandomState(3)
    g = {}
    for (T, B, C, H, K) in [(5, 2, 64, 4, 15), (12, 2, 64, 4, 7)]:
        m = DynamicConv1dTBC(C, K, padding_l=K - 1, num_heads=H, weight_softmax=True).eval()
        w = torch.from_numpy(rs.standard_normal((H * K, C)).astype(np.float32) / 8)
        m.weight_linear.weight.data.copy_(w)
        x = torch.from_numpy(rs.standard_normal((T, B, C)).astype(np.float32))
        with torch.no_grad():
            y = m(x)
        tag = 'dynconv_T%d_K%d/' % (T, K)
        g[tag + 'x'], g[tag + 'w'], g[tag + 'y'] = x.numpy(), w.numpy(), y.numpy()
        assert (restate.dynamic_conv(x, w, K, H) - y).abs().max() < 1e-5
    # attention over an empty context ([B,1,0] faces, multi_head.py:349-364)
    E, H = 64, 4
    m = MultiHeadAttention(E, H, kdim=512, vdim=512).eval()
    sd = {('a.' + k): v.detach().clone() for k, v in m.state_dict().items()}
    q = torch.from_numpy(rs.standard_normal((4, 2, E)).astype(np.float32))
    key = torch.zeros(1, 2, 0)
    mask = torch.zeros(2, 1, dtype=torch.bool)
    with torch.no_grad():
        y, w = m(q, key, key, key_padding_mask=mask, static_kv=True, need_weights=True)
    for k, v in sd.items():
        g['mha_empty/' + k] = v.numpy()
    g['mha_empty/q'], g['mha_empty/y'], g['mha_empty/w'] = q.numpy(), y.numpy(), w.numpy()
    ry, rw = restate.multi_head_attention(q, key, mask, sd, 'a.', H, True)
    assert (ry - y).abs().max() < 1e-5 and (rw - w).abs().max() < 1e-6
    # LightweightConv1dTBC (lightweight.py:88-240): static taps [H,1,K] (+ optional channel bias)
    from tell.modules import LightweightConv1dTBC
    for (T, B, C, H, K, bias) in [(5, 2, 64, 4, 7, False), (12, 3, 64, 8, 3, False), (4, 2, 32, 2, 15, False)]:
        m = LightweightConv1dTBC(C, K, padding_l=K - 1, num_heads=H, weight_softmax=True, bias=bias).eval()
        w = torch.from_numpy(rs.standard_normal((H, 1, K)).astype(np.float32))
        m.weight.data.copy_(w)
        if bias:
            m.bias.data.copy_(torch.from_numpy(rs.standard_normal(C).astype(np.float32)))
        x = torch.from_numpy(rs.standard_normal((T, B, C)).astype(np.float32))
        with torch.no_grad():
            y = m(x)
            state, steps = {}, []
            for t in range(T):
                steps.append(m(x[t:t + 1], incremental_state=state))
            yi = torch.cat(steps, 0)
        assert (yi - y).abs().max() < 1e-5
        tag = 'lightconv_T%d_K%d/' % (T, K)
        g[tag + 'x'], g[tag + 'w'], g[tag + 'y'] = x.numpy(), w.numpy(), y.numpy()
        if bias:
            g[tag + 'bias'] = m.bias.detach().numpy()
        ry = restate.lightweight_conv(x, w, K, H, bias=m.bias.detach() if bias else None)
        assert (ry - y).abs().max() < 1e-5
    np.savez_compressed(os.path.join(OUT, 'ops.npz'), **g)
    print('ops ok')


def facenet_goldens():
    """tell/facenet: the reference's own InceptionResnetV1 (seeded synthetic weights -- the VGGFace2
    checkpoint is a network download) and P/R/O-Net with BOTH seeded synthetic weights and the real
    vendored checkpoints tell/facenet/data/{pnet,rnet,onet}.pt (copied into the fixture as data so
    the GPU box, which has no /root/reference, can load them)."""
    from tell.facenet.inception_resnet_v1 import InceptionResnetV1
    from tell.facenet.mtcnn import ONet, PNet, RNet
    g = {}
    rs = np.random.RandomState(7)
    net = InceptionResnetV1(num_classes=10).eval()
    sd = synth.shaped_state_dict({k: v.shape for k, v in net.state_dict().items()}, seed=3)
    net.load_state_dict(sd, strict=True)
    x = torch.from_numpy(rs.standard_normal((3, 3, 160, 160)).astype(np.float32))
    x = x * torch.tensor([0.6, 1.0, 1.6]).view(3, 1, 1, 1) + torch.tensor([-0.5, 0.0, 0.7]).view(3, 1, 1, 1)
    with torch.no_grad():
        emb, logits = net(x)
        o_emb, o_logits = restate.inception_resnet_v1_forward(x, sd)
    assert (emb - o_emb).abs().max() < 1e-5 and (logits - o_logits).abs().max() < 1e-4
    g.update(irv1_x=x.numpy(), irv1_emb=emb.numpy(), irv1_logits=logits.numpy())
    sizes = {'pnet': (2, 3, 37, 53), 'rnet': (5, 3, 24, 24), 'onet': (4, 3, 48, 48)}
    fwd = {'pnet': restate.pnet_forward, 'rnet': restate.rnet_forward, 'onet': restate.onet_forward}
    for name, cls in (('pnet', PNet), ('rnet', RNet), ('onet', ONet)):
        x = torch.from_numpy(rs.standard_normal(sizes[name]).astype(np.float32))
        g[name + '_x'] = x.numpy()
        for kind in ('synth', 'real'):
            net = cls(pretrained=(kind == 'real')).eval()
            if kind == 'synth':
                sd = synth.shaped_state_dict({k: v.shape for k, v in net.state_dict().items()}, seed=5)
                net.load_state_dict(sd, strict=True)
            else:
                sd = {k: v.clone() for k, v in net.state_dict().items()}
                for k, v in sd.items():
                    g['%s_real_w.%s' % (name, k)] = v.numpy()
            with torch.no_grad():
                outs = net(x)
                o_outs = fwd[name](x, sd)
            for i, (a, b) in enumerate(zip(outs, o_outs)):
                assert (a - b).abs().max() < 1e-5, (name, kind, i)
                g['%s_%s_out%d' % (name, kind, i)] = a.numpy()
    np.savez_compressed(os.path.join(OUT, 'facenet.npz'), **g)
    print('facenet goldens:', len(g), 'arrays')


if __name__ == '__main__':
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)
    if '--facenet-only' in sys.argv:
        ref_loader.load()
        facenet_goldens()
        sys.exit(0)
    op_goldens()
    forward_glue_goldens()
    decoder_goldens(synth.CFG_TINY, 'faces_objects', 'tiny_faces_objects', gain=4.0)
    decoder_goldens(synth.CFG_TINY_NO_IMAGE, 'no_image', 'tiny_no_image', gain=4.0)
    decoder_goldens(synth.CFG_TINY_FLATTENED, 'flattened', 'tiny_flattened', gain=4.0)
    decoder_goldens(synth.CFG_TINY_FACES, 'faces_parallel', 'tiny_faces_parallel', gain=4.0)
    facenet_goldens()
