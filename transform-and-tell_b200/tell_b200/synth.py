"""Deterministic synthetic weights and NYTimes-shaped batches.

numpy RandomState streams are stable across machines, so the same (config, seed) reproduces the
same state dict in the build container (where the reference modules are loaded with it to produce
golden vectors) and on the GPU box (where those vectors are checked).  Key names and shapes follow
the reference decoder state dict.
"""
import math

import numpy as np
import torch

CFG_FULL = dict(vocab=50265, embed_dim=1024, heads=16, ffn=4096, kernels=(3, 7, 15, 31),
                cutoffs=(5000, 20000), max_pos=512,
                contexts=(('image', 2048), ('article', 1024), ('faces', 512), ('obj', 2048)))
CFG_NO_IMAGE = dict(CFG_FULL, contexts=(('article', 1024),))
CFG_TINY = dict(vocab=2000, embed_dim=64, heads=4, ffn=128, kernels=(3, 7), cutoffs=(200, 800),
                max_pos=512,
                contexts=(('image', 2048), ('article', 1024), ('faces', 512), ('obj', 2048)))
CFG_TINY_NO_IMAGE = dict(CFG_TINY, contexts=(('article', 1024),))
CFG_TINY_FLATTENED = dict(CFG_TINY, contexts=(('image', 2048), ('article', 1024)))
CFG_TINY_FACES = dict(CFG_TINY, contexts=(('image', 2048), ('article', 1024), ('faces', 512)))


def full_cutoffs(cfg):
    return list(cfg['cutoffs']) + [cfg['vocab']]


def oracle_cfg(cfg):
    """The dict oracle/restate.py's decoder functions take."""
    return dict(embed_dim=cfg['embed_dim'], heads=cfg['heads'], kernels=list(cfg['kernels']),
                cutoffs=full_cutoffs(cfg), ctx_names=[n for n, _ in cfg['contexts']])


def sinusoidal_table(n_embeds, embed_dim, padding_idx):
    n_ts = embed_dim // 2
    increment = math.log(10000.0) / (n_ts - 1)
    inv = torch.exp(torch.arange(n_ts, dtype=torch.float) * -increment)
    st = torch.arange(n_embeds, dtype=torch.float).unsqueeze(1) * inv.unsqueeze(0)
    sig = torch.cat([torch.sin(st), torch.cos(st)], dim=1)
    sig[padding_idx, :] = 0
    return sig


def decoder_state_dict(cfg, seed=0, logit_gain=1.0):
    """Reference-shaped decoder weights.  logit_gain > 1 scales the output-side embeddings so that
    greedy argmax margins are comfortably above the numerical noise of either implementation."""
    rs = np.random.RandomState(seed)

    def n(*shape, std=1.0):
        return torch.from_numpy((rs.standard_normal(shape) * std).astype(np.float32))

    def xavier(o, i):
        return n(o, i, std=math.sqrt(2.0 / (o + i)))

    E, H, V = cfg['embed_dim'], cfg['heads'], cfg['vocab']
    cut = full_cutoffs(cfg)
    sd = {}
    pre = 'embedder.token_embedder_adaptive.embeddings.'
    for i, hi in enumerate(cut):
        lo = cut[i - 1] if i else 0
        w = n(hi - lo, E, std=logit_gain / math.sqrt(E))
        w[0] = 0                                           # nn.Embedding(padding_idx=0)
        sd[pre + '%d.0.weight' % i] = w
        sd[pre + '%d.1.weight' % i] = xavier(E, E)
    sd['embedder.token_embedder_position.weights'] = sinusoidal_table(cfg['max_pos'] + 1, E, 1)

    def gehring(prefix, o, i):
        v = n(o, i, std=math.sqrt(0.9 / i))
        sd[prefix + 'bias'] = n(o, std=0.02)
        sd[prefix + 'weight_g'] = v.norm(dim=1, keepdim=True) * (1 + 0.1 * n(o, 1))
        sd[prefix + 'weight_v'] = v

    def ln(prefix):
        sd[prefix + 'weight'] = 1 + 0.1 * n(E)
        sd[prefix + 'bias'] = 0.05 * n(E)

    for li, K in enumerate(cfg['kernels']):
        p = 'layers.%d.' % li
        gehring(p + 'linear1.', 2 * E, E)
        sd[p + 'conv.weight_linear.weight'] = xavier(H * K, E)
        gehring(p + 'linear2.', E, E)
        ln(p + 'conv_layer_norm.')
        for name, kdim in cfg['contexts']:
            a = p + 'context_attns.%s.' % name
            if kdim == E:
                sd[a + 'in_proj_weight'] = xavier(3 * E, E)
            else:
                sd[a + 'k_proj_weight'] = xavier(E, kdim)
                sd[a + 'v_proj_weight'] = xavier(E, kdim)
                sd[a + 'q_proj_weight'] = xavier(E, E)
            sd[a + 'in_proj_bias'] = n(3 * E, std=0.02)
            sd[a + 'bias_k'] = n(1, 1, E, std=math.sqrt(2.0 / (1 + E)))
            sd[a + 'bias_v'] = n(1, 1, E, std=math.sqrt(2.0 / (1 + E)))
            sd[a + 'out_proj.weight'] = xavier(E, E)
            sd[a + 'out_proj.bias'] = n(E, std=0.02)
            ln(p + 'context_attn_lns.%s.' % name)
        gehring(p + 'context_fc.', E, E * len(cfg['contexts']))
        gehring(p + 'fc1.', cfg['ffn'], E)
        gehring(p + 'fc2.', E, cfg['ffn'])
        ln(p + 'final_layer_norm.')
    sd['adaptive_softmax.head.word_proj.weight'] = sd[pre + '0.0.weight']
    sd['adaptive_softmax.head.class_proj.weight'] = xavier(len(cut) - 1, E) * logit_gain
    sd['adaptive_softmax.head._float_tensor'] = torch.zeros(1)
    for i in range(len(cut) - 1):
        sd['adaptive_softmax.tail.%d.0.weight' % i] = xavier(E, E)
        sd['adaptive_softmax.tail.%d.2.weight' % i] = sd[pre + '%d.0.weight' % (i + 1)]
    sd['adaptive_softmax.version'] = torch.LongTensor([1])
    sd['version'] = torch.Tensor([2])
    return sd


def caption_batch(B, T, vocab, rs, min_len=None, cutoffs=(5000, 20000)):
    """[B,T] ids: col0 = <s>=0, body Zipf-like over the three bands (80/15/5 %), </s>=2 last real
    token, pad=1 afterwards."""
    min_len = min_len or max(3, T // 2)
    ids = np.ones((B, T), dtype=np.int64)
    for b in range(B):
        L = rs.randint(min_len, T + 1)
        band = rs.choice(3, size=L, p=[0.8, 0.15, 0.05])
        lo = np.array([4, cutoffs[0], cutoffs[1]])[band]
        hi = np.array([cutoffs[0], cutoffs[1], vocab])[band]
        body = (lo + (rs.random_sample(L) ** 2 * (hi - lo))).astype(np.int64)
        ids[b, :L] = body
        ids[b, 0] = 0
        ids[b, L - 1] = 2
    return torch.from_numpy(ids)


def article_batch(B, S, vocab, rs, min_len=None):
    min_len = min_len or max(2, S // 2)
    ids = np.ones((B, S), dtype=np.int64)
    for b in range(B):
        L = rs.randint(min_len, S + 1)
        ids[b, :L] = rs.randint(4, vocab, size=L)
        ids[b, 0] = 0
        ids[b, L - 1] = 2
    return torch.from_numpy(ids)


def nan_padded(B, n_max, dim, rs, kind):
    """faces: unit-norm rows; objects: relu(N(0,1)); per-sample count U{0..n_max}, NaN padded."""
    x = np.full((B, n_max, dim), np.nan, dtype=np.float32)
    for b in range(B):
        k = rs.randint(0, n_max + 1)
        v = rs.standard_normal((k, dim)).astype(np.float32)
        if kind == 'faces':
            v /= np.linalg.norm(v, axis=1, keepdims=True) + 1e-12
        else:
            v = np.maximum(v, 0)
        x[b, :k] = v
    return torch.from_numpy(x)


def decoder_inputs(cfg, B, T, S, F=4, O=16, P=49, seed=1234):
    """Decoder-level synthetic inputs: caption ids [B,T+1] and the four contexts in the layout
    transformer_faces_objects.py:384-395 hands to the decoder ([len,B,dim] + [B,len] masks)."""
    rs = np.random.RandomState(seed)
    names = [n for n, _ in cfg['contexts']]
    cap = caption_batch(B, T + 1, cfg['vocab'], rs, cutoffs=cfg['cutoffs'])
    art_ids = article_batch(B, S, cfg['vocab'], rs)
    contexts = {}
    if 'image' in names:
        contexts['image'] = torch.from_numpy(
            np.maximum(rs.standard_normal((P, B, 2048)), 0).astype(np.float32))
        contexts['image_mask'] = torch.zeros(B, P, dtype=torch.bool)
    contexts['article'] = torch.from_numpy(rs.standard_normal((S, B, 1024)).astype(np.float32))
    contexts['article_mask'] = art_ids == 1
    if 'faces' in names:
        f = nan_padded(B, F, 512, rs, 'faces')
        m = torch.isnan(f).any(-1)
        f[m] = 0
        contexts['faces'] = f.transpose(0, 1).contiguous()
        contexts['faces_mask'] = m
    if 'obj' in names:
        o = nan_padded(B, O, 2048, rs, 'obj')
        m = torch.isnan(o).any(-1)
        o[m] = 0
        contexts['obj'] = o.transpose(0, 1).contiguous()
        contexts['obj_mask'] = m
    return cap, contexts


def resnet_state_dict(layers=(3, 8, 36, 3), seed=0, bn3_gain=0.5):
    """torchvision-keyed ResNet weights with non-trivial BatchNorm statistics.  With running
    statistics that do not track the activations the residual stream grows by about
    sqrt(1 + bn3_gain^2) per block in eval mode (0.5 -> ~1e5 after 50 blocks; the full-depth golden
    uses 0.25)."""
    rs = np.random.RandomState(seed)

    def n(*shape, std=1.0):
        return torch.from_numpy((rs.standard_normal(shape) * std).astype(np.float32))

    sd = {}

    def conv(name, cout, cin, k):
        sd[name + '.weight'] = n(cout, cin, k, k, std=math.sqrt(2.0 / (cin * k * k)))

    def bn(name, c, gain=1.0):
        sd[name + '.weight'] = gain * (1 + 0.1 * n(c))
        sd[name + '.bias'] = 0.05 * n(c)
        sd[name + '.running_mean'] = 0.1 * n(c)
        sd[name + '.running_var'] = torch.from_numpy((0.5 + rs.random_sample(c)).astype(np.float32))
        sd[name + '.num_batches_tracked'] = torch.tensor(0, dtype=torch.long)

    conv('conv1', 64, 3, 7)
    bn('bn1', 64)
    inplanes = 64
    for li, (planes, nb) in enumerate(zip((64, 128, 256, 512), layers)):
        for bi in range(nb):
            p = 'layer%d.%d.' % (li + 1, bi)
            stride = 2 if (li > 0 and bi == 0) else 1
            conv(p + 'conv1', planes, inplanes, 1); bn(p + 'bn1', planes)
            conv(p + 'conv2', planes, planes, 3); bn(p + 'bn2', planes)
            conv(p + 'conv3', planes * 4, planes, 1); bn(p + 'bn3', planes * 4, gain=bn3_gain)
            if bi == 0 and (stride != 1 or inplanes != planes * 4):
                conv(p + 'downsample.0', planes * 4, inplanes, 1); bn(p + 'downsample.1', planes * 4)
            inplanes = planes * 4
    sd['fc.weight'] = n(1000, 2048, std=0.01)
    sd['fc.bias'] = torch.zeros(1000)
    return sd


def roberta_state_dict(n_layers, embed_dim, ffn, vocab, max_pos, seed=0):
    """fairseq-keyed (decoder.sentence_encoder.*) RoBERTa encoder weights."""
    rs = np.random.RandomState(seed)

    def n(*shape, std=1.0):
        return torch.from_numpy((rs.standard_normal(shape) * std).astype(np.float32))

    E = embed_dim
    p = 'decoder.sentence_encoder.'
    sd = {p + 'embed_tokens.weight': n(vocab, E, std=0.5),
          p + 'embed_positions.weight': n(max_pos, E, std=0.5),
          p + 'emb_layer_norm.weight': 1 + 0.1 * n(E), p + 'emb_layer_norm.bias': 0.05 * n(E)}
    sd[p + 'embed_tokens.weight'][1] = 0
    sd[p + 'embed_positions.weight'][1] = 0
    for i in range(n_layers):
        lp = p + 'layers.%d.' % i
        sd[lp + 'self_attn.in_proj_weight'] = n(3 * E, E, std=1.0 / math.sqrt(E))
        sd[lp + 'self_attn.in_proj_bias'] = n(3 * E, std=0.02)
        sd[lp + 'self_attn.out_proj.weight'] = n(E, E, std=1.0 / math.sqrt(E))
        sd[lp + 'self_attn.out_proj.bias'] = n(E, std=0.02)
        sd[lp + 'self_attn_layer_norm.weight'] = 1 + 0.1 * n(E)
        sd[lp + 'self_attn_layer_norm.bias'] = 0.05 * n(E)
        sd[lp + 'fc1.weight'] = n(ffn, E, std=1.0 / math.sqrt(E))
        sd[lp + 'fc1.bias'] = n(ffn, std=0.02)
        sd[lp + 'fc2.weight'] = n(E, ffn, std=1.0 / math.sqrt(ffn))
        sd[lp + 'fc2.bias'] = n(E, std=0.02)
        sd[lp + 'final_layer_norm.weight'] = 1 + 0.1 * n(E)
        sd[lp + 'final_layer_norm.bias'] = 0.05 * n(E)
    return sd


def shaped_state_dict(shapes, seed=0):
    """Seeded synthetic weights for a convnet given {key: shape} (the keys of the reference module's
    state_dict): He-scaled conv / linear weights, near-identity BatchNorm with non-trivial running
    statistics, PReLU slopes around 0.25.  Every tensor is drawn from its own RandomState seeded by
    (seed, crc32(key)), so the result does not depend on iteration order or on which side (the
    reference module or the B200 module) supplied the shapes."""
    import zlib
    sd = {}
    for key in sorted(shapes):
        shape = tuple(shapes[key])
        rs = np.random.RandomState((seed * 1000003 + zlib.crc32(key.encode())) % (2 ** 31))
        leaf = key.rsplit('.', 1)[-1]
        owner = key.rsplit('.', 2)[-2] if key.count('.') >= 1 else ''

        def n(std=1.0):
            return torch.from_numpy((rs.standard_normal(shape) * std).astype(np.float32))
        if leaf == 'num_batches_tracked':
            t = torch.tensor(0, dtype=torch.long)
        elif leaf == 'running_var':
            t = torch.from_numpy((0.5 + rs.random_sample(shape)).astype(np.float32))
        elif leaf == 'running_mean':
            t = 0.1 * n()
        elif owner.startswith('prelu'):
            t = torch.from_numpy((0.1 + 0.3 * rs.random_sample(shape)).astype(np.float32))
        elif owner.startswith('bn') or owner == 'last_bn':
            t = 1 + 0.1 * n() if leaf == 'weight' else 0.05 * n()
        elif leaf == 'weight':
            fan_in = int(np.prod(shape[1:])) if len(shape) > 1 else shape[0]
            t = n(math.sqrt(2.0 / fan_in))
        else:   # conv / linear bias
            t = 0.05 * n()
        sd[key] = t
    return sd
