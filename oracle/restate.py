"""TEST INFRASTRUCTURE ONLY -- CPU restatement (plain PyTorch fp32) of the reference algorithm
for the caption hot path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this file; the product (transform-and-tell_b200/) never does.

Parity status: PINNED.  Every function below is checked in tests/test_oracle_cpu.py against
golden vectors produced by the reference's own, unmodified modules
(oracle/gen_golden.py imports them from /root/reference via oracle/ref_loader.py; the vectors are
committed under tests/golden/) and against the reference's only known-answer test
(tell/modules/token_embedders/tests/test_positional.py:12-32).
Exceptions, stated in DESIGN.md: the RoBERTa encoder (fairseq, un-vendored) and torchvision's
Bottleneck are third-party code absent from the reference tree -- their restatements here follow
the published architectures and are "parity unpinned".

All functions are stateless and take the reference state_dict (Appendix B key names), so the same
dict feeds the oracle and the CUDA modules.  Citations are relative to the reference repo root.
"""
import math

import torch
import torch.nn.functional as F

LN_EPS = 1e-5


# ------------------------------------------------------------------------------------- embeddings
def make_positions(ids, padding_idx, left_pad):
    """tell/modules/token_embedders/positional.py:231-268."""
    mask = ids.ne(padding_idx)
    T = ids.shape[1]
    positions = torch.arange(padding_idx + 1, padding_idx + 1 + T).expand_as(ids)
    if left_pad:
        positions = positions - (T - mask.long().sum(dim=1, keepdim=True))
    return torch.where(mask, positions, torch.full_like(ids, padding_idx))


def sinusoidal_table(n_embeds, embed_dim, padding_idx):
    """positional.py:124-164 get_embedding (tensor2tensor layout: sin || cos)."""
    n_ts = embed_dim // 2
    increment = math.log(10000.0) / (n_ts - 1)
    inv = torch.exp(torch.arange(n_ts, dtype=torch.float) * -increment)
    st = torch.arange(n_embeds, dtype=torch.float).unsqueeze(1) * inv.unsqueeze(0)
    sig = torch.cat([torch.sin(st), torch.cos(st)], dim=1)
    if embed_dim % 2 == 1:
        sig = torch.cat([sig, torch.zeros(n_embeds, 1)], dim=1)
    if padding_idx is not None:
        sig[padding_idx, :] = 0
    return sig


def positional_embedding(ids, table, padding_idx=1, left_pad=False, start_pos=0):
    """positional.py:167-211 forward (start_pos = incremental position)."""
    pos = make_positions(ids, padding_idx, left_pad)
    pos = torch.where(pos != padding_idx, pos + start_pos, pos)
    return table.index_select(0, pos.reshape(-1)).view(ids.shape[0], ids.shape[1], -1)


def adaptive_embedding(ids, sd, prefix, cutoffs, embed_dim, scale=True):
    """adaptive.py:61-76.  cutoffs includes the vocab size as last entry."""
    out = torch.zeros(ids.shape + (embed_dim,))
    for i, hi in enumerate(cutoffs):
        lo = cutoffs[i - 1] if i > 0 else 0
        mask = (ids >= lo) & (ids < hi)
        if mask.any():
            emb = sd[prefix + 'embeddings.%d.0.weight' % i]
            proj = sd[prefix + 'embeddings.%d.1.weight' % i]
            # nn.Embedding(..., padding_idx=0): local row 0 receives no gradient
            out[mask] = F.linear(F.embedding(ids[mask] - lo, emb, padding_idx=0), proj)
    return (math.sqrt(embed_dim) if scale else 1.0) * out


def sum_embedder(ids, sd, cutoffs, embed_dim, start_pos=0, prefix='embedder.'):
    """sum_text_field_embedder.py:70-118: adaptive + sinusoidal, summed."""
    a = adaptive_embedding(ids, sd, prefix + 'token_embedder_adaptive.', cutoffs, embed_dim)
    p = positional_embedding(ids, sd[prefix + 'token_embedder_position.weights'], 1, False,
                             start_pos)
    return a + p


# ------------------------------------------------------------------------------------- linear
def gehring_linear(x, sd, prefix):
    """linear.py:8-34: weight-norm (dim 0) linear."""
    v, g = sd[prefix + 'weight_v'], sd[prefix + 'weight_g']
    w = v * (g / v.norm(dim=1, keepdim=True))
    return F.linear(x, w, sd.get(prefix + 'bias'))


# ------------------------------------------------------------------------------------- dynconv
def dynamic_conv(x, w_filter, K, H, weight_softmax=True):
    """convolutions/dynamic.py:285-336 in closed form:
    out[t,b,hR+r] = sum_k softmax_k(W_f x_t)[h,k] * x[t-(K-1)+k, b, hR+r], zero for negative time.
    The softmax runs over all K taps even when some hit the padding."""
    T, B, C = x.shape
    R = C // H
    z = F.linear(x, w_filter).view(T, B, H, K)
    p = F.softmax(z, dim=-1) if weight_softmax else z
    xp = torch.cat([x.new_zeros(K - 1, B, C), x], dim=0)
    win = xp.unfold(0, K, 1)                       # [T, B, C, K]
    win = win.view(T, B, H, R, K)
    return torch.einsum('tbhrk,tbhk->tbhr', win, p).reshape(T, B, C)


def lightweight_conv(x, w, K, H, weight_softmax=True, bias=None):
    """convolutions/lightweight.py:88-240 in closed form: static taps w [H,1,K] shared by every
    (t, b); the softmax runs over all K taps and the taps that fall before t = 0 are dropped without
    renormalisation -- both in _forward_expanded (K > T: weight.narrow AFTER the softmax, :197-199)
    and in the incremental path (:174-176)."""
    T, B, C = x.shape
    R = C // H
    p = w.view(H, K)
    p = F.softmax(p, dim=-1) if weight_softmax else p
    xp = torch.cat([x.new_zeros(K - 1, B, C), x], dim=0)
    win = xp.unfold(0, K, 1).view(T, B, H, R, K)
    out = torch.einsum('tbhrk,hk->tbhr', win, p).reshape(T, B, C)
    return out if bias is None else out + bias.view(1, 1, -1)


# ------------------------------------------------------------------------------------- attention
def multi_head_attention(query, key, key_padding_mask, sd, prefix, H, need_weights=False):
    """attention/multi_head.py:288-486, static_kv=True / incremental_state=None path."""
    T, B, E = query.shape
    d = E // H
    bias = sd[prefix + 'in_proj_bias']
    if (prefix + 'in_proj_weight') in sd:
        w = sd[prefix + 'in_proj_weight']
        wq, wk, wv = w[:E], w[E:2 * E], w[2 * E:]
    else:
        wq, wk, wv = (sd[prefix + 'q_proj_weight'], sd[prefix + 'k_proj_weight'],
                      sd[prefix + 'v_proj_weight'])
    q = F.linear(query, wq, bias[:E]) * d ** -0.5
    S = key.shape[0] if key.shape[2] > 0 else 0
    bk = sd[prefix + 'bias_k'].repeat(1, B, 1)
    bv = sd[prefix + 'bias_v'].repeat(1, B, 1)
    if S > 0:
        k = torch.cat([F.linear(key, wk, bias[E:2 * E]), bk])
        v = torch.cat([F.linear(key, wv, bias[2 * E:]), bv])
        mask = torch.cat([key_padding_mask, key_padding_mask.new_zeros(B, 1)], dim=1)
    else:
        k, v = bk, bv
        mask = key_padding_mask.new_zeros(B, 1)
    q = q.contiguous().view(T, B * H, d).transpose(0, 1)
    k = k.contiguous().view(-1, B * H, d).transpose(0, 1)
    v = v.contiguous().view(-1, B * H, d).transpose(0, 1)
    k = torch.cat([k, k.new_zeros(B * H, 1, d)], dim=1)          # add_zero_attn
    v = torch.cat([v, v.new_zeros(B * H, 1, d)], dim=1)
    mask = torch.cat([mask, mask.new_zeros(B, 1)], dim=1)
    L = k.shape[1]
    w = torch.bmm(q, k.transpose(1, 2)).view(B, H, T, L)
    w = w.masked_fill(mask.bool().unsqueeze(1).unsqueeze(2), float('-inf')).view(B * H, T, L)
    w = F.softmax(w, dim=-1, dtype=torch.float32)
    attn = torch.bmm(w, v).transpose(0, 1).contiguous().view(T, B, E)
    attn = F.linear(attn, sd[prefix + 'out_proj.weight'], sd[prefix + 'out_proj.bias'])
    weights = w.view(B, H, T, L).sum(dim=1) / H if need_weights else None
    return attn, weights


# ------------------------------------------------------------------------------------- decoder
def _ln(x, sd, prefix):
    return F.layer_norm(x, (x.shape[-1],), sd[prefix + 'weight'], sd[prefix + 'bias'], LN_EPS)


def decoder_layer(X, contexts, sd, prefix, K, H, ctx_names, conv_state=None):
    """models/decoder_faces_objects.py:255-365 (eval / dropout-free); also covers
    decoder_flattened_no_image.py's layer (ctx_names = ['article']).
    conv_state: None, or a list [buffer] holding the DynamicConv input buffer
    (dynamic.py:95-99 incremental semantics)."""
    residual = X
    h = gehring_linear(X, sd, prefix + 'linear1.')
    h = F.glu(h, dim=-1)
    if conv_state is not None:
        prev = conv_state[0]
        if prev is not None:
            h = torch.cat([prev, h], dim=0)
        conv_state[0] = h[-K + 1:] if K > 1 else h[:0]
        o = dynamic_conv(h, sd[prefix + 'conv.weight_linear.weight'], K, H)
        if prev is not None:
            o = o[prev.shape[0]:]
    else:
        o = dynamic_conv(h, sd[prefix + 'conv.weight_linear.weight'], K, H)
    h = gehring_linear(o, sd, prefix + 'linear2.')
    X = _ln(residual + h, sd, prefix + 'conv_layer_norm.')
    outs, attns = [], {}
    for name in ctx_names:
        a, w = multi_head_attention(X, contexts[name], contexts[name + '_mask'], sd,
                                    prefix + 'context_attns.%s.' % name, H, need_weights=True)
        outs.append(_ln(X + a, sd, prefix + 'context_attn_lns.%s.' % name))
        attns[name] = w
    Xc = torch.cat(outs, dim=-1)
    X = gehring_linear(Xc, sd, prefix + 'context_fc.')
    residual = X
    h = F.relu(gehring_linear(X, sd, prefix + 'fc1.'))
    h = gehring_linear(h, sd, prefix + 'fc2.')
    X = _ln(residual + h, sd, prefix + 'final_layer_norm.')
    return X, attns


def decoder_forward(ids, contexts, sd, cfg, incremental_state=None):
    """decoder_faces_objects.py:95-142.  cfg: dict(embed_dim, heads, kernels, cutoffs (with vocab),
    ctx_names).  incremental_state: None or dict {'pos': int, 'conv': [[buf] per layer]}.
    Returns [B,T,E] and per-layer attention dicts."""
    start = 0
    if incremental_state is not None:
        start = incremental_state.setdefault('pos', 0)
        incremental_state['pos'] = start + ids.shape[1]
        conv = incremental_state.setdefault('conv', [[None] for _ in cfg['kernels']])
    X = sum_embedder(ids, sd, cfg['cutoffs'], cfg['embed_dim'], start)
    X = X.transpose(0, 1)
    attns = []
    for i, K in enumerate(cfg['kernels']):
        X, a = decoder_layer(X, contexts, sd, 'layers.%d.' % i, K, cfg['heads'], cfg['ctx_names'],
                             conv[i] if incremental_state is not None else None)
        attns.append(a)
    return X.transpose(0, 1), attns


# ------------------------------------------------------------------------------------- softmax / loss
def adaptive_logits(X, target, sd, cutoffs, prefix='adaptive_softmax.'):
    """softmax.py:144-191: (list of logits, list of targets); tied weights."""
    X = X.contiguous().view(-1, X.shape[-1])
    target = target.reshape(-1)
    head_w = torch.cat([sd[prefix + 'head.word_proj.weight'], sd[prefix + 'head.class_proj.weight']])
    new_target = [target.clone()]
    logits = [F.linear(X, head_w)]
    for i in range(len(cutoffs) - 1):
        mask = (target >= cutoffs[i]) & (target < cutoffs[i + 1])
        new_target[0][mask] = cutoffs[0] + i
        if mask.any():
            idx = mask.nonzero().squeeze(1)
            new_target.append(target[mask] - cutoffs[i])
            h = F.linear(X.index_select(0, idx), sd[prefix + 'tail.%d.0.weight' % i])
            logits.append(F.linear(h, sd[prefix + 'tail.%d.2.weight' % i]))
        else:
            new_target.append(None)
            logits.append(None)
    return logits, new_target


def adaptive_loss(X, target, sd, cutoffs, padding_idx=1):
    """criteria/adaptive_loss.py:27-73 + transformer_faces_objects.py:82-90.
    Returns (loss_sum, ntokens, loss = sum / ln2 / ntokens).  ignore_index is applied to the
    CLUSTER-LOCAL targets, exactly as the reference does."""
    logits, tgt = adaptive_logits(X, target, sd, cutoffs)
    loss = X.new_zeros(1)
    for lg, t in zip(logits, tgt):
        if t is not None:
            loss = loss + F.cross_entropy(lg, t, ignore_index=padding_idx, reduction='sum')
    ntokens = int(target.ne(padding_idx).sum())
    return loss, ntokens, loss / math.log(2) / ntokens


def adaptive_log_prob(X, sd, cutoffs, prefix='adaptive_softmax.'):
    """softmax.py:193-222 get_log_prob: [B,T,vocab]."""
    B, T, E = X.shape
    X2 = X.contiguous().view(-1, E)
    head_w = torch.cat([sd[prefix + 'head.word_proj.weight'], sd[prefix + 'head.class_proj.weight']])
    head_lp = F.log_softmax(F.linear(X2, head_w), dim=1)
    parts = [head_lp[:, :cutoffs[0]]]
    for i in range(len(cutoffs) - 1):
        h = F.linear(X2, sd[prefix + 'tail.%d.0.weight' % i])
        t = F.log_softmax(F.linear(h, sd[prefix + 'tail.%d.2.weight' % i]), dim=1)
        parts.append(t + head_lp[:, cutoffs[0] + i, None])
    return torch.cat(parts, dim=1).view(B, T, -1)


# ------------------------------------------------------------------------------------- model level
def nan_mask_(x):
    """transformer_faces_objects.py:374-379: mask = isnan(x).any(-1); x[mask] = 0 (in place)."""
    mask = torch.isnan(x).any(dim=-1)
    x[mask] = 0
    return mask


def build_contexts(image_feats, article_hiddens, bert_weight, article_ids, face_embeds, obj_embeds,
                   padding_idx=1):
    """transformer_faces_objects.py:335-395.  image_feats [B,2048,7,7] (ResNet output),
    article_hiddens: list of 25 [B,S,1024] (RoBERTa all-layer features)."""
    B = article_ids.shape[0]
    contexts = {}
    if image_feats is not None:
        X_image = image_feats.permute(0, 2, 3, 1).reshape(B, -1, image_feats.shape[1])
        contexts['image'] = X_image.transpose(0, 1)
        contexts['image_mask'] = X_image.new_zeros(B, X_image.shape[1]).bool()
    X_article = torch.stack(article_hiddens, dim=2)                 # [B,S,25,1024]
    w = F.softmax(bert_weight, dim=0).unsqueeze(0).unsqueeze(1).unsqueeze(3)
    X_article = (X_article * w).sum(dim=2)
    contexts['article'] = X_article.transpose(0, 1)
    contexts['article_mask'] = article_ids == padding_idx
    if face_embeds is not None:
        fm = nan_mask_(face_embeds)
        contexts['faces'] = face_embeds.transpose(0, 1)
        contexts['faces_mask'] = fm
    if obj_embeds is not None:
        om = nan_mask_(obj_embeds)
        contexts['obj'] = obj_embeds.transpose(0, 1)
        contexts['obj_mask'] = om
    return contexts


def shift_caption(caption_ids, padding_idx=1):
    """transformer_faces_objects.py:321-329: inputs = ids[:, :-1], targets = ids[:, 1:]."""
    target = caption_ids[:, 1:].contiguous()
    inp = caption_ids[:, :-1].contiguous()
    return inp, target


def greedy_generate(seed_ids, contexts, sd, cfg, gen_len=100, eos=2, padding_idx=1,
                    early_exit=True):
    """transformer_faces_objects.py:399-494 _generate with sampling_topk=1 (argmax), written
    without the active-row compaction but emitting the identical matrices: token_ids [B, 1+n]
    (seed column first) and log_probs [B, n]; finished rows emit pad (1) with log-prob 0."""
    B = seed_ids.shape[0]
    state = {}
    active = torch.ones(B, dtype=torch.bool)
    prev = seed_ids[:, 0:1]
    cols, lps = [prev.clone()], []
    for _ in range(gen_len):
        X, _ = decoder_forward(prev, contexts, sd, cfg, state)
        lp = adaptive_log_prob(X[:, -1:], sd, cfg['cutoffs'])[:, 0]
        best_lp, best = lp.max(dim=-1)
        tok = torch.where(active, best, torch.full_like(best, padding_idx))
        cols.append(tok.unsqueeze(1))
        lps.append(torch.where(active, best_lp, torch.zeros_like(best_lp)).unsqueeze(1))
        active = active & (tok != eos)
        prev = tok.unsqueeze(1)
        if early_exit and not bool(active.any()):
            break
    return torch.cat(cols, dim=1), torch.cat(lps, dim=1)


# ------------------------------------------------------------------------------------- encoders
def _resnet_bn(x, sd, p, prefix, bn_mode, bn_eps, momentum, stats, storage=None):
    """nn.BatchNorm2d in eval() ('running') or train() ('batch') mode on a raw convolution output.
    storage (tests only): models where the CUDA path keeps bf16 -- the raw convolution in batch
    mode (running mode folds the scale into the weights, see _resnet_conv_bn)."""
    if bn_mode == 'running':
        return F.batch_norm(x, sd[p + 'running_mean'], sd[p + 'running_var'], sd[p + 'weight'],
                            sd[p + 'bias'], False, 0.0, bn_eps)
    if storage is not None:
        x = storage(x)
    n = x.numel() // x.shape[1]
    mean = x.mean(dim=(0, 2, 3))
    var = x.var(dim=(0, 2, 3), unbiased=False)
    name = p[len(prefix):]
    stats[name + 'running_mean'] = (1 - momentum) * sd[p + 'running_mean'] + momentum * mean
    stats[name + 'running_var'] = ((1 - momentum) * sd[p + 'running_var']
                                   + momentum * var * n / max(n - 1, 1))
    y = (x - mean.view(1, -1, 1, 1)) * torch.rsqrt(var + bn_eps).view(1, -1, 1, 1)
    return y * sd[p + 'weight'].view(1, -1, 1, 1) + sd[p + 'bias'].view(1, -1, 1, 1)


def _resnet_conv_bn(x, sd, conv, bn, prefix, bn_mode, bn_eps, momentum, stats, storage, **kw):
    w = sd[conv + 'weight']
    if storage is None:
        return _resnet_bn(F.conv2d(x, w, **kw), sd, bn, prefix, bn_mode, bn_eps, momentum, stats)
    if bn_mode == 'running':       # the CUDA path rounds the FOLDED weight, bias stays fp32
        scale = sd[bn + 'weight'] / torch.sqrt(sd[bn + 'running_var'] + bn_eps)
        bias = sd[bn + 'bias'] - sd[bn + 'running_mean'] * scale
        return F.conv2d(x, storage(w * scale.view(-1, 1, 1, 1)), **kw) + bias.view(1, -1, 1, 1)
    return _resnet_bn(F.conv2d(x, storage(w), **kw), sd, bn, prefix, bn_mode, bn_eps, momentum,
                      stats, storage)


def resnet_bottleneck(x, sd, p, stride, prefix='', bn_mode='running', bn_eps=1e-5, momentum=0.1,
                      stats=None, storage=None):
    """torchvision Bottleneck.forward (v1.5: the stride sits on the 3x3 convolution)."""
    stats = {} if stats is None else stats
    st = storage if storage is not None else (lambda t: t)
    a = (sd, None, None, prefix, bn_mode, bn_eps, momentum, stats, storage)

    def cb(x, conv, bn, **kw):
        return _resnet_conv_bn(x, sd, p + conv, p + bn, prefix, bn_mode, bn_eps, momentum, stats,
                               storage, **kw)
    idn = x
    o = st(F.relu(cb(x, 'conv1.', 'bn1.')))
    o = st(F.relu(cb(o, 'conv2.', 'bn2.', stride=stride, padding=1)))
    o = cb(o, 'conv3.', 'bn3.')
    if (p + 'downsample.0.weight') in sd:
        idn = st(cb(x, 'downsample.0.', 'downsample.1.', stride=stride))
    return st(F.relu(o + idn))


def resnet152_forward(image, sd, prefix='resnet.', bn_eps=1e-5, blocks=(3, 8, 36, 3),
                      bn_mode='running', return_stats=False, momentum=0.1, storage=None,
                      collect=None):
    """models/resnet.py:92-108 (torchvision Bottleneck stack): [B,3,224,224] -> [B,2048,7,7].
    bn_mode 'running' = eval() semantics (evaluate / demo); 'batch' = train() semantics -- what the
    reference's training step runs on the frozen backbone (callback_apex_trainer.py:259 calls
    model.train(); only the PARAMETERS are frozen, config.yaml `no_grad`): every BatchNorm normalises
    with the biased batch variance and moves its running statistics by `momentum` towards the batch
    mean / UNBIASED batch variance (returned as `stats` when return_stats).
    Pinned by tests/golden/resnet152.npz (the reference's own ResNetFeatureExtractor, both modes).
    storage: None = the reference arithmetic (fp32 throughout).  A rounding function (e.g.
    lambda t: t.bfloat16().float()) is applied wherever the CUDA path keeps a value in bf16 -- image,
    (folded) weights, raw convolutions in batch mode, every block-internal activation -- giving the
    "same algorithm, bf16 storage" model the tests use to separate storage precision from mistakes.
    collect: optional list receiving (name, input, output) of the stem and of every block."""
    stats = {}
    st = storage if storage is not None else (lambda t: t)
    x = _resnet_conv_bn(st(image), sd, prefix + 'conv1.', prefix + 'bn1.', prefix, bn_mode, bn_eps,
                        momentum, stats, storage, stride=2, padding=3)
    x = F.max_pool2d(st(F.relu(x)), 3, 2, 1)
    if collect is not None:
        collect.append(('stem', image, x))
    for li, n in enumerate(blocks):
        for bi in range(n):
            p = prefix + 'layer%d.%d.' % (li + 1, bi)
            stride = 2 if (li > 0 and bi == 0) else 1
            y = resnet_bottleneck(x, sd, p, stride, prefix, bn_mode, bn_eps, momentum, stats, storage)
            if collect is not None:
                collect.append((p[len(prefix):-1], x, y))
            x = y
    return (x, stats) if return_stats else x


def roberta_forward(ids, sd, n_layers, heads, prefix='roberta.', padding_idx=1, eps=1e-5):
    """fairseq RoBERTa encoder `extract_features(return_all_hiddens=True)` (call site
    transformer_faces_objects.py:352-353): learned positions (pad-aware, offset by pad+1),
    embedding LayerNorm, post-LN encoder layers with GELU; returns n_layers+1 [B,S,E] tensors.
    Un-vendored third-party model (fairseq @ 2f7e3f3323): parity unpinned; key names follow
    fairseq's `decoder.sentence_encoder.*` layout."""
    p = prefix + 'decoder.sentence_encoder.'
    B, S = ids.shape
    pad = ids.eq(padding_idx)
    pos = make_positions(ids, padding_idx, False)
    x = F.embedding(ids, sd[p + 'embed_tokens.weight']) + F.embedding(pos, sd[p + 'embed_positions.weight'])
    x = F.layer_norm(x, (x.shape[-1],), sd[p + 'emb_layer_norm.weight'], sd[p + 'emb_layer_norm.bias'], eps)
    x = x * (~pad).unsqueeze(-1).type_as(x)
    E = x.shape[-1]
    d = E // heads
    hiddens = [x]
    for i in range(n_layers):
        lp = p + 'layers.%d.' % i
        qkv = F.linear(x, sd[lp + 'self_attn.in_proj_weight'], sd[lp + 'self_attn.in_proj_bias'])
        q, k, v = qkv.chunk(3, dim=-1)
        q = (q * d ** -0.5).view(B, S, heads, d).transpose(1, 2)
        k = k.view(B, S, heads, d).transpose(1, 2)
        v = v.view(B, S, heads, d).transpose(1, 2)
        w = torch.matmul(q, k.transpose(-1, -2))
        w = w.masked_fill(pad.view(B, 1, 1, S), float('-inf'))
        w = F.softmax(w, dim=-1)
        a = torch.matmul(w, v).transpose(1, 2).reshape(B, S, E)
        a = F.linear(a, sd[lp + 'self_attn.out_proj.weight'], sd[lp + 'self_attn.out_proj.bias'])
        x = F.layer_norm(x + a, (E,), sd[lp + 'self_attn_layer_norm.weight'],
                         sd[lp + 'self_attn_layer_norm.bias'], eps)
        h = F.gelu(F.linear(x, sd[lp + 'fc1.weight'], sd[lp + 'fc1.bias']))
        h = F.linear(h, sd[lp + 'fc2.weight'], sd[lp + 'fc2.bias'])
        x = F.layer_norm(x + h, (E,), sd[lp + 'final_layer_norm.weight'],
                         sd[lp + 'final_layer_norm.bias'], eps)
        hiddens.append(x)
    return hiddens


# ------------------------------------------------------------------------------------- optimizer
def bert_adam_schedule(step, t_total, warmup, schedule='warmup_linear'):
    """pytorch-pretrained-bert 0.6.2 `_LRSchedule.get_lr` + `WarmupLinearSchedule.get_lr_` /
    `WarmupConstantSchedule.get_lr_` (the dependency behind `type: bert_adam`,
    expt/nytimes/9_transformer_objects/config.yaml:126-132; un-vendored, pulled in by allennlp 0.9).
    Plain Python floats, as there.  Parity unpinned: the package is absent from /root/reference and
    the reference has no test for it; this follows the published source."""
    if t_total < 0 or schedule in (None, 'none'):
        return 1.0
    warmup = max(warmup, 0.0)
    progress = float(step) / t_total
    if progress < warmup:
        return progress / warmup
    if schedule == 'warmup_constant':
        return 1.0
    return max((progress - 1.0) / (warmup - 1.0), 0.0)


def bert_adam_step(params, grads, state, lr, warmup=-1, t_total=-1, schedule='warmup_linear',
                   b1=0.9, b2=0.999, e=1e-6, weight_decay=0.01, max_grad_norm=1.0):
    """One `BertAdam.step()` (pytorch-pretrained-bert 0.6.2 optimization.py, the optimizer stepped at
    tell/training/callback_apex_trainer.py:238) over lists of fp32 tensors, in place.
    state: list of dicts {'step', 'next_m', 'next_v'} (created on first use)."""
    for p, g, st in zip(params, grads, state):
        if not st:
            st.update(step=0, next_m=torch.zeros_like(p), next_v=torch.zeros_like(p))
        g = g.clone()
        if max_grad_norm > 0:                       # clip_grad_norm_(p, max_grad_norm): per tensor
            total_norm = g.norm(2)
            clip_coef = max_grad_norm / (total_norm + 1e-6)
            if clip_coef < 1:
                g.mul_(clip_coef)
        st['next_m'].mul_(b1).add_(g, alpha=1 - b1)
        st['next_v'].mul_(b2).addcmul_(g, g, value=1 - b2)
        update = st['next_m'] / (st['next_v'].sqrt() + e)
        if weight_decay > 0.0:
            update = update + weight_decay * p
        lr_scheduled = lr * bert_adam_schedule(st['step'], t_total, warmup, schedule)
        p.add_(-(lr_scheduled * update))
        st['step'] += 1


# ------------------------------------------------------------------------------------- face encoders
def _basic_conv(x, sd, p, stride=1, padding=0):
    """BasicConv2d (tell/facenet/inception_resnet_v1.py:10-34): conv (no bias), BatchNorm eps 1e-3
    in eval mode, ReLU."""
    x = F.conv2d(x, sd[p + 'conv.weight'], None, stride, padding)
    x = F.batch_norm(x, sd[p + 'bn.running_mean'], sd[p + 'bn.running_var'], sd[p + 'bn.weight'],
                     sd[p + 'bn.bias'], False, 0.0, 1e-3)
    return F.relu(x)


def _res_block(x, sd, p, tails, scale, relu=True):
    """Block35 / Block17 / Block8 (:37-119): branches -> cat -> 1x1 conv (bias) * scale + x -> ReLU."""
    outs = [_basic_conv(x, sd, p + 'branch0.')]
    for i, tail in enumerate(tails, start=1):
        h = _basic_conv(x, sd, p + 'branch%d.0.' % i)
        for j, pad in enumerate(tail, start=1):
            h = _basic_conv(h, sd, p + 'branch%d.%d.' % (i, j), 1, pad)
        outs.append(h)
    out = F.conv2d(torch.cat(outs, 1), sd[p + 'conv2d.weight'], sd[p + 'conv2d.bias'])
    out = out * scale + x
    return F.relu(out) if relu else out


def inception_resnet_v1_forward(x, sd):
    """InceptionResnetV1.forward in eval mode (tell/facenet/inception_resnet_v1.py:264-299):
    x [B,3,H,W] -> (l2-normalised embedding [B,512], logits)."""
    x = _basic_conv(x, sd, 'conv2d_1a.', 2)
    x = _basic_conv(x, sd, 'conv2d_2a.')
    x = _basic_conv(x, sd, 'conv2d_2b.', 1, 1)
    x = F.max_pool2d(x, 3, 2)
    x = _basic_conv(x, sd, 'conv2d_3b.')
    x = _basic_conv(x, sd, 'conv2d_4a.')
    x = _basic_conv(x, sd, 'conv2d_4b.', 2)
    for i in range(5):
        x = _res_block(x, sd, 'repeat_1.%d.' % i, [[1], [1, 1]], 0.17)
    x = torch.cat([_basic_conv(x, sd, 'mixed_6a.branch0.', 2),                     # Mixed_6a :122-144
                   _basic_conv(_basic_conv(_basic_conv(x, sd, 'mixed_6a.branch1.0.'), sd,
                                           'mixed_6a.branch1.1.', 1, 1), sd, 'mixed_6a.branch1.2.', 2),
                   F.max_pool2d(x, 3, 2)], 1)
    for i in range(10):
        x = _res_block(x, sd, 'repeat_2.%d.' % i, [[(0, 3), (3, 0)]], 0.10)
    b0 = _basic_conv(_basic_conv(x, sd, 'mixed_7a.branch0.0.'), sd, 'mixed_7a.branch0.1.', 2)   # :147-181
    b1 = _basic_conv(_basic_conv(x, sd, 'mixed_7a.branch1.0.'), sd, 'mixed_7a.branch1.1.', 2)
    b2 = _basic_conv(_basic_conv(_basic_conv(x, sd, 'mixed_7a.branch2.0.'), sd, 'mixed_7a.branch2.1.', 1, 1),
                     sd, 'mixed_7a.branch2.2.', 2)
    x = torch.cat([b0, b1, b2, F.max_pool2d(x, 3, 2)], 1)
    for i in range(5):
        x = _res_block(x, sd, 'repeat_3.%d.' % i, [[(0, 1), (1, 0)]], 0.20)
    x = _res_block(x, sd, 'block8.', [[(0, 1), (1, 0)]], 1.0, relu=False)
    x = x.mean(dim=(2, 3))
    x = F.linear(x, sd['last_linear.weight'])
    x = F.batch_norm(x, sd['last_bn.running_mean'], sd['last_bn.running_var'], sd['last_bn.weight'],
                     sd['last_bn.bias'], False, 0.0, 1e-3)
    x = F.normalize(x, p=2, dim=1)
    return x, F.linear(x, sd['logits.weight'], sd['logits.bias'])


def _conv_prelu(x, sd, i):
    return F.prelu(F.conv2d(x, sd['conv%d.weight' % i], sd['conv%d.bias' % i]), sd['prelu%d.weight' % i])


def pnet_forward(x, sd):
    """PNet.forward (tell/facenet/mtcnn.py:40-51) -> (box regression, face probability)."""
    x = F.max_pool2d(_conv_prelu(x, sd, 1), 2, 2, ceil_mode=True)
    x = _conv_prelu(_conv_prelu(x, sd, 2), sd, 3)
    a = F.softmax(F.conv2d(x, sd['conv4_1.weight'], sd['conv4_1.bias']), dim=1)
    return F.conv2d(x, sd['conv4_2.weight'], sd['conv4_2.bias']), a


def rnet_forward(x, sd):
    """RNet.forward (mtcnn.py:86-101); note the (w, h, c) flatten before dense4."""
    x = F.max_pool2d(_conv_prelu(x, sd, 1), 3, 2, ceil_mode=True)
    x = F.max_pool2d(_conv_prelu(x, sd, 2), 3, 2, ceil_mode=True)
    x = _conv_prelu(x, sd, 3).permute(0, 3, 2, 1).contiguous()
    x = F.prelu(F.linear(x.view(x.shape[0], -1), sd['dense4.weight'], sd['dense4.bias']), sd['prelu4.weight'])
    a = F.softmax(F.linear(x, sd['dense5_1.weight'], sd['dense5_1.bias']), dim=1)
    return F.linear(x, sd['dense5_2.weight'], sd['dense5_2.bias']), a


def onet_forward(x, sd):
    """ONet.forward (mtcnn.py:136-159) -> (box, landmarks, probability)."""
    x = F.max_pool2d(_conv_prelu(x, sd, 1), 3, 2, ceil_mode=True)
    x = F.max_pool2d(_conv_prelu(x, sd, 2), 3, 2, ceil_mode=True)
    x = F.max_pool2d(_conv_prelu(x, sd, 3), 2, 2, ceil_mode=True)
    x = _conv_prelu(x, sd, 4).permute(0, 3, 2, 1).contiguous()
    x = F.prelu(F.linear(x.view(x.shape[0], -1), sd['dense5.weight'], sd['dense5.bias']), sd['prelu5.weight'])
    a = F.softmax(F.linear(x, sd['dense6_1.weight'], sd['dense6_1.bias']), dim=1)
    return (F.linear(x, sd['dense6_2.weight'], sd['dense6_2.bias']),
            F.linear(x, sd['dense6_3.weight'], sd['dense6_3.bias']), a)
