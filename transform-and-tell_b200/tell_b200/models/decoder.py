"""The DynamicConv decoder family of tell/models/decoder_*.py on the B200 kernels.

One N-context layer covers the four reference variants, which differ only in their context list:
    dynamic_conv_decoder_faces_objects      image, article, faces, obj   (decoder_faces_objects.py)
    dynamic_conv_decoder_faces_parallel     image, article, faces        (decoder_faces_parallel.py)
    dynamic_conv_decoder_flattened          image, article               (decoder_flattened.py)
    dynamic_conv_decoder_flattened_no_image article                      (decoder_flattened_no_image.py)
State-dict keys and constructor kwargs are the reference's.
"""
import torch
import torch.nn as nn

from .. import config
from .. import functional as Fn
from .. import ops
from .. import twin
from .. import weight_bank
from ..modules import (AdaptiveEmbedding, AdaptiveSoftmax, DynamicConv1dTBC, GehringLinear,
                       LayerNorm, LightweightConv1dTBC, MultiHeadAttention, TextFieldEmbedder)
from ..registry import Registrable
from ..utils import eval_str_list


class Decoder(nn.Module, Registrable):
    pass


class DecoderLayer(nn.Module, Registrable):
    pass


_KV_KEY = 'tell_b200.kv_cache'   # our own incremental-state entry (projected contexts)


class DynamicConvDecoderLayer(DecoderLayer):
    """decoder_faces_objects.py:183-379 generalised over the context list.
    context_dims: ordered {name: kdim}."""

    def __init__(self, decoder_embed_dim, decoder_conv_dim, decoder_glu, decoder_conv_type,
                 weight_softmax, decoder_attention_heads, weight_dropout, dropout, relu_dropout,
                 input_dropout, decoder_normalize_before, attention_dropout, decoder_ffn_embed_dim,
                 context_dims, kernel_size=0):
        super().__init__()
        if decoder_normalize_before:
            raise NotImplementedError('decoder_normalize_before=True is not used by the shipped configs')
        self.embed_dim = decoder_embed_dim
        self.conv_dim = decoder_conv_dim
        self.glu = bool(decoder_glu)
        self.linear1 = GehringLinear(self.embed_dim, (2 if self.glu else 1) * self.conv_dim)
        if decoder_conv_type == 'lightweight':
            self.conv = LightweightConv1dTBC(self.conv_dim, kernel_size, padding_l=kernel_size - 1,
                                             weight_softmax=weight_softmax,
                                             num_heads=decoder_attention_heads,
                                             weight_dropout=weight_dropout)
        elif decoder_conv_type == 'dynamic':
            self.conv = DynamicConv1dTBC(self.conv_dim, kernel_size, padding_l=kernel_size - 1,
                                         weight_softmax=weight_softmax,
                                         num_heads=decoder_attention_heads,
                                         weight_dropout=weight_dropout)
        else:
            raise NotImplementedError
        self.linear2 = GehringLinear(self.conv_dim, self.embed_dim)
        self.dropout = dropout
        self.relu_dropout = relu_dropout
        self.input_dropout = input_dropout
        self.normalize_before = decoder_normalize_before
        self.conv_layer_norm = LayerNorm(self.embed_dim)
        self.context_names = list(context_dims.keys())
        self.context_attns = nn.ModuleDict()
        self.context_attn_lns = nn.ModuleDict()
        for name, kdim in context_dims.items():
            self.context_attns[name] = MultiHeadAttention(
                self.embed_dim, decoder_attention_heads, kdim=kdim, vdim=kdim,
                dropout=attention_dropout)
            self.context_attn_lns[name] = LayerNorm(self.embed_dim)
        self.context_fc = GehringLinear(self.embed_dim * len(context_dims), self.embed_dim)
        self.fc1 = GehringLinear(self.embed_dim, decoder_ffn_embed_dim)
        self.fc2 = GehringLinear(decoder_ffn_embed_dim, self.embed_dim)
        self.final_layer_norm = LayerNorm(self.embed_dim)
        self.need_attn = True

    def _p(self, p):
        return p if self.training else 0.0

    @staticmethod
    def _seed(p):
        return config.next_seed() if p > 0 else 0

    def forward(self, X, contexts, incremental_state, kv_cache=None):
        """X [T,B,E] -> ([T,B,E], {ctx: head-averaged attention [B,T,S+2]} in eval mode)."""
        T, B, E = X.shape
        N = T * B
        X2 = X.reshape(N, E)
        # ---- conv block (decoder_faces_objects.py:256-266)
        p_in = self._p(self.input_dropout)
        h = Fn.DropoutFn.apply(X2, p_in, self._seed(p_in), True) if p_in > 0 else X2
        h = self.linear1(h)
        if self.glu:
            h = Fn.GLUFn.apply(h)
        h = self.conv(h.view(T, B, self.conv_dim), incremental_state=incremental_state)
        h = self.linear2(h.reshape(N, self.conv_dim))
        p = self._p(self.dropout)
        X2 = Fn.ResidualLayerNormFn.apply(h, X2.contiguous(), self.conv_layer_norm.weight,
                                          self.conv_layer_norm.bias, p, self._seed(p),
                                          self.conv_layer_norm.eps)
        # ---- parallel cross-attention branches (:272-352), all reading the same X2:
        #      one fused Q GEMM, n attention launches on column blocks, n out_proj GEMMs
        need_w = (not self.training) and self.need_attn
        names = self.context_names
        n = len(names)
        mhas = [self.context_attns[nm] for nm in names]
        kvs, masks = [], []
        for nm, mha in zip(names, mhas):
            if kv_cache is not None and nm in kv_cache:
                kv = kv_cache[nm]
            else:
                kv = mha.project_kv(contexts[nm], need_weights=(not self.training) and self.need_attn)
                if kv_cache is not None:
                    kv_cache[nm] = kv
            kvs.append(kv)
            m = contexts.get(nm + '_mask/u8')      # converted once per forward by the decoder
            if m is None and contexts[nm + '_mask'] is not None:
                m = contexts[nm + '_mask'].to(torch.uint8).contiguous()
            masks.append(m if kv is not None else None)
        E_ = self.embed_dim
        q_ws = [mha._weights()[0] for mha in mhas]
        q_bs = [mha._bias_q() for mha in mhas]
        Q_all = Fn.FusedQProjFn.apply(X2, mhas[0].scaling, n, *q_ws, *q_bs)
        p_att = self._p(mhas[0].dropout)
        seeds_a = tuple(self._seed(p_att) for _ in range(n))
        slabs = tuple(kv_cache.get(nm + '/slab') if kv_cache is not None else None for nm in names)
        lens = tuple(contexts.get(nm + '_mask/len') for nm in names)     # valid key counts (long contexts)
        extra = (slabs, lens) if any(s_ is not None for s_ in slabs + lens) else ()
        hm = [kv_cache.get(nm + '/hm') if kv_cache is not None else None for nm in names]
        if T == 1 and not need_w and not torch.is_grad_enabled() and \
                all(h_ is not None or kv is None for h_, kv in zip(hm, kvs)) and any(h_ is not None for h_ in hm):
            # incremental decoding over the head-major K|V cache (built once by the generate loop)
            A_all = torch.empty_like(Q_all)
            tw = Fn._tw()
            A16 = torch.empty(Q_all.shape, dtype=torch.bfloat16, device=Q_all.device) if tw else None
            if config.attn_multi and n <= 4 and mhas[0].head_dim == 64:
                # all contexts of the layer in ONE launch (blockIdx.z walks them)
                items = []
                for c, (mha, kv) in enumerate(zip(mhas, kvs)):
                    sl = slice(c * E_, (c + 1) * E_)
                    S_c = hm[c][0].shape[2] if kv is not None else 0
                    items.append(dict(q=Q_all[:, sl], k=hm[c][0] if kv is not None else None,
                                      v=hm[c][1] if kv is not None else None,
                                      bias_k=mha.bias_k.view(-1) if mha.bias_k is not None else None,
                                      bias_v=mha.bias_v.view(-1) if mha.bias_v is not None else None,
                                      mask=masks[c], out=A_all[:, sl], lse=None, S=S_c,
                                      kv_len=kv_cache.get(names[c] + '/len'),
                                      out16=A16[:, sl] if tw else None))
                ops.attn_decode_hm_multi(items, B, mhas[0].num_heads, mhas[0].head_dim, mhas[0].add_zero_attn)
                if tw:
                    twin.put(A_all, A16)
            else:
                for c, (mha, kv) in enumerate(zip(mhas, kvs)):
                    sl = slice(c * E_, (c + 1) * E_)
                    bk = mha.bias_k.view(-1) if mha.bias_k is not None else None
                    bv = mha.bias_v.view(-1) if mha.bias_v is not None else None
                    if kv is None:        # empty context: only the bias / zero rows
                        ops.attn_fwd(Q_all[:, sl], None, None, bk, bv, None, 1, B, 0, mha.num_heads,
                                     mha.head_dim, mha.add_zero_attn, tc=True, out=A_all[:, sl])
                    else:
                        ops.attn_decode_hm(Q_all[:, sl], hm[c][0], hm[c][1], bk, bv, masks[c], A_all[:, sl],
                                           mha.add_zero_attn)
            attns = {}
        else:
            res = Fn.MultiCtxAttentionFn.apply(Q_all, T, B, mhas[0].num_heads, mhas[0].add_zero_attn,
                                               p_att, seeds_a, need_w, n, *kvs,
                                               *[m.bias_k for m in mhas], *[m.bias_v for m in mhas],
                                               *masks, *extra)
            A_all = res[0]
            attns = {nm: w for nm, w in zip(names, res[1:])} if need_w else {}
        lns = [self.context_attn_lns[nm] for nm in names]
        seeds = tuple(self._seed(p) for _ in range(n))
        out_ws = [m.out_proj.weight for m in mhas]
        if Fn.OutProjContextLNFn.usable(A_all, n, out_ws):
            # out-projections (one batched GEMM) + the n residual LayerNorms (one launch) as one node
            Xc = Fn.OutProjContextLNFn.apply(A_all, X2, p, seeds, lns[0].eps, n, *out_ws,
                                             *[m.out_proj.bias for m in mhas],
                                             *[l.weight for l in lns], *[l.bias for l in lns])
        else:
            hs = Fn.FusedOutProjFn.apply(A_all, n, *out_ws, *[m.out_proj.bias for m in mhas])
            Xc = Fn.ContextLayerNormFn.apply(X2, p, seeds, lns[0].eps, n, *hs,
                                             *[l.weight for l in lns], *[l.bias for l in lns])
        # ---- context_fc + FFN (:354-364)
        X2 = self.context_fc(Xc, twin_out=True)            # fc1 reads it as a GEMM operand
        p_relu = self._p(self.relu_dropout)
        h = self.fc1(X2, act=ops.ACT_RELU, twin_out=p_relu == 0)
        if p_relu > 0:
            h = Fn.DropoutFn.apply(h, p_relu, self._seed(p_relu), True)
        h = self.fc2(h)
        X2 = Fn.ResidualLayerNormFn.apply(h, X2, self.final_layer_norm.weight,
                                          self.final_layer_norm.bias, p, self._seed(p),
                                          self.final_layer_norm.eps)
        return X2.view(T, B, E), attns

    def make_generation_fast_(self, need_attn=False, **kwargs):
        self.need_attn = need_attn


class _DynamicConvDecoderBase(Decoder):
    CONTEXTS = None  # ordered {name: kdim}, set by subclasses

    def __init__(self, vocab, embedder: TextFieldEmbedder, max_target_positions, dropout,
                 share_decoder_input_output_embed, decoder_output_dim, decoder_conv_dim,
                 decoder_glu, decoder_conv_type, weight_softmax, decoder_attention_heads,
                 weight_dropout, relu_dropout, input_dropout, decoder_normalize_before,
                 attention_dropout, decoder_ffn_embed_dim, decoder_kernel_size_list,
                 adaptive_softmax_cutoff=None, tie_adaptive_weights=False,
                 adaptive_softmax_dropout=0, tie_adaptive_proj=False, adaptive_softmax_factor=0,
                 decoder_layers=6, final_norm=True, padding_idx=0, namespace='target_tokens',
                 vocab_size=None, section_attn=False, swap=False, article_embed_size=1024):
        super().__init__()
        self.vocab = vocab
        vocab_size = vocab_size or vocab.get_vocab_size(namespace)
        self.dropout = dropout
        self.share_input_output_embed = share_decoder_input_output_embed
        embed_dim = embedder.get_output_dim()
        self.embed_dim = embed_dim
        self.max_target_positions = max_target_positions
        self.embedder = embedder
        context_dims = dict(self.CONTEXTS)
        if 'article' in context_dims:
            context_dims['article'] = article_embed_size
        self.layers = nn.ModuleList([
            DynamicConvDecoderLayer(embed_dim, decoder_conv_dim, decoder_glu, decoder_conv_type,
                                    weight_softmax, decoder_attention_heads, weight_dropout,
                                    dropout, relu_dropout, input_dropout, decoder_normalize_before,
                                    attention_dropout, decoder_ffn_embed_dim, context_dims,
                                    kernel_size=decoder_kernel_size_list[i])
            for i in range(decoder_layers)])
        self.adaptive_softmax = None
        if adaptive_softmax_cutoff is None:
            raise NotImplementedError('the shipped configs always use the adaptive softmax')
        adaptive_inputs = None
        if isinstance(embedder, AdaptiveEmbedding):
            adaptive_inputs = embedder
        elif hasattr(embedder, 'token_embedder_adaptive'):
            adaptive_inputs = embedder.token_embedder_adaptive
        elif tie_adaptive_weights:
            raise ValueError('Cannot locate adaptive_inputs.')
        self.adaptive_softmax = AdaptiveSoftmax(
            vocab_size, embed_dim, eval_str_list(adaptive_softmax_cutoff, type=int),
            dropout=adaptive_softmax_dropout, adaptive_inputs=adaptive_inputs,
            factor=adaptive_softmax_factor, tie_proj=tie_adaptive_proj)
        self.register_buffer('version', torch.Tensor([2]))
        self.normalize = decoder_normalize_before and final_norm

    @classmethod
    def _from_params(cls, params, **extras):
        params = dict(params)
        embedder = TextFieldEmbedder.from_params(params.pop('embedder'))
        import inspect
        sig = inspect.signature(cls.__init__)
        unknown = [k for k in params if k not in sig.parameters]
        if unknown:
            from ..registry import ConfigurationError
            raise ConfigurationError('Extra parameters passed to %s: %s' % (cls.__name__, unknown))
        return cls(vocab=extras.get('vocab'), embedder=embedder, **params)

    # ------------------------------------------------------------------ weight bank
    use_weight_bank = True

    def weight_scope(self, refresh=True):
        """Context manager inside which every weight operand of this decoder comes from ONE
        tt_weight_prep launch (weight_bank.py).  refresh=False reuses the operands prepared by an
        earlier scope (incremental decoding: the weights do not change between steps).  Without a
        usable bank (parity mode, CPU-less construction, disabled) it is a no-op scope."""
        bank = None
        if self.use_weight_bank and config.precision == 'bf16' \
                and next(self.parameters()).is_cuda:
            bank = getattr(self, '_bank', None)
            if bank is None or bank._moved():
                bank = weight_bank.WeightBank(self)
                object.__setattr__(self, '_bank', bank)      # not a submodule / not in state_dict
                refresh = True
        scope = weight_bank.activate(bank)
        if bank is not None:
            if refresh or not bank.prepared_once or bank.log:
                bank.current = bank.effective_weights()
                bank.prepared_once = True
            else:
                bank.current = {id(e.v): e.w32 for e in bank.wn}
        return scope

    def forward_tbc(self, prev_target, contexts, incremental_state=None, use_layers=None):
        """Hot-path entry: returns X in decoder layout [T,B,E] (no final transpose) + extras."""
        if weight_bank.ACTIVE is None or weight_bank.ACTIVE.module is not self:
            first = incremental_state is None or _KV_KEY not in incremental_state
            with self.weight_scope(refresh=first):
                return self._forward_tbc(prev_target, contexts, incremental_state, use_layers)
        return self._forward_tbc(prev_target, contexts, incremental_state, use_layers)

    batch_kv_layers = True

    def _project_contexts_all_layers(self, contexts, caches):
        """Key|value projections of every context for ALL layers up front: one GEMM per context
        (N = L*2E) instead of one per (layer, context); see Fn.AllLayerKVProjFn.  Fills the
        per-layer caches the layers read their projected contexts from."""
        L = len(self.layers)
        for nm in self.layers[0].context_names:
            if nm in caches[0]:
                continue
            key = contexts[nm]
            if key is None or key.shape[2] == 0 or key.shape[0] == 0:
                continue                      # empty context: the per-layer path returns None
            S, B, kd = key.shape
            mhas = [layer.context_attns[nm] for layer in self.layers]
            ws = [m._weights() for m in mhas]
            E = mhas[0].embed_dim
            biases = [m._bias_kv() for m in mhas]
            need_w = (not self.training) and any(layer.need_attn for layer in self.layers)
            kv16 = Fn.kv16_ok(mhas[0].head_dim, need_w)
            slab = Fn.GradSlab(S * B, 2 * E, L, key.device,
                               torch.bfloat16 if kv16 else torch.float32) \
                if torch.is_grad_enabled() else None
            kvs = Fn.AllLayerKVProjFn.apply(key.reshape(S * B, kd), L, slab, kv16,
                                            *[w[1] for w in ws], *[w[2] for w in ws], *biases)
            for l in range(L):
                caches[l][nm] = kvs[l]
                if slab is not None:
                    caches[l][nm + '/slab'] = (slab, l)

    def build_decode_cache(self, incremental_state, contexts=None):
        """After the first incremental step: repack every projected context of every layer from the
        token-major [S*B, 2E] bf16 views into head-major K, V [B,H,S,64] (tt_kv_repack_heads).  Returns
        the number of caches built (0 when keys|values are not bf16 / head_dim is not 64)."""
        caches = incremental_state.get(_KV_KEY) if incremental_state is not None else None
        if not caches:
            return 0
        built = 0
        # per context: how many leading keys of each sample can be unmasked at all (padding is
        # trailing): the decode kernel never reads the cache rows beyond
        lens = {}
        if contexts is not None:
            for nm in self.layers[0].context_names:
                m = contexts.get(nm + '_mask')
                if m is not None and m.dim() == 2 and m.shape[1] > 0:
                    pos = torch.arange(1, m.shape[1] + 1, device=m.device, dtype=torch.int32)
                    lens[nm] = ((~m.bool()).to(torch.int32) * pos).amax(dim=1).to(torch.int32).contiguous()
        for layer, cache in zip(self.layers, caches):
            for nm in layer.context_names:
                kv = cache.get(nm)
                mha = layer.context_attns[nm]
                if kv is None or kv.dtype != torch.bfloat16 or mha.head_dim != 64 or nm + '/hm' in cache:
                    continue
                E = mha.embed_dim
                B = cache['/B']
                S = kv.shape[0] // B
                if (S + 2) * 8 * 4 > 40 * 1024:          # key set beyond the decode kernel's budget
                    continue
                cache[nm + '/hm'] = ops.kv_repack_heads(kv[:, :E], kv[:, E:], S, B, mha.num_heads, 64)
                if nm in lens and lens[nm].shape[0] == B:
                    cache[nm + '/len'] = lens[nm]
                built += 1
        return built

    def _forward_tbc(self, prev_target, contexts, incremental_state=None, use_layers=None):
        twin.clear()                         # operand twins never outlive one forward + backward
        for layer in self.layers:            # parameter splits belong to one forward's autograd graph
            for mha in layer.context_attns.values():
                mha.begin_step()
        X2, ids = self.embedder.embed_tbc(prev_target, incremental_state)
        B, T = ids.shape
        p = self.dropout if self.training else 0.0
        if p > 0:
            X2 = Fn.DropoutFn.apply(X2, p, config.next_seed())
        X = X2.view(T, B, self.embed_dim)
        attns, inner_states = [], [X]
        if incremental_state is not None:
            caches = incremental_state.setdefault(_KV_KEY, [dict() for _ in self.layers])
        else:
            caches = [dict() for _ in self.layers]
        for c_ in caches:
            c_['/B'] = B
        contexts = dict(contexts)            # key-padding masks as uint8, once for all layers
        for nm in self.layers[0].context_names:
            m = contexts.get(nm + '_mask')
            if m is not None and nm + '_mask/u8' not in contexts:
                contexts[nm + '_mask/u8'] = m.to(torch.uint8).contiguous()
            # long contexts (the article): per-sample count of leading keys that can be unmasked, so
            # that the attention kernels skip the key tiles of trailing padding (full forward only;
            # the incremental step takes it from the decode cache)
            if config.attn_skip_padding and m is not None and m.dim() == 2 and m.shape[1] >= 128 \
                    and X.shape[0] > 1 and nm + '_mask/len' not in contexts:
                pos = torch.arange(1, m.shape[1] + 1, device=m.device, dtype=torch.int32)
                contexts[nm + '_mask/len'] = ((~m.bool()).to(torch.int32) * pos).amax(dim=1) \
                    .to(torch.int32).contiguous()
        if self.batch_kv_layers and not use_layers:
            self._project_contexts_all_layers(contexts, caches)
        for i, layer in enumerate(self.layers):
            attn = None
            if not use_layers or i in use_layers:
                X, attn = layer(X, contexts, incremental_state,
                                caches[i] if caches is not None else None)
                inner_states.append(X)
            attns.append(attn)
        return X, {'attn': attns, 'inner_states': inner_states}

    def forward(self, prev_target, contexts, incremental_state=None, use_layers=None, **kwargs):
        """decoder_faces_objects.py:95-142: -> (X [B,T,E], {'attn', 'inner_states'})."""
        X, extra = self.forward_tbc(prev_target, contexts, incremental_state, use_layers)
        return Fn.Transpose01Fn.apply(X), extra

    def max_positions(self):
        return self.max_target_positions

    def get_normalized_probs(self, net_output, log_probs, sample=None):
        out = self.adaptive_softmax.get_log_prob(net_output[0], None)
        return out if log_probs else out.exp()

    def filter_incremental_state(self, incremental_state, active_idx):
        """decoder_faces_objects.py:175-180 (+ our projected-context cache, same batch axis)."""
        if incremental_state is None:
            return
        for key in incremental_state:
            if 'DynamicConv1dTBC' in key:
                incremental_state[key] = incremental_state[key][:, active_idx]
        if _KV_KEY in incremental_state:
            raise NotImplementedError('row compaction with a projected-context cache: use '
                                      'the masked (non-compacting) greedy loop of the model')


@Decoder.register('dynamic_conv_decoder_faces_objects')
class DynamicConvFacesObjectsDecoder(_DynamicConvDecoderBase):
    CONTEXTS = (('image', 2048), ('article', 1024), ('faces', 512), ('obj', 2048))


@Decoder.register('dynamic_conv_decoder_faces_parallel')
class DynamicConvFacesParallelDecoder(_DynamicConvDecoderBase):
    CONTEXTS = (('image', 2048), ('article', 1024), ('faces', 512))


@Decoder.register('dynamic_conv_decoder_flattened')
class DynamicConvFlattenedDecoder(_DynamicConvDecoderBase):
    CONTEXTS = (('image', 2048), ('article', 1024))


@Decoder.register('dynamic_conv_decoder_flattened_no_image')
class DynamicConvDecoderNoImage(_DynamicConvDecoderBase):
    CONTEXTS = (('article', 1024),)
