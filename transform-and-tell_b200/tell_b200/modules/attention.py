"""tell/modules/attention/multi_head.py:207-552 (static_kv cross-attention mode) on B200 kernels."""
import torch
import torch.nn as nn

from .. import config
from .. import functional as Fn
from .linear import Linear, linear


class MultiHeadAttention(nn.Module):
    """Same constructor, parameters and state-dict keys as the reference module.  forward()
    implements the mode every decoder layer uses: incremental_state=None, attn_mask=None,
    key is value (decoder_faces_objects.py:275-282)."""

    def __init__(self, embed_dim, num_heads, kdim=None, vdim=None, dropout=0., bias=True,
                 add_bias_kv=True, add_zero_attn=True, self_attention=False,
                 encoder_decoder_attention=False, out_dim=None):
        super().__init__()
        self.embed_dim = embed_dim
        self.kdim = kdim if kdim is not None else embed_dim
        self.vdim = vdim if vdim is not None else embed_dim
        self.qkv_same_dim = self.kdim == embed_dim and self.vdim == embed_dim
        self.num_heads = num_heads
        self.dropout = dropout
        self.head_dim = embed_dim // num_heads
        assert self.head_dim * num_heads == embed_dim, 'embed_dim must be divisible by num_heads'
        self.scaling = self.head_dim ** -0.5
        self.self_attention = self_attention
        self.encoder_decoder_attention = encoder_decoder_attention
        if self.qkv_same_dim:
            self.in_proj_weight = nn.Parameter(torch.empty(3 * embed_dim, embed_dim))
            nn.init.xavier_uniform_(self.in_proj_weight)
        else:
            self.k_proj_weight = nn.Parameter(torch.empty(embed_dim, self.kdim))
            self.v_proj_weight = nn.Parameter(torch.empty(embed_dim, self.vdim))
            self.q_proj_weight = nn.Parameter(torch.empty(embed_dim, embed_dim))
            for w in (self.k_proj_weight, self.v_proj_weight, self.q_proj_weight):
                nn.init.xavier_uniform_(w)
        self.in_proj_bias = nn.Parameter(torch.zeros(3 * embed_dim)) if bias else None
        self.out_proj = Linear(embed_dim, out_dim if out_dim else embed_dim, bias=bias)
        if add_bias_kv:
            self.bias_k = nn.Parameter(torch.empty(1, 1, embed_dim))
            self.bias_v = nn.Parameter(torch.empty(1, 1, embed_dim))
            nn.init.xavier_normal_(self.bias_k)
            nn.init.xavier_normal_(self.bias_v)
        else:
            self.bias_k = self.bias_v = None
        self.add_zero_attn = add_zero_attn

    # -- projections (multi_head.py:491-518)
    def begin_step(self):
        """Forget the parameter split of the previous forward (it belongs to that autograd graph)."""
        self._split = None

    def _splits(self):
        """(Wq, Wk, Wv, bq, bkv).  With autograd on, the slices of in_proj_weight / in_proj_bias come
        from ONE InProjSplitFn node per forward (see functional.InProjSplitFn)."""
        E = self.embed_dim
        w = self.in_proj_weight if self.qkv_same_dim else None
        b = self.in_proj_bias
        need = torch.is_grad_enabled() and ((w is not None and w.requires_grad) or
                                            (b is not None and b.requires_grad))
        if not need:
            ws = (w[:E], w[E:2 * E], w[2 * E:]) if w is not None else \
                (self.q_proj_weight, self.k_proj_weight, self.v_proj_weight)
            return ws + ((b[:E], b[E:]) if b is not None else (None, None))
        sp = getattr(self, '_split', None)
        if sp is None:
            outs = list(Fn.InProjSplitFn.apply(w, b, E))
            ws = tuple(outs[:3]) if w is not None else \
                (self.q_proj_weight, self.k_proj_weight, self.v_proj_weight)
            bs = tuple(outs[-2:]) if b is not None else (None, None)
            sp = ws + bs
            object.__setattr__(self, '_split', sp)
        return sp

    def _weights(self):
        return self._splits()[:3]

    def _bias_q(self):
        return self._splits()[3]

    def _bias_kv(self):
        return self._splits()[4]

    def project_q(self, query2d):
        wq = self._weights()[0]
        bq = self._bias_q()
        return Fn.LinearFn.apply(query2d, wq, bq, self.scaling)      # q *= scaling (:353)

    def project_kv(self, key, need_weights=True):
        """key [S,B,kdim] -> fused [S*B, 2E] (k | v).  Returns None for an empty context
        ([.,.,0]: multi_head.py:349-352)."""
        if key.shape[2] == 0 or key.shape[0] == 0:
            return None
        _, wk, wv = self._weights()
        bkv = self._bias_kv()
        S, B, kd = key.shape
        return Fn.KVProjFn.apply(key.reshape(S * B, kd), wk, wv, bkv,
                                 Fn.kv16_ok(self.head_dim, need_weights))

    def attend(self, query, kv, key_padding_mask, need_weights=False):
        """query [T,B,E]; kv from project_kv.  Returns (out_proj input [T*B,E], weights)."""
        T, B, E = query.shape
        q = self.project_q(query.reshape(T * B, E))
        S = kv.shape[0] // B if kv is not None else 0
        mask = None
        if key_padding_mask is not None and S > 0:
            mask = key_padding_mask.to(torch.uint8).contiguous()
        p = self.dropout if self.training else 0.0
        out, weights = Fn.AttentionFn.apply(q, kv, self.bias_k, self.bias_v, mask, T, B, S,
                                            self.num_heads, self.add_zero_attn, p,
                                            config.next_seed() if p > 0 else 0, need_weights)
        return out, weights

    def forward(self, query, key, value, key_padding_mask=None, incremental_state=None,
                need_weights=True, static_kv=False, attn_mask=None):
        """query [T,B,E], key = value [S,B,kdim], key_padding_mask [B,S] (True = pad).
        Returns (attn [T,B,E], head-averaged weights [B,T,S+2] or None)."""
        T, B, E = query.shape
        assert E == self.embed_dim
        self.begin_step()
        if incremental_state is not None or attn_mask is not None:
            raise NotImplementedError('only the static_kv cross-attention mode of the decoder '
                                      '(incremental_state=None, attn_mask=None) is implemented')
        if value is not key and not (value.shape == key.shape and value.data_ptr() == key.data_ptr()):
            raise NotImplementedError('key and value must be the same context tensor')
        kv = self.project_kv(key, need_weights)
        a, weights = self.attend(query, kv, key_padding_mask, need_weights)
        out = linear(a, self.out_proj.weight, self.out_proj.bias).view(T, B, -1)
        return out, weights
