"""tell/modules/token_embedders/{adaptive,positional,sum_text_field_embedder}.py on B200 kernels."""
import math

import torch
import torch.nn as nn

from .. import functional as Fn
from .. import ops
from ..registry import Registrable
from ..utils import get_incremental_state, set_incremental_state


class TokenEmbedder(nn.Module, Registrable):
    pass


class TextFieldEmbedder(nn.Module, Registrable):
    pass


class _Weight(nn.Module):
    def __init__(self, *shape):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(*shape))


@TokenEmbedder.register('adaptive')
class AdaptiveEmbedding(TokenEmbedder):
    """adaptive.py:12-80.  embeddings.{i}.0.weight [V_i, E_i] (row padding_idx zero),
    embeddings.{i}.1.weight [E, E_i]."""

    def __init__(self, vocab, namespace, padding_idx, initial_dim, factor, output_dim, cutoff,
                 vocab_size=None, scale_embeds=False):
        super().__init__()
        vocab_size = vocab_size or vocab.get_vocab_size(namespace)
        cutoff = list(cutoff)
        if not cutoff or vocab_size > cutoff[-1]:
            cutoff.append(vocab_size)
        assert vocab_size == cutoff[-1]
        if factor != 1:
            raise NotImplementedError('factor != 1 (shrinking band widths) is not used by the '
                                      'shipped configs')
        self.cutoff = cutoff
        self.embed_size = output_dim
        self.padding_idx = padding_idx
        self.embed_scale = math.sqrt(output_dim) if scale_embeds else 1
        self.embeddings = nn.ModuleList()
        for i in range(len(cutoff)):
            prev = cutoff[i - 1] if i > 0 else 0
            emb = _Weight(cutoff[i] - prev, initial_dim)
            emb.weight.data.normal_(mean=0, std=math.sqrt(1 / initial_dim))
            emb.weight.data[padding_idx].fill_(0)
            proj = _Weight(output_dim, initial_dim)
            nn.init.xavier_uniform_(proj.weight)
            self.embeddings.append(nn.ModuleList([emb, proj]))

    def weights_for_band(self, band):
        return self.embeddings[band][0].weight, self.embeddings[band][1].weight

    def tables(self):
        return [e[0].weight for e in self.embeddings], [e[1].weight for e in self.embeddings]

    def forward(self, X, incremental_state=None):
        """ids [B,T] -> [B,T,E] (public API layout)."""
        out = embed_tokens(X, self, None, 0)
        return ops_transpose_tb(out, X.shape[1], X.shape[0])

    def get_output_dim(self):
        return self.embed_size


# incremental_state key holding the running position as a device tensor (see start_position)
POSITION_DEV_KEY = '__tt_position_dev__'


@TokenEmbedder.register('sinusoidal_positional')
class SinusoidalPositionalEmbedding(TokenEmbedder):
    """positional.py:85-229; buffer `weights` [init_size+1, E], row padding_idx zero."""

    def __init__(self, vocab, embedding_dim, padding_idx, left_pad, init_size=1024):
        super().__init__()
        self.embedding_dim = embedding_dim
        self.padding_idx = padding_idx
        self.left_pad = left_pad
        self.register_buffer('weights', self.get_embedding(init_size + 1, embedding_dim,
                                                           padding_idx))

    @staticmethod
    def get_embedding(n_embeds, embed_dim, padding_idx=None):
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

    def start_position(self, incremental_state, seq_len):
        """positional.py:170-176: running position kept in the incremental state."""
        if incremental_state is None:
            return 0
        dev = incremental_state.get(POSITION_DEV_KEY)
        if dev is not None:          # captured decode step: int32 [1] device tensor, advanced by the caller
            return dev
        start = get_incremental_state(self, incremental_state, 'position') or 0
        set_incremental_state(self, incremental_state, 'position', start + seq_len)
        return start

    def ensure_size(self, max_pos):
        if max_pos > self.weights.shape[0]:
            w = self.get_embedding(max_pos, self.embedding_dim, self.padding_idx)
            self.register_buffer('weights', w.to(self.weights.device))

    def forward(self, X, incremental_state=None, timestep=None):
        B, T = X.shape
        start = self.start_position(incremental_state, T)
        if not torch.is_tensor(start):
            self.ensure_size(start + T + 1 + self.padding_idx)
        pos = ops.make_positions(X, self.padding_idx, self.left_pad, start, tbc=False)
        return ops.gather_rows(self.weights, pos.view(-1)).view(B, T, -1)

    def get_output_dim(self):
        return self.embedding_dim


def ops_transpose_tb(x2d, T, B):
    """[T*B, E] (row = t*B+b) -> [B,T,E]."""
    E = x2d.shape[-1]
    return Fn.Transpose01Fn.apply(x2d.view(T, B, E))


def embed_tokens(ids, adaptive, positional, start_pos):
    """Fused adaptive (+ positional) embedding in decoder layout: returns [T*B, E], row = t*B+b."""
    tables, projs = adaptive.tables()
    if positional is not None:
        pos_table, pad = positional.weights, positional.padding_idx
    else:
        # all-zero table: the position lookup contributes nothing
        pos_table = torch.zeros((ids.shape[1] + 2, adaptive.embed_size), device=ids.device)
        pad, start_pos = 0, 0
    return Fn.EmbedFn.apply(ids, pos_table, start_pos, pad, float(adaptive.embed_scale),
                            tuple(adaptive.cutoff), (len(tables), int(adaptive.padding_idx)), *tables, *projs)


@TextFieldEmbedder.register('sum')
class SumTextFieldEmbedder(TextFieldEmbedder):
    """sum_text_field_embedder.py:16-118: sums its token embedders' outputs.  The shipped
    configs use {'adaptive': AdaptiveEmbedding, 'position': SinusoidalPositionalEmbedding}, which
    is evaluated as one fused kernel sequence."""

    def __init__(self, token_embedders, embedder_to_indexer_map=None, allow_unmatched_keys=False):
        super().__init__()
        self._token_embedders = token_embedders
        self._embedder_to_indexer_map = embedder_to_indexer_map
        self._allow_unmatched_keys = allow_unmatched_keys
        for key, embedder in token_embedders.items():
            self.add_module('token_embedder_%s' % key, embedder)

    @classmethod
    def _from_params(cls, params, **extras):
        emb_params = params.pop('token_embedders')
        embedders = {k: TokenEmbedder.from_params(v, vocab=None) for k, v in emb_params.items()}
        return cls(embedders, params.pop('embedder_to_indexer_map', None),
                   params.pop('allow_unmatched_keys', False))

    def get_output_dim(self):
        return max(e.get_output_dim() for e in self._token_embedders.values())

    def _ids(self, text_field_input, key):
        if self._embedder_to_indexer_map is not None:
            return text_field_input[self._embedder_to_indexer_map[key][0]]
        return text_field_input[key]

    def embed_tbc(self, text_field_input, incremental_state=None):
        """Decoder-layout result [T*B, E] (row = t*B + b) without the public-API transpose."""
        keys = sorted(self._token_embedders.keys())
        adaptive = [k for k in keys if isinstance(self._token_embedders[k], AdaptiveEmbedding)]
        positional = [k for k in keys
                      if isinstance(self._token_embedders[k], SinusoidalPositionalEmbedding)]
        if len(adaptive) != 1 or len(positional) > 1 or len(adaptive) + len(positional) != len(keys):
            raise NotImplementedError('SumTextFieldEmbedder supports one adaptive embedder plus an '
                                      'optional sinusoidal positional embedder (shipped configs)')
        ids = self._ids(text_field_input, adaptive[0]).contiguous()
        ad = self._token_embedders[adaptive[0]]
        if not positional:
            return embed_tokens(ids, ad, None, 0), ids
        pe = self._token_embedders[positional[0]]
        if pe.left_pad:
            raise NotImplementedError('left-padded targets are not used by the shipped configs')
        start = pe.start_position(incremental_state, ids.shape[1])
        if not torch.is_tensor(start):
            pe.ensure_size(start + ids.shape[1] + 1 + pe.padding_idx)
        return embed_tokens(ids, ad, pe, start), ids

    def forward(self, text_field_input, num_wrapping_dims=0, incremental_state=None):
        out, ids = self.embed_tbc(text_field_input, incremental_state)
        return ops_transpose_tb(out, ids.shape[1], ids.shape[0])
