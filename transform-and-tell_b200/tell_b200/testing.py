"""Helpers shared by tests, smoke() and bench.py: build a decoder from a synth config."""
from . import synth
from .modules import AdaptiveEmbedding, SinusoidalPositionalEmbedding, SumTextFieldEmbedder


def build_embedder(cfg):
    E = cfg['embed_dim']
    return SumTextFieldEmbedder(
        {'adaptive': AdaptiveEmbedding(None, 'bpe', 0, E, 1, E, list(cfg['cutoffs']), cfg['vocab'],
                                       True),
         'position': SinusoidalPositionalEmbedding(None, E, 1, False, cfg['max_pos'])},
        {'adaptive': ['roberta'], 'position': ['roberta']}, True)


def build_decoder(cfg, cls, state_dict=None, dropout=0.1):
    E = cfg['embed_dim']
    dec = cls(None, build_embedder(cfg), cfg['max_pos'], dropout, True, E, E, True, 'dynamic', True,
              cfg['heads'], dropout, 0.0, dropout, False, dropout, cfg['ffn'],
              list(cfg['kernels']), list(cfg['cutoffs']), True, 0, False, 1, len(cfg['kernels']),
              False, 0, 'bpe', cfg['vocab'])
    if state_dict is not None:
        dec.load_state_dict(state_dict, strict=True)
    return dec
