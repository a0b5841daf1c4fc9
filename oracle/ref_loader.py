"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Imports the *unmodified* reference modules from /root/reference (read-only) so
that golden vectors can be generated from the reference's own code
(SURVEY.md Appendix A).  /root/reference exists only in the build container, so
this loader is used by `oracle/gen_golden.py` (fixtures committed under
tests/golden/) and by CPU tests that are skipped when the tree is absent.

AllenNLP / overrides are not installed here; they are replaced by the minimal
stubs below (registry decorator = identity, TokenEmbedder = nn.Module).  The
`tell` package is injected as a namespace package so `tell/__init__.py`
(which imports allennlp) is bypassed.  Nothing is written under /root/reference.
"""
import os
import sys
import types

import torch.nn as nn

REF = os.environ.get('TT_REFERENCE_ROOT', '/root/reference')


def available():
    return os.path.isdir(os.path.join(REF, 'tell', 'modules'))


class _Registrable:
    @classmethod
    def register(cls, name):
        return lambda sub: sub


def _mod(name):
    m = types.ModuleType(name)
    sys.modules[name] = m
    return m


_loaded = False


def load():
    """Install the stubs and make `tell.modules`, `tell.models.decoder_*` importable."""
    global _loaded
    if _loaded:
        return
    if not available():
        raise RuntimeError('reference tree not present at %s' % REF)
    names = ['allennlp', 'allennlp.common', 'allennlp.common.registrable',
             'allennlp.common.checks', 'allennlp.data', 'allennlp.data.vocabulary',
             'allennlp.modules', 'allennlp.modules.text_field_embedders',
             'allennlp.modules.text_field_embedders.text_field_embedder',
             'allennlp.modules.time_distributed', 'allennlp.modules.token_embedders',
             'allennlp.modules.token_embedders.token_embedder', 'overrides']
    for n in names:
        _mod(n)
    sys.modules['allennlp.common'].Registrable = _Registrable
    sys.modules['allennlp.common'].Params = object
    sys.modules['allennlp.common.registrable'].Registrable = _Registrable
    sys.modules['allennlp.common.checks'].ConfigurationError = Exception
    sys.modules['allennlp.data'].Vocabulary = object
    sys.modules['allennlp.data.vocabulary'].Vocabulary = object

    class TokenEmbedder(nn.Module, _Registrable):
        pass

    class TextFieldEmbedder(nn.Module, _Registrable):
        pass

    for k in ['allennlp.modules.token_embedders',
              'allennlp.modules.token_embedders.token_embedder']:
        sys.modules[k].TokenEmbedder = TokenEmbedder
    for k in ['allennlp.modules.text_field_embedders',
              'allennlp.modules.text_field_embedders.text_field_embedder']:
        sys.modules[k].TextFieldEmbedder = TextFieldEmbedder
    sys.modules['allennlp.modules.time_distributed'].TimeDistributed = object
    sys.modules['overrides'].overrides = lambda f: f

    tell = _mod('tell')
    tell.__path__ = [REF + '/tell']
    import tell.utils  # noqa: F401
    import tell.modules  # noqa: F401
    models = _mod('tell.models')
    models.__path__ = [REF + '/tell/models']
    _loaded = True


def load_model_module():
    """Makes tell.models.transformer_faces_objects importable (class only; its constructor needs
    network access, so tests call its `_forward` / `_generate` methods unbound on a stub self)."""
    load()
    for n in ['allennlp.models', 'allennlp.models.model', 'allennlp.nn', 'allennlp.nn.initializers',
              'pycocoevalcap', 'pycocoevalcap.bleu', 'pycocoevalcap.bleu.bleu_scorer']:
        if n not in sys.modules:
            _mod(n)

    class Model(nn.Module, _Registrable):
        def __init__(self, vocab=None):
            super().__init__()

    sys.modules['allennlp.models.model'].Model = Model
    sys.modules['allennlp.models'].Model = Model
    sys.modules['allennlp.nn.initializers'].InitializerApplicator = lambda *a, **k: (lambda m: None)
    sys.modules['allennlp.nn'].InitializerApplicator = sys.modules['allennlp.nn.initializers'].InitializerApplicator
    sys.modules['pycocoevalcap.bleu.bleu_scorer'].BleuScorer = object
    import importlib
    return importlib.import_module('tell.models.transformer_faces_objects')


def build_embedder(vocab_size=50265, embed_dim=1024, cutoff=(5000, 20000), max_pos=512):
    load()
    from tell.modules.token_embedders import (AdaptiveEmbedding,
                                              SinusoidalPositionalEmbedding,
                                              SumTextFieldEmbedder)
    emb = SumTextFieldEmbedder(
        {'adaptive': AdaptiveEmbedding(None, 'bpe', 0, embed_dim, 1, embed_dim,
                                       list(cutoff), vocab_size, True),
         'position': SinusoidalPositionalEmbedding(None, embed_dim, 1, False, max_pos)},
        {'adaptive': ['roberta'], 'position': ['roberta']}, True)
    return emb


def build_decoder(kind='faces_objects', vocab_size=50265, embed_dim=1024, heads=16,
                  ffn=4096, kernels=(3, 7, 15, 31), cutoff=(5000, 20000), dropout=0.1,
                  attention_dropout=0.1, weight_dropout=0.1, input_dropout=0.1):
    """Reference decoder with the ctor argument order of
    tell/models/decoder_faces_objects.py:23-33 / decoder_flattened_no_image.py."""
    load()
    emb = build_embedder(vocab_size, embed_dim, cutoff)
    if kind == 'faces_objects':
        from tell.models.decoder_faces_objects import DynamicConvFacesObjectsDecoder as D
    elif kind == 'no_image':
        from tell.models.decoder_flattened_no_image import DynamicConvDecoderNoImage as D
    elif kind == 'flattened':
        from tell.models.decoder_flattened import DynamicConvDecoder as D
    elif kind == 'faces_parallel':
        from tell.models.decoder_faces_parallel import DynamicConvFacesParallelDecoder as D
    else:
        raise ValueError(kind)
    dec = D(None, emb, 512, dropout, True, embed_dim, embed_dim, True, 'dynamic', True,
            heads, weight_dropout, 0.0, input_dropout, False, attention_dropout, ffn,
            list(kernels), list(cutoff), True, 0, False, 1, len(kernels), False, 0,
            'bpe', vocab_size)
    return dec


def adaptive_loss():
    load()
    from tell.modules.criteria import AdaptiveLoss
    return AdaptiveLoss(1)
