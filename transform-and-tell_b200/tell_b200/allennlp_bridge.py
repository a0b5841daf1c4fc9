"""Registers the B200 classes under the reference's registry names on AllenNLP's own Registrable
bases, so that `tell train expt/nytimes/9_transformer_objects/config.yaml` constructs THESE modules
from the unchanged `model:` block (tell/commands/train.py:65-75 resolves `type:` through
allennlp.common.Registrable).  Import it once, after `import tell`:

    import tell_b200.allennlp_bridge

AllenNLP and the reference package are NOT dependencies of tell_b200: without them this module
raises ImportError with the reason; tell_b200.registry is the stand-alone equivalent used by the
tests and the benchmark."""
try:
    from allennlp.models import Model
    from allennlp.modules.text_field_embedders import TextFieldEmbedder
    from allennlp.modules.token_embedders import TokenEmbedder
    from tell.models.decoder_flattened import Decoder
    from tell.modules.criteria import Criterion
except ImportError as e:     # pragma: no cover - neither package exists in the build image
    raise ImportError('tell_b200.allennlp_bridge needs allennlp 0.9 and the reference `tell` package '
                      'on sys.path (%s); use tell_b200.registry for stand-alone construction' % e)

from . import models as M
from . import modules as O


def _register(base, name, cls):
    """Replace an existing registration (the reference registers the same names at import time)."""
    try:
        base.register(name, exist_ok=True)(cls)
    except TypeError:        # allennlp 0.9.0: register() has no exist_ok
        from allennlp.common.registrable import Registrable
        Registrable._registry[base].pop(name, None)
        base.register(name)(cls)


for _name, _cls in (('transformer_faces_objects', M.TransformerFacesObjectModel),
                    ('transformer_faces', M.TransformerFacesModel),
                    ('transformer_flattened', M.TransformerFlattenedModel)):
    _register(Model, _name, _cls)
for _name, _cls in M.decoder.Decoder.registered_names().items():
    _register(Decoder, _name, _cls)
_register(TextFieldEmbedder, 'sum', O.SumTextFieldEmbedder)
_register(TokenEmbedder, 'adaptive', O.AdaptiveEmbedding)
_register(TokenEmbedder, 'sinusoidal_positional', O.SinusoidalPositionalEmbedding)
_register(Criterion, 'adaptive_loss', O.AdaptiveLoss)
