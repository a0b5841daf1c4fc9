from .decoder import (Decoder, DecoderLayer, DynamicConvDecoderLayer, DynamicConvDecoderNoImage,
                      DynamicConvFacesObjectsDecoder, DynamicConvFacesParallelDecoder,
                      DynamicConvFlattenedDecoder)
from .resnet import ResNetFeatureExtractor, resnet152
from .roberta import RobertaEncoder
from .transformer import (Model, TransformerFacesModel, TransformerFacesObjectModel,
                          TransformerFlattenedModel)
