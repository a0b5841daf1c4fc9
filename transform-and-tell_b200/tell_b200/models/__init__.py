from .decoder import (Decoder, DecoderLayer, DynamicConvDecoderLayer, DynamicConvDecoderNoImage,
                      DynamicConvFacesObjectsDecoder, DynamicConvFacesParallelDecoder,
                      DynamicConvFlattenedDecoder)
