"""Operator surface of tell/modules/__init__.py:1-11, B200-native."""
from .attention import MultiHeadAttention
from .convolutions import DynamicConv1dTBC, LightweightConv1dTBC
from .linear import GehringLinear, LayerNorm, Linear, TiedLinear
from .softmax import AdaptiveLoss, AdaptiveSoftmax, Criterion
from .token_embedders import (AdaptiveEmbedding, SinusoidalPositionalEmbedding,
                              SumTextFieldEmbedder, TextFieldEmbedder, TokenEmbedder)
