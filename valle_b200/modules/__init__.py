from .activation import MultiheadAttention, ValleARMask
from .embedding import SinePositionalEmbedding, TokenEmbedding
from .transformer import (AdaptiveLayerNorm, LayerNorm, TransformerEncoder, TransformerEncoderLayer)
