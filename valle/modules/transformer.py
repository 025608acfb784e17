"""valle/modules/transformer.py surface: LayerNorm, AdaptiveLayerNorm, TransformerEncoderLayer, TransformerEncoder."""
from valle_b200.modules.transformer import (AdaptiveLayerNorm, LayerNorm, TransformerEncoder,  # noqa: F401
                                            TransformerEncoderLayer)
