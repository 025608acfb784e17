"""valle/modules/embedding.py surface: TokenEmbedding, SinePositionalEmbedding."""
from valle_b200.modules.embedding import SinePositionalEmbedding, TokenEmbedding  # noqa: F401
