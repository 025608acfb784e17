"""valle/modules/activation.py surface: MultiheadAttention."""
from valle_b200.modules.activation import MultiheadAttention, ValleARMask  # noqa: F401
