"""valle/data surface on the hot path: the EnCodec tokenizer (valle/data/tokenizer.py:211-361)."""
from .tokenizer import *  # noqa: F401,F403
