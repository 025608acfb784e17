"""valle/data/tokenizer.py surface: AudioTokenizer, tokenize_audio, AudioTokenExtractor."""
from valle_b200.data.tokenizer import *  # noqa: F401,F403
from valle_b200.data import tokenizer as _t

__all__ = [n for n in dir(_t) if not n.startswith("_")]
