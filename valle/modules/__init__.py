"""valle/modules operator surface."""
from . import activation, embedding, transformer  # noqa: F401
