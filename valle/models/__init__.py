"""valle/models/__init__.py:1-136 surface: model flags, factory, VALLE."""
from valle_b200.models import (NUM_AUDIO_TOKENS, NUM_MEL_BINS, NUM_SPEAKER_CLASSES, NUM_TEXT_TOKENS,  # noqa: F401
                               SPEAKER_EMBEDDING_DIM, VALLE, PromptedFeatures, add_model_arguments, get_model, str2bool)
from . import macros, valle  # noqa: F401
