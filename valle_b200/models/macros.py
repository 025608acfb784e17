# constants of valle/models/macros.py (part of the checkpoint / token-id contract)
NUM_TEXT_TOKENS = 512     # macros.py:2
NUM_AUDIO_TOKENS = 1024   # macros.py:5  EnCodec RVQ bins
NUM_MEL_BINS = 100
NUM_SPEAKER_CLASSES = 4096
SPEAKER_EMBEDDING_DIM = 64
