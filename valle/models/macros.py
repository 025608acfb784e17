"""valle/models/macros.py constants."""
from valle_b200.models.macros import *  # noqa: F401,F403
