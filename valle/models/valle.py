"""valle/models/valle.py surface: VALLE, topk_sampling, top_k_top_p_filtering."""
from valle_b200.models.valle import VALLE, PromptedFeatures, top_k_top_p_filtering, topk_sampling  # noqa: F401
