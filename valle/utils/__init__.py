"""valle/utils surface used by the hot path: make_pad_mask (icefall.utils, called at valle/models/valle.py:804-805),
Transpose (valle/utils/__init__.py:12-16), AttributeDict (icefall.utils, valle/bin/infer.py:133)."""
import torch
import torch.nn as nn


def make_pad_mask(lengths: torch.Tensor, max_len: int = 0) -> torch.Tensor:
    """bool [N, max_len], True at padded positions"""
    assert lengths.ndim == 1, lengths.ndim
    max_len = max(max_len, int(lengths.max()))
    return torch.arange(max_len, device=lengths.device)[None, :] >= lengths[:, None]


class Transpose(nn.Identity):
    """(N, T, D) -> (N, D, T)"""

    def forward(self, input: torch.Tensor) -> torch.Tensor:
        return input.transpose(1, 2)


class AttributeDict(dict):
    """dict with attribute access (icefall.utils.AttributeDict): `get_model(AttributeDict(checkpoint))`"""

    def __getattr__(self, key):
        if key in self:
            return self[key]
        raise AttributeError(f"No such attribute '{key}'")

    def __setattr__(self, key, value):
        self[key] = value

    def __delattr__(self, key):
        if key in self:
            del self[key]
            return
        raise AttributeError(f"No such attribute '{key}'")
