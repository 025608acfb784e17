"""TEST INFRASTRUCTURE ONLY -- codec oracle for `valle_b200.data.AudioTokenizer`.

PARITY UNPINNED: the reference obtains the EnCodec arithmetic from the un-vendored, unpinned PyPI package
`encodec` (valle/data/tokenizer.py:23-24,219-221,239-242) whose weights are downloaded at run time; neither
the package nor the weights exist offline, and the reference's only check at this boundary is a
`__main__` self-test that needs the weights (tokenizer.py:364-376).  The stand-in oracle is
`transformers.EncodecModel(EncodecConfig())` -- the same published 24 kHz architecture (ratios 8*5*4*2 = 320,
32 filters, causal reflect-padded convs, weight-norm, 2-layer LSTM, 1024 x 128 Euclidean codebooks,
8 codebooks at 6 kbps) -- with SEEDED RANDOM weights and random codebooks (the default init zeros the
codebooks).  Identical tensors are loaded into the engine, so the comparison checks the kernels, not the
training.  Only tests / smoke / bench baselines may import this module.
"""
from __future__ import annotations

import torch


def build_codec(seed: int = 0):
    from transformers import EncodecConfig, EncodecModel
    torch.manual_seed(seed)
    m = EncodecModel(EncodecConfig()).eval()
    g = torch.Generator().manual_seed(seed + 1)
    # Codebooks placed where the (random-weight) encoder output actually lives: stage 0 around the
    # per-dimension mean/std of the embeddings of a calibration signal, later stages zero-mean with a
    # shrinking scale like real residuals -- otherwise one code wins every frame and the check is vacuous.
    with torch.no_grad():
        cal = (torch.randn(2, 1, 24000, generator=g) * 0.3).clamp(-1, 1)
        emb = m.encoder(cal)                                   # [2, 128, 75]
        mu = emb.mean(dim=(0, 2))
        sd = emb.std(dim=(0, 2)) + 1e-6
        for q, layer in enumerate(m.quantizer.layers):
            e = torch.randn(layer.codebook.embed.shape, generator=g) * sd * (0.8 ** q)
            if q == 0:
                e = e + mu
            layer.codebook.embed.copy_(e)
    return m


@torch.no_grad()
def encode(m, wav: torch.Tensor, bandwidth: float = 6.0):
    """wav [B,1,N] -> (codes [B,8,T'], encoder embeddings [B,128,T'])."""
    emb = m.encoder(wav)
    codes = m.quantizer.encode(emb, bandwidth)            # [n_q, B, T']
    return codes.transpose(0, 1).contiguous(), emb


@torch.no_grad()
def decode(m, codes: torch.Tensor) -> torch.Tensor:
    """codes [B,8,T'] -> wav [B,1,T'*320]."""
    emb = m.quantizer.decode(codes.transpose(0, 1))
    return m.decoder(emb)


@torch.no_grad()
def rvq_margins(m, emb: torch.Tensor, n_q: int = 8) -> torch.Tensor:
    """top-1 minus top-2 (negative squared) distance per frame and stage: [n_q, B, T']."""
    residual = emb
    out = []
    for layer in m.quantizer.layers[:n_q]:
        x = residual.permute(0, 2, 1)
        e = layer.codebook.embed
        dist = -(x.pow(2).sum(-1, keepdim=True) - 2 * x @ e.t() + e.pow(2).sum(1)[None, None])
        top2 = dist.topk(2, dim=-1).values
        out.append(top2[..., 0] - top2[..., 1])
        idx = dist.argmax(-1)
        residual = residual - layer.codebook.embed[idx].permute(0, 2, 1)
    return torch.stack(out)
