"""Host-side logic of the batched engine that needs no GPU: the packed index maps (rows / positions of the
text and audio segments of every utterance) that generate() / _nar() ship to the device in one copy."""
import numpy as np

from valle_b200.engine import _seg_ranges


def _naive(starts, lens):
    rows = [s + i for s, n in zip(starts, lens) for i in range(n)]
    pos = [i for n in lens for i in range(n)]
    return rows, pos


def test_seg_ranges_matches_python_loops():
    rng = np.random.default_rng(0)
    for _ in range(50):
        B = int(rng.integers(1, 9))
        lens = rng.integers(0, 40, size=B).tolist()   # ragged, empty segments included
        starts = rng.integers(0, 1000, size=B).tolist()
        rows, pos = _seg_ranges(starts, lens)
        exp_rows, exp_pos = _naive(starts, lens)
        assert rows.tolist() == exp_rows and pos.tolist() == exp_pos
        assert rows.dtype == np.int64 and pos.dtype == np.int64


def test_seg_ranges_empty_and_packed_layout():
    rows, pos = _seg_ranges([5, 9], [0, 0])
    assert rows.size == 0 and pos.size == 0
    # the packed [text_b | audio_b] layout of generate(): text rows then audio rows of each utterance
    S, T = [3, 2], [4, 1]
    cu = np.concatenate([[0], np.cumsum(np.add(S, T))])
    trow, tpos = _seg_ranges(cu[:-1], S)
    arow, apos = _seg_ranges(cu[:-1] + np.asarray(S), T)
    assert trow.tolist() == [0, 1, 2, 7, 8] and tpos.tolist() == [0, 1, 2, 0, 1]
    assert arow.tolist() == [3, 4, 5, 6, 9] and apos.tolist() == [0, 1, 2, 3, 0]
    assert sorted(trow.tolist() + arow.tolist()) == list(range(int(cu[-1])))


def test_layernorm_fold_identity_of_the_decode_chain():
    """The algebra behind vb_ln_fold (include/valle_b200.h): LayerNorm(x) W^T + b = rstd (x (W gamma)^T - mean c) + d with
    c = row sums of the folded weights and d = b + W beta, the moments taken over split k-ranges and added up -- what
    gemm_decode_x_kernel and its consumers compute on the device (here in torch fp64 against F.layer_norm, and with the
    bf16 roundings of the device path against its own tolerance)."""
    import torch
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(3)
    B, K, N, eps = 5, 256, 96, 1e-5
    x = torch.randn(B, K, generator=g, dtype=torch.float64) * 1.7 + 0.3
    W = torch.randn(N, K, generator=g, dtype=torch.float64) / 16
    gamma = 1 + 0.1 * torch.randn(K, generator=g, dtype=torch.float64)
    beta = 0.1 * torch.randn(K, generator=g, dtype=torch.float64)
    b = torch.randn(N, generator=g, dtype=torch.float64)
    ref = F.linear(F.layer_norm(x, (K,), gamma, beta, eps), W, b)
    wf = W * gamma
    c, d = wf.sum(1), b + W @ beta
    # moments as per-split partial sums over k-ranges (4 splits), then added up
    s1 = sum(x[:, i:i + 64].sum(1) for i in range(0, K, 64))
    s2 = sum((x[:, i:i + 64] ** 2).sum(1) for i in range(0, K, 64))
    mean = s1 / K
    rstd = torch.rsqrt(s2 / K - mean * mean + eps)
    acc = sum(x[:, i:i + 64] @ wf[:, i:i + 64].T for i in range(0, K, 64))
    out = rstd[:, None] * (acc - mean[:, None] * c[None, :]) + d[None, :]
    assert torch.allclose(out, ref, atol=1e-10, rtol=1e-10)
    # device roundings: x and W gamma in bf16, c summed from the ROUNDED folded weights, fp32 accumulation
    xb, wfb = x.float().bfloat16().float(), wf.float().bfloat16().float()
    accb = xb @ wfb.T
    outb = rstd.float()[:, None] * (accb - mean.float()[:, None] * wfb.sum(1)[None, :]) + d.float()[None, :]
    assert (outb.double() - ref).abs().max() < 3e-2
