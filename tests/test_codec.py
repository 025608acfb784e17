"""EnCodec path (AudioTokenizer.encode/decode) against the HF EncodecModel oracle with identical
seeded random weights.  fp32; encoder embeddings <= 2e-4 * scale, codes identical wherever the
oracle's RVQ top-2 margin exceeds 1e-3, waveform <= 2e-3 * scale."""
import pytest
import torch

from oracle import encodec_oracle as E


def test_oracle_codec_shapes_cpu():
    m = E.build_codec(0)
    wav = torch.randn(2, 1, 4800).clamp(-1, 1) * 0.1
    codes, emb = E.encode(m, wav)
    assert codes.shape == (2, 8, 15) and emb.shape == (2, 128, 15)
    assert codes.min() >= 0 and codes.max() < 1024 and codes.unique().numel() > 20
    assert E.decode(m, codes).shape == (2, 1, 4800)


@pytest.mark.gpu
def test_audio_tokenizer_encode_decode_vs_oracle():
    from valle_b200.data.tokenizer import AudioTokenizer
    m = E.build_codec(0)
    tok = AudioTokenizer(device="cuda:0", weights=m.state_dict())
    assert tok.sample_rate == 24000 and tok.channels == 1
    g = torch.Generator().manual_seed(3)
    wav = (torch.randn(3, 1, 24000 + 123, generator=g) * 0.1).clamp(-1, 1)   # ragged tail: extra padding path
    ref_codes, ref_emb = E.encode(m, wav)
    # encoder stack
    emb = tok.codec._run(tok.codec.enc, wav.cuda()).cpu()
    scale = ref_emb.abs().max().item()
    assert emb.shape == ref_emb.shape
    assert (emb - ref_emb).abs().max().item() <= 2e-4 * scale, (emb - ref_emb).abs().max().item() / scale
    # codes
    (codes, none), = tok.encode(wav)
    codes = codes.cpu()
    assert none is None and codes.shape == ref_codes.shape and codes.dtype == torch.int64
    margins = E.rvq_margins(m, ref_emb).transpose(0, 1)            # [B, 8, T']
    diff = codes != ref_codes
    # a flipped early stage changes the residual of every later stage: compare stage by stage
    first_bad = diff.float().cumsum(1) > 0
    assert not (diff[:, 0] & (margins[:, 0] > 1e-3)).any()
    frac = 1.0 - first_bad.any(1).float().mean().item()
    assert frac >= 0.97, frac
    # decoder on the oracle's codes
    wav_ref = E.decode(m, ref_codes)
    wav_out = tok.decode([(ref_codes.cuda(), None)]).cpu()
    assert wav_out.shape == wav_ref.shape
    s = wav_ref.abs().max().item()
    assert (wav_out - wav_ref).abs().max().item() <= 2e-3 * s, (wav_out - wav_ref).abs().max().item() / s


@pytest.mark.gpu
def test_audio_tokenizer_requires_weights_and_cuda():
    from valle_b200 import _lib
    from valle_b200.data.tokenizer import AudioTokenizer
    with pytest.raises(_lib.VbError):
        AudioTokenizer()
