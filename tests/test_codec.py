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


def test_random_encodec_weights_have_the_published_architecture():
    """product-side synthetic weights: same tensors (names, shapes) as transformers' 24 kHz EncodecModel"""
    from transformers import EncodecConfig, EncodecModel
    from valle_b200.data.tokenizer import random_encodec_weights
    sd = random_encodec_weights(0)
    ref = EncodecModel(EncodecConfig()).state_dict()
    n_conv = 0
    for k, v in sd.items():
        k2 = k.replace(".conv.weight", ".conv.parametrizations.weight.original1")
        assert k2 in ref and tuple(ref[k2].shape) == tuple(v.shape), k
        n_conv += k.endswith(".conv.weight")
    assert n_conv == sum(k.endswith("original1") for k in ref)


def test_compute_num_frames_matches_the_hop_count():
    from valle_b200.data.tokenizer import compute_num_frames
    for n in (320, 321, 24000, 239999, 240000, 240161, 479):
        exp = compute_num_frames(round(n / 24000, ndigits=12), 320.0 / 24000, 24000)
        assert abs(exp - -(-n // 320)) <= 1          # tokenizer.py:305: within one frame of what the codec emits


@pytest.mark.gpu
def test_persistent_lstm_layer_equals_the_stepwise_path():
    """one cooperative launch for all T steps (W_hh slices resident in shared memory, grid barrier per step) against the
    launch-per-step kernels: same arithmetic up to the k-group summation order"""
    from valle_b200 import _lib as L
    from valle_b200.data.tokenizer import AudioTokenizer
    m = E.build_codec(0)
    tok = AudioTokenizer(device="cuda:0", weights=m.state_dict())
    lstm = [mod for kind, mod in tok.codec.enc if kind == "lstm"][0]
    g = torch.Generator().manual_seed(8)
    for B, T in ((5, 37), (16, 75), (64, 20)):
        x = torch.randn(B, 512, T, generator=g).cuda()
        a = lstm(x)
        L.check(L.load().vb_tune_set(b"VB_LSTM_STEPWISE", 1))
        try:
            b = lstm(x)
        finally:
            L.load().vb_tune_set(b"VB_LSTM_STEPWISE", 0)
        assert float((a - b).abs().max()) < 2e-5, (B, T, float((a - b).abs().max()))


@pytest.mark.gpu
def test_extract_batch_ragged_waveforms_vs_oracle():
    """AudioTokenExtractor.extract_batch (tokenizer.py:326-361): ragged waveforms, zero-padded to the batch maximum as
    the reference does, trimmed to the expected frame count; codes against the HF oracle on the same padded batch."""
    from valle.data import AudioTokenExtractor, AudioTokenizer
    m = E.build_codec(0)
    ext = AudioTokenExtractor(tokenizer=AudioTokenizer(device="cuda:0", weights=m.state_dict()), max_batch=4)
    g = torch.Generator().manual_seed(4)
    lens = [24000, 7000, 15321, 9600, 12000]
    waves = [(torch.randn(1, n, generator=g) * 0.1).clamp(-1, 1) for n in lens]
    out = ext.extract_batch(waves, 24000, None)
    assert len(out) == len(lens)
    single = ext.extract(waves[2], 24000)
    assert single.shape == out[2].shape
    order = sorted(range(len(lens)), key=lambda i: -lens[i])
    agree, total = 0, 0
    for b0 in range(0, len(order), 4):
        ids = order[b0:b0 + 4]
        pad = torch.zeros(len(ids), 1, lens[ids[0]])
        for j, i in enumerate(ids):
            pad[j, 0, : lens[i]] = waves[i][0]
        ref, _ = E.encode(m, pad)                       # [B, 8, T]
        for j, i in enumerate(ids):
            o = torch.from_numpy(out[i])                # [T_i, 8]
            assert o.shape == (-(-lens[i] // 320), 8) or abs(o.shape[0] - -(-lens[i] // 320)) <= 1
            r = ref[j, :, : o.shape[0]].t()
            agree += int((o == r).all(dim=1).sum())
            total += o.shape[0]
    assert agree / total >= 0.97, agree / total
