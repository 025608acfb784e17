"""GPU tests of the mirrored operator surface (valle/modules/*): every module's own forward(), called the way the
reference calls it (boolean attn_mask tensors, key_padding_mask, (x, stage_embedding) tuples, return_layer_states),
against the CPU restatement of the reference's arithmetic (oracle) / torch.nn.functional on the same weights."""
import pytest
import torch
import torch.nn.functional as F

from oracle import valle_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _mha_ref(mod, x, attn_mask=None, kpm=None):
    """what valle/modules/activation.py:408-427 calls"""
    out, _ = F.multi_head_attention_forward(
        x.transpose(0, 1), x.transpose(0, 1), x.transpose(0, 1), mod.embed_dim, mod.num_heads,
        mod.in_proj_weight.cpu(), mod.in_proj_bias.cpu(), None, None, False, 0.0,
        mod.out_proj.weight.cpu(), mod.out_proj.bias.cpu(), training=False, key_padding_mask=kpm,
        need_weights=False, attn_mask=attn_mask)
    return out.transpose(0, 1)


def test_multihead_attention_forward_bool_masks_and_key_padding():
    from valle.modules.activation import MultiheadAttention
    torch.manual_seed(0)
    d, H, B, Lq = 256, 4, 3, 37
    mod = MultiheadAttention(d, H, dropout=0.1, batch_first=True).eval()
    with torch.no_grad():
        mod.in_proj_bias.normal_(0, 0.1)
        mod.out_proj.bias.normal_(0, 0.1)
    mod = mod.to(DEV)
    x = torch.randn(B, Lq, d)
    ar = O.ar_inference_mask(9, Lq - 9)                                   # the mask VALLE.inference builds
    rnd = torch.rand(Lq, Lq) < 0.3
    rnd[torch.arange(Lq), torch.arange(Lq)] = False                        # no fully blocked row
    lens = torch.tensor([Lq, 20, 5])
    kpm = torch.arange(Lq)[None, :] >= lens[:, None]
    for name, am, kp in (("none", None, None), ("valle_ar", ar, None), ("dense", rnd, None), ("kpm", None, kpm),
                         ("dense+kpm", rnd, kpm)):
        xq = x.to(DEV)   # self-attention: the same tensor object for q, k, v (valle.py passes xy_dec three times)
        with torch.no_grad():
            got, _ = mod(xq, xq, xq, key_padding_mask=None if kp is None else kp.to(DEV), need_weights=False,
                         attn_mask=None if am is None else am.to(DEV))
            ref = _mha_ref(mod, x, am, kp)
        got = got.cpu()
        if kp is not None:   # padded query rows: the reference computes them too, the engine leaves them zero
            valid = ~kp
            assert torch.allclose(got[valid], ref[valid], atol=3e-5, rtol=1e-4), name
            assert float(got[kp].abs().max()) == 0.0
        else:
            assert torch.allclose(got, ref, atol=3e-5, rtol=1e-4), (name, float((got - ref).abs().max()))


def test_layernorm_and_adaptive_layernorm_forward_accept_tuples():
    from valle.modules.transformer import AdaptiveLayerNorm, LayerNorm
    torch.manual_seed(1)
    d = 256
    ln = LayerNorm(d)
    with torch.no_grad():
        ln.weight.normal_(1, 0.2)
        ln.bias.normal_(0, 0.2)
    x = torch.randn(2, 11, d) * 2 + 0.3
    emb = torch.randn(1, d)
    ref = F.layer_norm(x, (d,), ln.weight, ln.bias, 1e-5)
    ln = ln.to(DEV)
    with torch.no_grad():
        y = ln(x.to(DEV))
        yt, e = ln((x.to(DEV), emb.to(DEV)))
    assert torch.allclose(y.cpu(), ref, atol=2e-5, rtol=1e-5) and torch.equal(yt, y) and torch.equal(e.cpu(), emb)
    ada = AdaptiveLayerNorm(d, norm=torch.nn.LayerNorm(d))
    with torch.no_grad():
        ada.norm.weight.normal_(1, 0.2)
        ada.norm.bias.normal_(0, 0.2)
    ref = O.ada_layer_norm(x, emb, ada.project_layer.weight, ada.project_layer.bias, ada.norm.weight, ada.norm.bias)
    ada = ada.to(DEV)
    with torch.no_grad():
        y = ada(x.to(DEV), emb.to(DEV))
        yt, e = ada((x.to(DEV), emb.to(DEV)))
    assert torch.allclose(y.cpu(), ref.detach(), atol=1e-4, rtol=1e-5) and torch.equal(yt, y)


def test_token_and_sine_positional_embedding_forward():
    from valle.modules.embedding import SinePositionalEmbedding, TokenEmbedding
    torch.manual_seed(2)
    d = 256
    te = TokenEmbedding(d, 1025).eval()
    tok = torch.randint(0, 1025, (2, 13))
    ref = F.embedding(tok, te.weight)
    te = te.to(DEV)
    with torch.no_grad():
        got = te(tok.to(DEV))
    assert torch.equal(got.cpu(), ref.detach())
    assert torch.equal(te.embedding(5).cpu(), te.weight[5:6].cpu())
    from valle_b200 import ops
    te(torch.tensor([[1025]], device=DEV))           # clamped read + device flag ...
    with pytest.raises(IndexError):
        ops.check_oob(DEV)                           # ... reported at the next check
    pe = SinePositionalEmbedding(d, dropout=0.1, scale=False, alpha=True).eval()
    with torch.no_grad():
        pe.alpha.fill_(0.37)
    x = torch.randn(2, 13, d)
    ref = O.pos_embed(x, pe.alpha.detach())
    pe = pe.to(DEV)
    with torch.no_grad():
        got = pe(x.to(DEV))
    assert torch.equal(got.cpu(), ref)


@pytest.mark.parametrize("adaptive", [False, True])
def test_transformer_encoder_forward_masks_states_and_single_layer(adaptive):
    from valle.modules.transformer import AdaptiveLayerNorm, LayerNorm, TransformerEncoder, TransformerEncoderLayer
    torch.manual_seed(3)
    d, H, nl, B, Lq = 256, 4, 2, 2, 29
    enc = TransformerEncoder(
        TransformerEncoderLayer(d, H, dim_feedforward=4 * d, dropout=0.1, batch_first=True, norm_first=True,
                                adaptive_layer_norm=adaptive),
        num_layers=nl, norm=AdaptiveLayerNorm(d, norm=torch.nn.LayerNorm(d)) if adaptive else LayerNorm(d)).eval()
    with torch.no_grad():
        for p in enc.parameters():
            if p.dim() == 1:
                p.normal_(0.0, 0.1)
        for lyr in enc.layers:
            for n in (lyr.norm1, lyr.norm2):
                (n.norm if adaptive else n).weight.add_(1.0)
        (enc.norm.norm if adaptive else enc.norm).weight.add_(1.0)
    sd = {"enc." + k: v.detach().clone() for k, v in enc.state_dict().items()}
    cfg = O.OracleConfig(d, H, nl, 1, 8)
    x = torch.randn(B, Lq, d)
    stage = torch.randn(1, d) if adaptive else None
    ar = O.ar_inference_mask(7, Lq - 7)
    rnd = torch.rand(Lq, Lq) < 0.25
    rnd[torch.arange(Lq), torch.arange(Lq)] = False
    lens = torch.tensor([Lq, 17])
    kpm = torch.arange(Lq)[None, :] >= lens[:, None]
    enc = enc.to(DEV)
    src = (x.to(DEV), stage.to(DEV)) if adaptive else x.to(DEV)
    for name, am, kp in (("none", None, None), ("valle_ar", ar, None), ("dense", rnd, None), ("kpm", None, kpm)):
        with torch.no_grad():
            out = enc(src, mask=None if am is None else am.to(DEV),
                      src_key_padding_mask=None if kp is None else kp.to(DEV))
            ref = O.encoder(sd, "enc", x, cfg, blocked=am, key_padding=kp, stage_emb=stage)
        got = (out[0] if adaptive else out).cpu()
        if adaptive:
            assert torch.equal(out[1].cpu(), stage)
        valid = ~kp if kp is not None else torch.ones(B, Lq, dtype=torch.bool)
        err = float((got[valid] - ref[valid]).abs().max())
        assert err < 2e-4, (name, err)
    # return_layer_states: (states, output); the last state before the final norm, the output after it
    with torch.no_grad():
        states, out = enc(src, mask=ar.to(DEV), return_layer_states=True)
        ref = O.encoder(sd, "enc", x, cfg, blocked=ar, stage_emb=stage)
    assert len(states) == nl and states[0].shape == (B, Lq, d)
    assert float(((out[0] if adaptive else out).cpu() - ref).abs().max()) < 2e-4
    # one layer on its own (transformer.py:265-312)
    cfg1 = O.OracleConfig(d, H, 1, 1, 8)
    with torch.no_grad():
        l0 = enc.layers[0](src, src_mask=ar.to(DEV))
    l0 = (l0[0] if adaptive else l0).cpu()
    assert float((l0 - states[0].cpu()).abs().max()) < 2e-5
