"""CPU tests of the oracle: (1) against the reference itself when /root/reference is mounted
(build container), (2) against the committed reference-generated golden fixtures (always)."""
import random

import pytest
import torch

from conftest import assert_checksums, build_model, load_golden
from oracle import valle_oracle as O
from oracle.ref_loader import load_reference, reference_available

needs_ref = pytest.mark.skipif(not reference_available(), reason="/root/reference not mounted")


def _sd(m):
    return {k: v.detach() for k, v in m.state_dict().items()}


@pytest.mark.parametrize("name", ["tiny_pm0.pt", "tiny_pm1.pt", "tiny_pm2.pt"])
def test_oracle_reproduces_golden_codes(name):
    g = load_golden(name)
    m = build_model(g["config"], g["weight_seed"])
    assert_checksums(m, g["checksums"])      # our init == the reference's init, bit for bit
    cfg = O.OracleConfig(**g["config"])
    x, y = g["x"], g["y"]
    xl = torch.tensor([x.shape[1]], dtype=torch.int32)
    enroll = torch.tensor([g["enroll"]], dtype=torch.int32) if "enroll" in g else None
    tr = O.InferenceTrace([], [], [], [])
    with torch.no_grad():
        codes = O.inference(_sd(m), cfg, x, xl, y, enroll, top_k=1, trace=tr)
        cont = O.continual(_sd(m), cfg, x, xl, y)
    assert torch.equal(codes, g["codes"].long())
    assert torch.equal(cont, g["continual"].long())
    steps = g["ar_logit_steps"].tolist()
    got = torch.stack([tr.ar_logits[i] for i in steps])
    assert torch.allclose(got, g["ar_logits"], atol=1e-5, rtol=0)


def test_kv_cache_equals_full_recompute():
    """The engine's algorithm (prefill + cached single-row steps) is exact under the reference's
    mask (valle.py:1019-1030): same greedy tokens, logits within 1e-5."""
    g = load_golden("tiny_pm1.pt")
    m = build_model(g["config"], g["weight_seed"])
    cfg = O.OracleConfig(**g["config"])
    x, y = g["x"], g["y"]
    xl = torch.tensor([x.shape[1]], dtype=torch.int32)
    with torch.no_grad():
        toks, logits = O.ar_decode_kv(_sd(m), cfg, x, xl, y, collect_logits=True)
    assert torch.equal(toks, g["codes"][..., 0].long())
    steps = g["ar_logit_steps"].tolist()
    got = torch.stack([logits[i] for i in steps])
    assert torch.allclose(got, g["ar_logits"], atol=2e-5, rtol=0)


def test_oracle_forward_matches_golden_loss():
    g = load_golden("config0.pt")
    m = build_model(g["config"], g["weight_seed"])
    cfg = O.OracleConfig(**g["config"])
    fw = g["forward"]
    for stage in (0, 1, 2):
        with torch.no_grad():
            loss, _ = O.forward_train(_sd(m), cfg, fw["x"], fw["x_lens"], fw["y"].long(), fw["y_lens"],
                                      fw["nar_stage"], fw["prefix_len"], train_stage=stage)
        ref = float(fw[f"loss_stage{stage}"])
        assert abs(float(loss) - ref) <= 1e-4 * abs(ref), (stage, float(loss), ref)


def test_sampling_restatement_greedy_is_argmax():
    torch.manual_seed(0)
    logits = torch.randn(4, 1025)
    tok = O.topk_sampling(logits.clone(), top_k=1)
    assert torch.equal(tok[:, 0], logits.argmax(-1))
    # top_k <= 0 leaves the distribution unfiltered (valle.py:1254)
    assert torch.equal(O.top_k_top_p_filtering(logits.clone(), top_k=-100), logits)


@needs_ref
@pytest.mark.parametrize("pm", [0, 1])
def test_oracle_vs_reference_inference_and_forward(pm):
    ref = load_reference()
    torch.manual_seed(0)
    m = ref.VALLE(256, 4, 2, norm_first=True, add_prenet=False, prefix_mode=pm, share_embedding=True,
                  nar_scale_factor=1.0, prepend_bos=False, num_quantizers=8).eval()
    sd = _sd(m)
    cfg = O.OracleConfig(256, 4, 2, pm, 8)
    g = torch.Generator().manual_seed(21)
    x = torch.randint(3, 100, (1, 7), generator=g)
    y = torch.randint(0, 1024, (1, 15, 8), generator=g)
    xl = torch.tensor([7], dtype=torch.int32)
    with torch.no_grad():
        assert torch.equal(m.inference(x, xl, y, None, top_k=1), O.inference(sd, cfg, x, xl, y, None, top_k=1))
        assert torch.equal(m.continual(x, xl, y), O.continual(sd, cfg, x, xl, y))
    # sampled decoding consumes torch's RNG identically
    with torch.no_grad():
        torch.manual_seed(3)
        a = m.inference(x, xl, y, None, top_k=5, temperature=0.9)
        torch.manual_seed(3)
        b = O.inference(sd, cfg, x, xl, y, None, top_k=5, temperature=0.9)
    assert torch.equal(a, b)
    N = 3
    xx = torch.randint(3, 100, (N, 9), generator=g)
    xls = torch.tensor([9, 7, 5], dtype=torch.int32)
    yy = torch.randint(0, 1024, (N, 40, 8), generator=g)
    yls = torch.tensor([40, 33, 28], dtype=torch.int32)
    m.rng = random.Random(0)
    ns = random.Random(0).choices(list(range(1, 8)), weights=[1 / 7] * 7, k=1)[0]
    torch.manual_seed(5)
    with torch.no_grad():
        (_, _), loss, _ = m(xx, xls, yy, yls)
    torch.manual_seed(5)
    int_low = (0.25 * yls.min()).type(torch.int64).item()
    pl = min(torch.randint(int_low, int_low * 2, size=()).item(), 225) if pm == 1 else 0
    with torch.no_grad():
        l2, _ = O.forward_train(sd, cfg, xx, xls, yy, yls, ns, pl)
    assert abs(float(loss) - float(l2)) <= 1e-5 * abs(float(loss))


@needs_ref
@pytest.mark.parametrize("prenet", [False, True])
def test_model_mirror_matches_reference_checkpoint_layout(prenet):
    """valle_b200.models.VALLE: same state_dict keys, shapes, init values and strict loading -- also with the
    pre-nets of add_prenet=True (Conv1d / BatchNorm1d / Linear containers at the reference's Sequential indices)."""
    ref = load_reference()
    cfg = dict(d_model=256, nhead=4, num_layers=2, prefix_mode=1, num_quantizers=8, add_prenet=prenet)
    torch.manual_seed(0)
    a = ref.VALLE(256, 4, 2, norm_first=True, add_prenet=prenet, prefix_mode=1, share_embedding=True,
                  nar_scale_factor=1.0, prepend_bos=False, num_quantizers=8)
    b = build_model(cfg, 0)
    sa, sb = a.state_dict(), b.state_dict()
    assert list(sa.keys()) == list(sb.keys())
    assert all(torch.equal(sa[k], sb[k]) for k in sa)
    b.load_state_dict(sa, strict=True)
    a.load_state_dict(sb, strict=True)
    assert [n for n, _ in a.named_parameters()] == [n for n, _ in b.named_parameters()]
    assert [n for n, _ in a.named_buffers()] == [n for n, _ in b.named_buffers()]
    assert (len(list(b.named_buffers())) == 0) == (not prenet)
