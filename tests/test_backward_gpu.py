"""Training backward (valle/bin/trainer.py:674 `scaler.scale(loss).backward()`): the gradient kernels
behind the C ABI against torch.autograd -- operator by operator on the same inputs, then the whole VALLE.forward loss
against the oracle (the reference's forward restated in plain torch, differentiated by torch on the host)."""
import ctypes as C
import random

import pytest
import torch
import torch.nn.functional as F

from conftest import assert_checksums, build_model, load_golden
from oracle import valle_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp(min=1e-12))


def _s():
    return torch.cuda.current_stream().cuda_stream


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-5), (torch.bfloat16, 2e-2)])
def test_linear_backward_vs_autograd(dtype, tol):
    from valle_b200 import _lib as L
    lib = L.load()
    torch.manual_seed(0)
    M, N, K = 333, 3072, 1024
    x = (torch.randn(M, K) * 0.5).to(dtype)
    w = (torch.randn(N, K) / 32).to(dtype)
    dy = (torch.randn(M, N) * 0.1).to(dtype)
    xr, wr = x.float().requires_grad_(), w.float().requires_grad_()
    b = torch.zeros(N, requires_grad=True)
    (F.linear(xr, wr, b) * dy.float()).sum().backward()
    xd, dyd = x.to(DEV), dy.to(DEV)
    wt = w.t().contiguous().to(DEV)
    dx = torch.empty(M, K, device=DEV)
    dw = torch.zeros(N, K, device=DEV)
    db = torch.zeros(N, device=DEV)
    dt = L.VB_F32 if dtype == torch.float32 else L.VB_BF16
    nb = lib.vb_linear_backward_workspace(dt, M, N, K)
    ws = torch.empty(nb, dtype=torch.uint8, device=DEV)
    L.check(lib.vb_linear_backward(xd.data_ptr(), dt, K, wt.data_ptr(), dyd.data_ptr(), N, dx.data_ptr(), L.VB_F32, K,
                                   L.VB_EPI_NONE, dw.data_ptr(), db.data_ptr(), M, N, K, ws.data_ptr(), nb, _s()))
    assert _rel(dx.cpu(), xr.grad) < tol and _rel(dw.cpu(), wr.grad) < tol and _rel(db.cpu(), b.grad) < tol
    # accumulate semantics: a second call doubles dW / db
    L.check(lib.vb_linear_backward(xd.data_ptr(), dt, K, wt.data_ptr(), dyd.data_ptr(), N, 0, L.VB_F32, K,
                                   L.VB_EPI_NONE, dw.data_ptr(), db.data_ptr(), M, N, K, ws.data_ptr(), nb, _s()))
    assert _rel(dw.cpu(), 2 * wr.grad) < tol and _rel(db.cpu(), 2 * b.grad) < tol


@pytest.mark.parametrize("adaptive", [False, True])
def test_layernorm_backward_vs_autograd(adaptive):
    from valle_b200 import _lib as L
    lib = L.load()
    torch.manual_seed(1)
    M, d = 301, 1024
    x = (torch.randn(M, d) * 2 + 0.3).requires_grad_()
    g = (torch.randn(d) * 0.2 + 1).requires_grad_()
    b = (torch.randn(d) * 0.2).requires_grad_()
    wb = (torch.randn(2 * d) * 0.3 + 0.5).requires_grad_() if adaptive else None
    rows = torch.randperm(M)[:200].to(torch.int32)
    dy = torch.randn(200, d)
    y = F.layer_norm(x[rows.long()], (d,), g, b, 1e-5)
    if adaptive:
        y = wb[:d] * y + wb[d:]
    (y * dy).sum().backward()
    dx = torch.full((M, d), 0.25, device=DEV)                    # accumulate semantics: dx += ...
    dg, db_, dwb = torch.zeros(d, device=DEV), torch.zeros(d, device=DEV), torch.zeros(2 * d, device=DEV)
    # (device copies are kept in variables: a temporary's memory may be reused before the kernel runs)
    xd, rd, gd, bd, dyd = x.detach().to(DEV), rows.to(DEV), g.detach().to(DEV), b.detach().to(DEV), dy.to(DEV)
    wbd = wb.detach().to(DEV) if adaptive else None
    L.check(lib.vb_layernorm_backward(xd.data_ptr(), d, rd.data_ptr(), 200, d, gd.data_ptr(), bd.data_ptr(),
                                      wbd.data_ptr() if adaptive else 0, 1e-5, dyd.data_ptr(), d, dx.data_ptr(), d, 0,
                                      L.VB_F32, dg.data_ptr(), db_.data_ptr(), dwb.data_ptr() if adaptive else 0, _s()))
    torch.cuda.synchronize()
    assert _rel(dx.cpu() - 0.25, x.grad) < 2e-5
    assert _rel(dg.cpu(), g.grad) < 2e-5 and _rel(db_.cpu(), b.grad) < 2e-5
    if adaptive:
        assert _rel(dwb.cpu(), wb.grad) < 2e-5


@pytest.mark.parametrize("mode", ["full", "valle_ar", "padded_ar", "padded"])
@pytest.mark.parametrize("dtype,tol", [(torch.float32, 3e-5), (torch.bfloat16, 3e-2)])
def test_attention_backward_vs_autograd(mode, dtype, tol):
    from valle_b200 import _lib as L
    lib = L.load()
    torch.manual_seed(2)
    H, d = 4, 256
    if mode in ("padded_ar", "padded"):
        B, Lp, seg1_start = 3, 150, 40
        lens, S, A = [Lp] * B, [40, 22, 9], [110, 75, 31]
    else:
        lens, S, A, seg1_start = [70, 5, 129, 200], [9, 2, 64, 30], None, 0
        B = len(lens)
    M = sum(lens)
    qkv = (torch.randn(M, 3 * d) * 0.7).to(dtype)
    dout = (torch.randn(M, d) * 0.3).to(dtype)
    cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32)
    mm = {"full": L.VB_MASK_FULL, "valle_ar": L.VB_MASK_VALLE_AR, "padded_ar": L.VB_MASK_PADDED_AR, "padded": L.VB_MASK_PADDED}[mode]
    tl = torch.tensor(S, dtype=torch.int32) if mode != "full" else None
    sl = torch.tensor(A, dtype=torch.int32) if A is not None else None
    # torch reference per sequence
    ref_d = torch.zeros(M, 3 * d)
    ref_o = torch.zeros(M, d)
    for b in range(B):
        r0, n = int(cu[b]), lens[b]
        blk = qkv[r0:r0 + n].float().clone().requires_grad_()
        q, k, v = blk[:, :d], blk[:, d:2 * d], blk[:, 2 * d:]
        qh, kh, vh = (t.reshape(n, H, 64).transpose(0, 1) for t in (q, k, v))
        sc = qh @ kh.transpose(-1, -2) * 0.125
        rows, cols = torch.arange(n)[:, None], torch.arange(n)[None, :]
        if mode == "valle_ar":
            ok = cols < torch.clamp(rows + 1, min=S[b])
        elif mode == "padded_ar":
            ok = (cols < S[b]) | ((rows >= seg1_start) & (cols >= seg1_start) & (cols < seg1_start + A[b]) & (cols <= rows))
            ok = ok | ((rows >= seg1_start + A[b]) & (cols >= seg1_start) & (cols < seg1_start + A[b]))
        elif mode == "padded":
            ok = ((cols < S[b]) | ((cols >= seg1_start) & (cols < seg1_start + A[b]))).expand(n, n)
        else:
            ok = torch.ones(n, n, dtype=torch.bool)
        sc = sc.masked_fill(~ok, float("-inf"))
        o = (torch.softmax(sc, -1) @ vh).transpose(0, 1).reshape(n, d)
        (o * dout[r0:r0 + n].float()).sum().backward()
        ref_d[r0:r0 + n] = blk.grad
        ref_o[r0:r0 + n] = o.detach()
    from valle_b200 import ops
    qd, cud, doutd = qkv.to(DEV), cu.to(DEV), dout.to(DEV)
    tld = tl.to(DEV) if tl is not None else None
    sld = sl.to(DEV) if sl is not None else None
    out = ops.attention(qd, cud, max(lens), H, mm, tld, seg1_lens=sld, seg1_start=seg1_start)
    assert _rel(out.float().cpu(), ref_o) < (3e-5 if dtype == torch.float32 else 2e-2)
    dq = torch.empty_like(qd)
    nb = lib.vb_attention_backward_workspace(M, H)
    ws = torch.empty(nb, dtype=torch.uint8, device=DEV)
    dt = L.VB_F32 if dtype == torch.float32 else L.VB_BF16
    L.check(lib.vb_attention_backward(qd.data_ptr(), out.data_ptr(), doutd.data_ptr(), dt, M, B, H, 64,
                                      cud.data_ptr(), tld.data_ptr() if tld is not None else 0,
                                      sld.data_ptr() if sld is not None else 0, seg1_start, max(lens), mm,
                                      dq.data_ptr(), ws.data_ptr(), nb, _s()))
    torch.cuda.synchronize()
    assert _rel(dq.float().cpu(), ref_d) < tol, _rel(dq.float().cpu(), ref_d)


def test_cross_entropy_and_embedding_backward_vs_autograd():
    from valle_b200 import autograd as AG
    torch.manual_seed(3)
    n, V = 77, 1025
    lg = (torch.randn(n, V) * 2).requires_grad_()
    tg = torch.randint(0, V, (n,))
    tg[::7] = 1024
    F.cross_entropy(lg, tg, ignore_index=1024, reduction="sum").mul(0.5).backward()
    l2 = lg.detach().to(DEV).requires_grad_()
    AG.CrossEntropySum.apply(l2, tg.to(DEV), 1024).mul(0.5).backward()
    assert _rel(l2.grad.cpu(), lg.grad) < 1e-5
    tabs = [torch.randn(1024, 256, requires_grad=True) for _ in range(3)]
    tok = torch.randint(0, 1024, (50, 3))
    ref = sum(t[tok[:, j]] for j, t in enumerate(tabs))
    dy = torch.randn(50, 256)
    (ref * dy).sum().backward()
    tabs_d = [t.detach().to(DEV).requires_grad_() for t in tabs]
    out = AG.EmbedSum.apply(tok.to(DEV), 3, 1, 50, *tabs_d)
    (out * dy.to(DEV)).sum().backward()
    for a, b in zip(tabs_d, tabs):
        assert _rel(a.grad.cpu(), b.grad) < 1e-5


def _oracle_grads(g, stage, nar_stage, prefix_len):
    m = build_model(g["config"], g["weight_seed"])
    sd = {k: v.detach().clone().requires_grad_() for k, v in m.state_dict().items()}
    c = g["config"]
    cfg = O.OracleConfig(c["d_model"], c["nhead"], c["num_layers"], c["prefix_mode"], c["num_quantizers"])
    fw = g["forward"]
    loss, _ = O.forward_train(sd, cfg, fw["x"], fw["x_lens"], fw["y"].long(), fw["y_lens"], nar_stage, prefix_len,
                              train_stage=stage)
    loss.backward()          # (forward_train already halves the stage-0 loss, valle.py:956-957)
    grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in sd.items()}
    return float(loss), grads


@pytest.mark.parametrize("stage", [0, 1, 2])
@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-3), (torch.bfloat16, 6e-2)])
def test_valle_forward_backward_matches_reference_autograd(stage, dtype, tol):
    """BASELINE configs[0] training batch (3 padded utterances, 32 phonemes, 128 x 8 codec tokens): loss.backward()
    through VALLE.forward in train() mode; every parameter gradient against torch.autograd of the oracle's forward
    (max-abs error relative to the gradient's max-abs: fp32 1e-3, bf16 6e-2); tied weights receive the sum."""
    g = load_golden("config0.pt")
    fw = g["forward"]
    ref_loss, ref = _oracle_grads(g, stage, int(fw["nar_stage"]), int(fw["prefix_len"]))
    m = build_model(g["config"], g["weight_seed"])
    assert_checksums(m, g["checksums"])
    m = m.to(DEV).train()
    m.engine_dtype = dtype
    m.rng = random.Random(0)
    torch.manual_seed(int(fw["torch_seed"]))
    (_, codes), loss, metrics = m(fw["x"].to(DEV), fw["x_lens"], fw["y"].long().to(DEV), fw["y_lens"], train_stage=stage)
    assert loss.requires_grad
    assert abs(float(loss) - ref_loss) <= (1e-4 if dtype == torch.float32 else 2e-2) * abs(ref_loss)
    loss.backward()
    names = dict(m.named_parameters())
    sd_keys = list(m.state_dict().keys())
    alias = {}                                             # state_dict key -> parameter name (tied weights)
    by_ptr = {p.data_ptr(): n for n, p in names.items()}
    for k, v in m.state_dict().items():
        alias[k] = by_ptr[v.data_ptr()]
    want = {}
    for k in sd_keys:
        want[alias[k]] = want.get(alias[k], 0) + ref[k]
    worst = ("", 0.0)
    checked = 0
    for n, p in names.items():
        if not p.requires_grad:          # e.g. the NAR positional alphas (alpha=False): frozen in the reference too
            assert p.grad is None
            continue
        gref = want[n]
        if float(gref.abs().max()) == 0.0:
            assert p.grad is None or float(p.grad.abs().max()) < 1e-6, n
            continue
        assert p.grad is not None, f"no gradient for {n}"
        e = _rel(p.grad.float().cpu(), gref)
        checked += 1
        if e > worst[1]:
            worst = (n, e)
    assert checked > 20
    assert worst[1] < tol, worst


def test_optimizer_step_changes_the_next_forward():
    """one SGD step on the gradients lowers the loss of the same batch (the engine re-packs changed parameters)"""
    g = load_golden("config0.pt")
    fw = g["forward"]
    m = build_model(g["config"], g["weight_seed"]).to(DEV).train()
    opt = torch.optim.SGD(m.parameters(), lr=2e-6)   # sum-reduced loss over ~800 target positions: large gradients
    losses = []
    for _ in range(3):
        m.rng = random.Random(0)
        torch.manual_seed(int(fw["torch_seed"]))
        (_, _), loss, _ = m(fw["x"].to(DEV), fw["x_lens"], fw["y"].long().to(DEV), fw["y_lens"], train_stage=0)
        opt.zero_grad()
        loss.backward()
        opt.step()
        losses.append(float(loss))
    assert losses[2] < losses[1] < losses[0], losses
