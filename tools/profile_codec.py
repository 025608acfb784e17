"""Per-layer CUDA-event timing of the EnCodec encoder / decoder stacks (B x 10 s, fp32, random weights)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from valle_b200.data.tokenizer import AudioTokenizer, random_encodec_weights  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
tok = AudioTokenizer(device="cuda:0", weights=random_encodec_weights(0))
g = torch.Generator().manual_seed(5)
wav = (torch.randn(B, 1, 240000, generator=g) * 0.1).clamp(-1, 1).cuda()


def run(stack, x, label):
    rows = []
    for rep in range(2):
        cur = x
        rows = []
        for kind, m in stack:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            if kind == "conv":
                y = m(cur)
            elif kind in ("conv_elu", "convt_elu"):
                y = m(cur, pre_elu=True)
            else:
                y = m(cur)
            e1.record()
            torch.cuda.synchronize()
            fl = 0.0
            if kind in ("conv", "conv_elu"):
                fl = 2.0 * m.cout * m.cin * m.k * y.shape[-1] * B
            elif kind == "convt_elu":
                fl = 2.0 * m.cout * m.cin * m.k * cur.shape[-1] * B
            elif kind == "res":
                for c in (m.c1, m.c2, m.sc):
                    fl += 2.0 * c.cout * c.cin * c.k * cur.shape[-1] * B
            elif kind == "lstm":
                fl = 2 * 2 * 2.0 * 2048 * 512 * cur.shape[-1] * B
            rows.append((kind, tuple(cur.shape), tuple(y.shape), e0.elapsed_time(e1), fl))
            cur = y
    tot = sum(r[3] for r in rows)
    print(f"== {label}: {tot:.2f} ms for B={B} ({tot / B:.3f} ms/utt)")
    for kind, si, so, ms, fl in rows:
        print(f"  {kind:10s} {str(si):22s} -> {str(so):22s} {ms:8.3f} ms  {fl / 1e9:8.1f} GFLOP  {fl / ms / 1e9 if ms > 0 else 0:7.2f} TFLOP/s")
    return cur


emb = run(tok.codec.enc, wav, "encoder")
(codes, _), = tok.encode(wav)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
(codes, _), = tok.encode(wav)
e1.record()
torch.cuda.synchronize()
print("encode total (incl. RVQ)", e0.elapsed_time(e1), "ms")
x = torch.randn(B, 128, 750, device="cuda:0")
run(tok.codec.dec, x, "decoder")
