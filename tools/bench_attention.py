"""Microbenchmark of the ragged bf16 flash attention at the NAR shapes of the bench (B x L, 16 heads of 64):
timed (optionally once per value of a tuning knob, VB_KNOB=<name>, in one process) and checked against an fp32 softmax(QK^T)V of the same
bf16 inputs on a sample of (sequence, head) pairs.  usage: bench_attention.py [B L [values...]]   (VB_KNOB=<name> selects the knob)"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from valle_b200 import _lib as L, ops  # noqa: E402

dev = "cuda:0"
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
Ls = int(sys.argv[2]) if len(sys.argv) > 2 else 1025
variants = [int(v) for v in sys.argv[3:]] or [0]
knob = os.environ.get("VB_KNOB", "")
H, d = 16, 1024
lib = L.load()
torch.manual_seed(0)
qkv = (torch.randn(B * Ls, 3 * d, device=dev) * 0.5).bfloat16()
# one sequence with peaked scores: exercises the rescale path (row maximum growing by more than 2^8)
qkv[:Ls, :d] *= 6.0
cu = (torch.arange(B + 1, dtype=torch.int32, device=dev) * Ls).contiguous()
out = torch.empty(B * Ls, d, dtype=torch.bfloat16, device=dev)
fl = 4.0 * B * H * Ls * Ls * 64


def reference(b, h):
    x = qkv[b * Ls:(b + 1) * Ls].float().view(Ls, 3, H, 64)
    q, k, v = x[:, 0, h], x[:, 1, h], x[:, 2, h]
    p = torch.softmax(q @ k.t() * 0.125, dim=-1)
    return p @ v


def timed(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for v in variants:
    if knob:
        lib.vb_tune_set(knob.encode(), v)
    out.zero_()
    ms = timed(lambda: ops.attention(qkv, cu, Ls, H, L.VB_MASK_FULL, None, out=out))
    err = 0.0
    for (b, h) in [(0, 0), (0, 5), (1, 3), (B - 1, H - 1)]:
        ref = reference(b, h)
        got = out[b * Ls:(b + 1) * Ls].view(Ls, H, 64)[:, h].float()
        err = max(err, float((got - ref).abs().max()))
    print(json.dumps(dict(knob=knob, value=v, B=B, L=Ls, ms=round(ms, 4), tflops=round(fl / ms / 1e9, 1), max_abs_err=err)), flush=True)

# library reference for context only (not the product): torch SDPA on the same shape
q = qkv.view(B, Ls, 3, H, 64).permute(2, 0, 3, 1, 4).contiguous()
ms_ref = timed(lambda: torch.nn.functional.scaled_dot_product_attention(q[0], q[1], q[2]))
print(json.dumps(dict(variant="torch_sdpa", B=B, L=Ls, ms=round(ms_ref, 4), tflops=round(fl / ms_ref / 1e9, 1))))
