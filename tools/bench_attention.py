"""Microbenchmark of the ragged bf16 flash attention at the NAR shape of the bench (B x L, 16 heads of 64)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from valle_b200 import _lib as L, ops  # noqa: E402

dev = "cuda:0"
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
Ls = int(sys.argv[2]) if len(sys.argv) > 2 else 1025
H, d = 16, 1024
qkv = (torch.randn(B * Ls, 3 * d, device=dev) * 0.5).bfloat16()
cu = (torch.arange(B + 1, dtype=torch.int32, device=dev) * Ls).contiguous()
out = torch.empty(B * Ls, d, dtype=torch.bfloat16, device=dev)
for _ in range(3):
    ops.attention(qkv, cu, Ls, H, L.VB_MASK_FULL, None, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 10
e0.record()
for _ in range(n):
    ops.attention(qkv, cu, Ls, H, L.VB_MASK_FULL, None, out=out)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
fl = 4.0 * B * H * Ls * Ls * 64
# library reference for context only (not the product): torch SDPA on the same shape
q = qkv.view(B, Ls, 3, H, 64).permute(2, 0, 3, 1, 4).contiguous()
for _ in range(3):
    torch.nn.functional.scaled_dot_product_attention(q[0], q[1], q[2])
e0.record()
for _ in range(n):
    torch.nn.functional.scaled_dot_product_attention(q[0], q[1], q[2])
e1.record()
torch.cuda.synchronize()
ms_ref = e0.elapsed_time(e1) / n
print(json.dumps(dict(B=B, L=Ls, ms=ms, tflops=fl / ms / 1e9, sdpa_ms=ms_ref, sdpa_tflops=fl / ms_ref / 1e9)))
