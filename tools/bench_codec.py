"""BASELINE.json configs[4] (single GPU share): EnCodec 24 kHz encode + decode of B x 10 s synthetic waveforms
with seeded random weights; prints seconds of audio per second and the HF-on-CPU stand-in baseline."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import encodec_oracle as E  # noqa: E402  (weights + CPU baseline only)
from valle_b200.data.tokenizer import AudioTokenizer  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
m = E.build_codec(0)
tok = AudioTokenizer(device="cuda:0", weights=m.state_dict())
g = torch.Generator().manual_seed(5)
wav = (torch.randn(B, 1, 240000, generator=g) * 0.1).clamp(-1, 1)
wd = wav.cuda()
(codes, _), = tok.encode(wd)
tok.decode([(codes, None)])
torch.cuda.synchronize()
e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
e0.record()
(codes, _), = tok.encode(wd)
e1.record()
out = tok.decode([(codes, None)])
e2.record()
torch.cuda.synchronize()
enc_ms, dec_ms = e0.elapsed_time(e1), e1.elapsed_time(e2)
t0 = time.perf_counter()
ref_codes, _ = E.encode(m, wav[:2])
t_cpu_enc = (time.perf_counter() - t0) / 2
agree = (codes[:2].cpu() == ref_codes).float().mean().item()
print(json.dumps(dict(B=B, seconds_audio=10 * B, encode_ms=enc_ms, decode_ms=dec_ms,
                      encode_audio_s_per_s=10 * B / (enc_ms / 1e3), decode_audio_s_per_s=10 * B / (dec_ms / 1e3),
                      cpu_hf_encode_s_per_utt=t_cpu_enc, code_agreement_first2=agree,
                      cpu_threads=torch.get_num_threads())))
