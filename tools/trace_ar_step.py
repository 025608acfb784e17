"""Device timeline of the AR decode step inside its CUDA graph (PDL-chained kernels), without a profiler attached.

    python -m valle_b200.build --trace
    VB_LIB_PATH=valle_b200/lib/libvalle_b200_trace.so python tools/trace_ar_step.py [B] [frames] [out.json]

The profiling build stamps %globaltimer when block 0 of every decode-step kernel passes its dependency wait and when
it finishes.  This tool decodes `frames` tokens at batch B (bf16), reads the ring of the last steps and prints, per
kernel of the chain, the time from its dependency being resolved to the next kernel's (= what the stage costs on the
critical path) and block 0's own duration.  ncu serialises the kernels and cannot see the PDL overlap; this can."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from valle_b200 import _lib  # noqa: E402

NAMES = {1: "ln_reduce", 2: "gemm_decode", 3: "attn_decode", 4: "relu_reduce", 5: "ar_sample", 6: "attn_combine",
         7: "fused"}


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    frames = int(sys.argv[2]) if len(sys.argv) > 2 else 430
    out = sys.argv[3] if len(sys.argv) > 3 else None
    lib = _lib.load()
    dev = torch.device("cuda:0")
    model = bench.build_model(dev)
    eng = model.engine(torch.bfloat16)
    eng.quiet = True
    texts, prompts = bench.make_batch(B, 5, dev)
    eng.generate(texts, prompts, top_k=1, max_new_tokens=40, return_device=True)   # warm-up + graph capture
    cap = 1 << 12
    ring = torch.zeros(cap, dtype=torch.int64, device=dev)
    cnt = torch.zeros(1, dtype=torch.int32, device=dev)
    _lib.check(lib.vb_trace_bind(ring.data_ptr(), cnt.data_ptr(), cap), "vb_trace_bind")
    eng.generate(texts, prompts, top_k=1, max_new_tokens=frames, return_device=True)
    torch.cuda.synchronize()
    ar_ms, steps = eng.stats.ar_ms, eng.stats.ar_steps
    n = int(cnt.item())
    lib.vb_trace_bind(0, 0, 1)
    raw = ring.cpu().tolist()
    seq = [raw[i % cap] for i in range(max(0, n - cap), n)]
    ev = [((v >> 8) & ((1 << 56) - 1), v & 0xff) for v in seq]
    ev.sort()
    # cut into steps at ar_sample starts
    starts = [(t, k >> 1) for t, k in ev if (k & 1) == 0]
    ends = {}
    for t, k in ev:
        if k & 1:
            ends.setdefault(k >> 1, []).append(t)
    # block 0's own end stamp that follows each start stamp of the same kernel kind (its duration inside the stage)
    own = {}
    pend = {}
    for t, k in ev:
        if (k & 1) == 0:
            pend[k >> 1] = t
        elif (k >> 1) in pend:
            own[pend.pop(k >> 1)] = t      # keyed by the start stamp
    idx = [i for i, (_, k) in enumerate(starts) if k == 5]
    steps_ev = [starts[idx[j] + 1: idx[j + 1] + 1] for j in range(len(idx) - 1)]
    steps_ev = [s for s in steps_ev if len(s) > 10]
    if not steps_ev:
        print("no complete step in the ring", n, len(ev))
        return
    L0 = len(steps_ev[-1])
    steps_ev = [s for s in steps_ev if len(s) == L0]
    print(f"B={B} frames={frames}: AR {ar_ms:.1f} ms / {steps} steps = {1000 * ar_ms / steps:.1f} us/step; "
          f"{len(steps_ev)} traced steps of {L0} kernels each")
    # per position in the chain: mean delta to the next kernel's start
    pos_stats = []
    for p in range(L0 - 1):
        d = [s[p + 1][0] - s[p][0] for s in steps_ev]
        pos_stats.append((NAMES.get(steps_ev[0][p][1], "?"), sum(d) / len(d) / 1000.0))
    total = sum(v for _, v in pos_stats)
    # 8 kernels per layer + LN / head / sampler, or (LayerNorm-folded chain) 6 per layer + head / sampler
    per_layer = (L0 - 3) // 12 if (L0 - 3) % 12 == 0 else ((L0 - 2) // 12 if (L0 - 2) % 12 == 0 else None)
    print(f"sum of stage times {total:.1f} us (block-0 stamps, last {len(steps_ev)} steps)")
    agg = {}
    for nm, v in pos_stats:
        agg.setdefault(nm, []).append(v)
    for nm, vs in agg.items():
        print(f"  {nm:14s} n={len(vs):3d}  mean {sum(vs) / len(vs):6.2f} us  total {sum(vs):7.1f} us ({100 * sum(vs) / total:4.1f} %)")
    # block 0: time from its dependency wait to its own end (the rest of the stage = the grid's tail + the hand-over)
    own_stats = []
    for p in range(L0 - 1):
        d = [own[s[p][0]] - s[p][0] for s in steps_ev if s[p][0] in own]
        own_stats.append(sum(d) / len(d) / 1000.0 if d else float("nan"))
    if per_layer:
        print(f"per-layer chain ({per_layer} kernels), mean over 12 layers: stage us | block 0 wait->end us")
        for j in range(per_layer):
            vs = [pos_stats[l * per_layer + j][1] for l in range(12)]
            os_ = [own_stats[l * per_layer + j] for l in range(12)]
            print(f"  [{j}] {pos_stats[j][0]:14s} {sum(vs) / 12:6.2f} us | {sum(os_) / 12:6.2f}")
    if out:
        json.dump(dict(B=B, frames=frames, ar_ms=ar_ms, steps=steps, kernels_per_step=L0,
                       stage_us=[dict(kernel=nm, us=v) for nm, v in pos_stats]), open(out, "w"), indent=1)


if __name__ == "__main__":
    main()
