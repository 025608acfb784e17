"""Per-CTA timeline of the KV-cache attention launches of the AR decode step (profiling build, every CTA stamps
%globaltimer when it becomes resident, when its dependency resolves and when it is done).

    python -m valle_b200.build --trace
    VB_LIB_PATH=valle_b200/lib/libvalle_b200_trace.so python tools/trace_attn_ctas.py [B] [frames] [out.json]

Prints, over the traced launches, the distribution (relative to the first CTA whose dependency resolves) of: residency,
dependency resolution, completion, and the CTA durations -- i.e. how much of the launch is CTA dispatch / tail."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from valle_b200 import _lib  # noqa: E402


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
    cap = 1 << 20
    ring = torch.zeros(cap, dtype=torch.int64, device=dev)
    cnt = torch.zeros(1, dtype=torch.int32, device=dev)
    _lib.check(lib.vb_trace_bind(ring.data_ptr(), cnt.data_ptr(), cap | (1 << 31)), "vb_trace_bind")
    eng.generate(texts, prompts, top_k=1, max_new_tokens=frames, return_device=True)
    torch.cuda.synchronize()
    n = int(cnt.item()) & 0xffffffff
    lib.vb_trace_bind(0, 0, 1)
    raw = ring.cpu().numpy().astype(np.uint64)
    if n > cap:   # ring wrapped: keep the newest cap entries, drop the oldest partially overwritten region
        raw = np.concatenate([raw[n % cap:], raw[:n % cap]])
    else:
        raw = raw[:n]
    t = (raw >> np.uint64(24)).astype(np.int64)
    bid = ((raw >> np.uint64(8)) & np.uint64(0xffff)).astype(np.int64)
    kind = (raw & np.uint64(0xff)).astype(np.int64)
    order = np.argsort(t, kind="stable")
    t, bid, kind = t[order], bid[order], kind[order]
    # launches: every CTA's stamps come in (16, 17, 18) triples in time order; cut the global sequence at the
    # "done" stamp count: launch k of a CTA = its k-th triple
    n_cta = int(bid.max()) + 1
    seqs = {}
    for k in (16, 17, 18):
        m = kind == k
        tk, bk = t[m], bid[m]
        o = np.lexsort((tk, bk))            # by CTA, then by time
        tk, bk = tk[o], bk[o]
        cuts = np.searchsorted(bk, np.arange(n_cta + 1))
        seqs[k] = [tk[cuts[c]:cuts[c + 1]] for c in range(n_cta)]
    n_launch = min(min(len(s) for s in seqs[k]) for k in (16, 17, 18))
    # align from the END (the ring may have cut the beginning)
    res, dep, done = (np.stack([s[len(s) - n_launch:] for s in seqs[k]]) for k in (16, 17, 18))   # [n_cta, n_launch]
    t0 = dep.min(axis=0, keepdims=True)
    use = slice(max(0, n_launch - 240), n_launch)
    r = {}
    for name, a in (("resident", res - t0), ("dependency", dep - t0), ("done", done - t0), ("duration", done - dep)):
        v = a[:, use].astype(np.float64) / 1000.0
        r[name] = {f"p{p}": float(np.percentile(v, p)) for p in (0, 10, 50, 90, 99, 100)}
        r[name]["mean"] = float(v.mean())
    span = (done.max(axis=0) - dep.min(axis=0))[use] / 1000.0
    late = ((dep - t0)[:, use] > 1000).mean()
    print(f"B={B}: {n_cta} CTAs per launch, {n_launch} launches traced, last {span.size} used; "
          f"launch span (first dependency -> last CTA done) mean {span.mean():.2f} us; "
          f"CTAs whose dependency resolves > 1 us after the first: {100 * late:.1f} %")
    for k, v in r.items():
        print(f"  {k:11s} " + "  ".join(f"{a}={b:7.2f}" for a, b in v.items()))
    if out:
        json.dump(dict(B=B, n_cta=n_cta, launches=int(n_launch), span_us=float(span.mean()), late_frac=float(late), **r),
                  open(out, "w"), indent=1)


if __name__ == "__main__":
    main()
