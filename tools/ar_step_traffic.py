"""Sum the DRAM traffic of ONE AR decode step from an ncu per-kernel CSV (metrics dram__bytes_read.sum,
dram__bytes_write.sum, gpu__time_duration.sum; eager launches, --cache-control none) and write
profiles/<name>.json.  A step = the kernels between two consecutive ar_sample_kernel launches.

    VB_NO_GRAPH=1 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
        --clock-control none --cache-control none -s 37000 -c 400 --csv --log-file gpurun_out/ar_step_dram.csv \
        python tools/profile_decode.py 64 380 bf16 ar_only
    python tools/ar_step_traffic.py gpurun_out/ar_step_dram.csv profiles/round1_ar_step_traffic.json 64 47 225 372
"""
import csv
import json
import os
import sys


def main():
    src, dst = sys.argv[1], sys.argv[2]
    B, S, Tp, n_gen = (int(v) for v in sys.argv[3:7])
    rows = []
    with open(src) as f:
        lines = [ln for ln in f if ln.startswith('"')]
    rd = csv.DictReader(lines)
    per = {}
    order = []
    for r in rd:
        kid = r["ID"]
        if kid not in per:
            per[kid] = {"name": r["Kernel Name"], "read": 0.0, "write": 0.0, "us": 0.0}
            order.append(kid)
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"].lower()
        scale = {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ns": 1e-3, "us": 1.0, "usecond": 1.0,
                 "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3}.get(unit, 1.0)
        if r["Metric Name"] == "dram__bytes_read.sum":
            per[kid]["read"] = v * scale
        elif r["Metric Name"] == "dram__bytes_write.sum":
            per[kid]["write"] = v * scale
        elif r["Metric Name"] == "gpu__time_duration.sum":
            per[kid]["us"] = v * scale
    seq = [per[k] for k in order]
    marks = [i for i, k in enumerate(seq) if "ar_sample_kernel" in k["name"]]
    if len(marks) < 2:
        raise SystemExit("need two ar_sample_kernel launches in the capture")
    step = seq[marks[0] + 1: marks[1] + 1]
    L = S + Tp + n_gen
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench  # the same algorithmic byte model as the bench line's roofline
    out = {
        "kernels_in_step": len(step),
        "context_len": L,
        "dram_read_bytes": sum(k["read"] for k in step),
        "dram_write_bytes": sum(k["write"] for k in step),
        "kernel_time_us_serialised": sum(k["us"] for k in step),
        "algorithmic_bytes": bench.ar_step_bytes(B, L, 2),
        "source": src,
    }
    out["traffic_bytes"] = out["dram_read_bytes"] + out["dram_write_bytes"]
    out["traffic_over_algorithmic"] = out["traffic_bytes"] / out["algorithmic_bytes"]
    with open(dst, "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
