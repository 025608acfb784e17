"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total, average, share."""
import collections
import csv
import sys

path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
seq = []
for row in csv.DictReader(lines):
    name = row["Kernel Name"].split("(")[0][:70] + " grid=" + row.get("Grid Size", "")
    v = float(row["Metric Value"].replace(",", ""))
    unit = row["Metric Unit"]
    v = v / 1000 if unit == "ns" else (v * 1000 if unit == "ms" else v)
    agg[name][0] += 1
    agg[name][1] += v
    seq.append((name, v))
tot = sum(v[1] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:100s} n={v[0]:4d} total={v[1]:10.1f}us avg={v[1] / v[0]:8.2f}us share={v[1] / tot * 100:5.1f}%")
print("total us", round(tot, 1), "launches", len(seq))
