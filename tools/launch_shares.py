"""Per-kernel launch count, average duration and share of the summed GPU time from an ncu launch list
(`ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file list.csv ...`).
    python tools/launch_shares.py list.csv"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1], newline="")) if len(r) > 14]
hdr = next(r for r in rows if "Kernel Name" in r)
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
acc = collections.defaultdict(list)
for r in rows:
    if r is hdr or len(r) != len(hdr):
        continue
    v = float(r[vi].replace(",", ""))
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[ui].replace("second", "s").replace("nsecond", "ns"), 1e-3)
    acc[r[ki]].append(v)
total = sum(sum(v) for v in acc.values())
for name, v in sorted(acc.items(), key=lambda kv: -sum(kv[1])):
    print(f"{name[:64]:64s} n={len(v):3d} avg={sum(v) / len(v):9.1f}us share={100 * sum(v) / total:5.1f}%")
