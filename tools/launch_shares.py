#!/usr/bin/env python
"""Shares of the step per kernel from an ncu launch list (`--metrics gpu__time_duration.sum --csv --log-file x.csv`):
    python tools/launch_shares.py gpurun_out/x.csv > profiles/x.summary.txt
Per-launch times under ncu are cold-cache and serialised: compare SHARES with the live CUDA-event shares, not absolutes."""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) >= 15 and r[0].isdigit()]
tot = sum(float(r[14]) for r in rows)
agg = collections.OrderedDict()
for r in rows:
    k = f"{r[4][:110]} grid {r[8]}"
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += float(r[14])
print(f"# {sys.argv[1]}: {len(rows)} launches, {tot / 1e6:.3f} ms in total (gpu__time_duration.sum)")
for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{100 * ns / tot:6.2f} %  {ns / 1e6:9.3f} ms  {n:4d}x  {k}")
