#!/usr/bin/env python
"""Condense an `ncu --metrics gpu__time_duration.sum --csv` launch list into one line per kernel (count, total, mean, share).
    python tools/launch_summary.py gpurun_out/launches_<tag>.csv > profiles/<tag>_launches_summary.txt
Times under ncu are cold-cache and serialised: the SHARES are what to compare with bench.py's live numbers."""
import collections
import csv
import sys

rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
h = rows[0]
ki, vi = h.index('Kernel Name'), h.index('Metric Value')
agg = collections.OrderedDict()
for r in rows[1:]:
    try:
        v = float(r[vi].replace(',', ''))
    except ValueError:
        continue
    a = agg.setdefault(r[ki][:110], [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
print(f"# {sys.argv[1]}: {sum(a[0] for a in agg.values())} launches, {tot / 1e6:.3f} ms under ncu (whole `bench.py --steps 2 --warmup 1` process:"
      " forward bench, Philox leg, e2e, heads + torch comparison, training steps)")
print(f"{'count':>6} {'total ms':>10} {'mean us':>10} {'share':>6}  kernel")
for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{a[0]:6d} {a[1] / 1e6:10.3f} {a[1] / a[0] / 1e3:10.1f} {100 * a[1] / tot:5.1f}%  {k}")
