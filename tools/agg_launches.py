"""Aggregate an ncu launch list (gpu__time_duration.sum CSV) by kernel for the launches after the last occurrence
of a marker kernel:  python tools/agg_launches.py gpurun_out/x.csv k_rows_delta"""
import collections
import csv
import re
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
rows = [(x["Kernel Name"], float(x["Metric Value"].replace(",", ""))) for x in csv.DictReader(lines)]
marker = sys.argv[2] if len(sys.argv) > 2 else None
if marker:
    idx = [i for i, (n, _) in enumerate(rows) if marker in n]
    rows = rows[idx[-1]:]
agg = collections.defaultdict(lambda: [0, 0.0])
for n, v in rows:
    k = re.sub(r"\(.*", "", n)
    agg[k][0] += 1
    agg[k][1] += v / 1e3
for k, v in sorted(agg.items(), key=lambda t: -t[1][1]):
    print(f"{k:50s} {v[0]:5d} {v[1]:10.1f} us")
print("total us", sum(v[1] for v in agg.values()))
