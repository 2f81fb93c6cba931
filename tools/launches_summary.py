"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into the per-kernel share table kept under profiles/.
Usage: python tools/launches_summary.py gpurun_out/launches.csv 'header' > profiles/rNN_launches_summary.txt"""
import collections
import csv
import sys

path = sys.argv[1]
header = sys.argv[2] if len(sys.argv) > 2 else ''
lines = [ln for ln in open(path, errors='replace') if not ln.startswith('==')]
rows = list(csv.reader(lines))
hdr = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
idx = {h: i for i, h in enumerate(rows[hdr])}
agg = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) <= idx['Metric Value'] or r[idx['Metric Name']] != 'gpu__time_duration.sum':
        continue
    v = float(r[idx['Metric Value']].replace(',', ''))
    unit = r[idx['Metric Unit']].lower()
    us = v * {'ns': 1e-3, 'nsecond': 1e-3, 'us': 1.0, 'usecond': 1.0, 'ms': 1e3, 'msecond': 1e3}.get(unit, 1e-3)
    a = agg.setdefault(r[idx['Kernel Name']], [0.0, 0])
    a[0] += us
    a[1] += 1
total = sum(a[0] for a in agg.values())
n = sum(a[1] for a in agg.values())
print(f'# {header}')
print('# Times are cold-cache, serialised (ncu replays every launch): compare SHARES, not absolute times.')
print(f'# total device time of the {n} captured launches: {total / 1e3:.2f} ms')
print(' share%   total_us launches    avg_us  kernel')
for k, (t, c) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:80]:
    print(f'{t / total * 100:7.2f} {t:10.1f} {c:8d} {t / c:9.1f}  {k[:150]}')
