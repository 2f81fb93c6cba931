"""Share of an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel family.  python tools/launch_categories.py launches.csv 'header'"""
import collections
import csv
import sys

path, header = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else '')
lines = [ln for ln in open(path, errors='replace') if not ln.startswith('==')]
rows = list(csv.reader(lines))
h = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
idx = {k: i for i, k in enumerate(rows[h])}
FAMILIES = [('conv_igemm', 'convolution forward / data gradient (tcgen05)'), ('conv_rows', 'convolution forward / data gradient (tcgen05)'),
            ('conv_wgrad', 'weight gradient (tcgen05)'), ('wgrad_reduce', 'weight gradient (tcgen05)'), ('upfirdn', 'upfirdn2d'),
            ('bias_act', 'bias_act'), ('demod_act', 'modulation / demodulation passes'), ('mod_scale', 'modulation / demodulation passes'),
            ('amax', 'fp16 x 3 operand preparation'), ('split_act', 'fp16 x 3 operand preparation'), ('pack_weight_pair', 'fp16 x 3 operand preparation'),
            ('slab_reduce', 'fp16 x 3 operand preparation'), ('band_reduce', 'reductions of the fused passes'), ('reduce_partials', 'reductions of the fused passes'),
            ('pack_weight', 'weight packing'), ('fc_', 'fully-connected'), ('modprep', 'operand preparation (modprep)'), ('rgb', 'ToRGB / FromRGB'),
            ('aug_', 'ADA pipe'), ('adam', 'Adam / EMA'), ('ema', 'Adam / EMA')]


def family(k):
    for key, name in FAMILIES:
        if key in k:
            return name
    if 'at::' in k or 'cutlass' in k or 'nvjet' in k or 'cublas' in k or 'gemm' in k.lower():
        return 'ATen / library glue (add, cat, copy, cast, randn, small GEMMs)'
    return 'other'


t, n = collections.Counter(), collections.Counter()
for r in rows[h + 1:]:
    if len(r) > idx['Metric Value'] and r[idx['Metric Name']] == 'gpu__time_duration.sum':
        v = float(r[idx['Metric Value']].replace(',', ''))
        us = v * {'ns': 1e-3, 'nsecond': 1e-3, 'us': 1.0, 'usecond': 1.0, 'ms': 1e3, 'msecond': 1e3}.get(r[idx['Metric Unit']].lower(), 1e-3)
        f = family(r[idx['Kernel Name']])
        t[f] += us
        n[f] += 1
total = sum(t.values())
print(f'# {header}')
print(f'# {sum(n.values())} launches, {total / 1e3:.2f} ms serialised (cold-cache ncu replays: compare shares)')
print(' share%   total_us launches  family')
for f, v in t.most_common():
    print(f'{v / total * 100:7.2f} {v:10.1f} {n[f]:8d}  {f}')
