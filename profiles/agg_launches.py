"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel.
usage: python profiles/agg_launches.py launches.csv [epochs_in_run] [--seq]"""
import collections
import csv
import re
import sys

path = sys.argv[1]
epochs = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
with open(path) as f:
    lines = [l for l in f if not l.startswith('==')]
agg = collections.defaultdict(lambda: [0, 0.0])
seq = []
for row in csv.DictReader(lines):
    name = row['Kernel Name']
    key = re.sub(r'\(.*', '', name)[:48]
    m = re.search(r'tc_gemm_kernel<\(na::tc::Mode\)(\d), \(bool\)(\d), \(bool\)(\d), (\d+)', name) or \
        re.search(r'tc_gemm_kernel<(\d), (\d), (\d), (\d+)', name)
    if m:
        mode = ['raw', 'fwd_sine', 'fwd_out', 'dx', 'dw', 'fwd_dot'][int(m.group(1))]
        key = f'tc_gemm[{mode}] BN{m.group(4)}'
    m = re.search(r'chain_kernel<(?:\(int\))?(\d+)', name)
    if m:
        key = f'chain<{m.group(1)}>'
    m2 = re.search(r'sgemm_kernel<(?:\(int\))?(\d)', name)
    if m2:
        key = 'sgemm[' + ['fwd_sine', 'fwd_out', 'dx', 'dw', 'fwd_eval', 'fwd_dot'][int(m2.group(1))] + ']'
    v = float(row['Metric Value'].replace(',', ''))
    unit = row['Metric Unit']
    v = v / 1e3 if unit in ('ns', 'nsecond') else v * 1e3 if unit in ('ms', 'msecond') else v
    agg[key][0] += 1
    agg[key][1] += v
    seq.append((key, row['Grid Size'], v))
tot = sum(v[1] for v in agg.values())
print(f'{"us total":>12} {"share":>6} {"launches":>8} {"avg us":>9} {"us/epoch":>9}  kernel')
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f'{v[1]:12.1f} {100 * v[1] / tot:5.1f}% {v[0]:8d} {v[1] / v[0]:9.1f} {v[1] / epochs:9.1f}  {k}')
print(f'{tot:12.1f} total')
if '--seq' in sys.argv:
    i0 = next((i for i, s in enumerate(seq) if s[0].startswith('chain') or 'fwd_sine' in s[0]), 0)
    for k, g, v in seq[i0:i0 + 45]:
        print(f'{v:9.1f} {g:18s} {k}')
