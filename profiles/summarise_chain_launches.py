"""Summarise the ncu launch list of the fused path per epoch and write ncu_traffic.json.

usage: python profiles/summarise_chain_launches.py profiles/launches_r01_chain.csv [epochs=3]

The list comes from
  NERFATTN_NO_GRAPH=1 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
      --clock-control none -c 200 --csv --log-file launches.csv python bench.py --steps 1 --warmup 0 --epochs 3 --no-e2e
Only the launches of the first `epochs` epochs (up to `epochs` x 5 Adam launches) are counted; what follows
belongs to bench.py's phase-isolation passes.  Times are serialised and cold-cache: compare shares.
"""
import collections
import csv
import json
import os
import re
import sys

path = sys.argv[1]
epochs = int(sys.argv[2]) if len(sys.argv) > 2 else 3
with open(path) as f:
    lines = [l for l in f if not l.startswith('==')]

launches = collections.OrderedDict()
for row in csv.DictReader(lines):
    rec = launches.setdefault(int(row['ID']), {'name': row['Kernel Name']})
    v = float(row['Metric Value'].replace(',', ''))
    unit = row['Metric Unit']
    if row['Metric Name'] == 'gpu__time_duration.sum':
        rec['us'] = v / 1e3 if unit.startswith('n') else v * 1e3 if unit.startswith('m') else v
    else:
        scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[unit]
        rec['rd' if 'read' in row['Metric Name'] else 'wr'] = v * scale


def label(name):
    m = re.search(r'chain_kernel<(?:\(int\))?(\d+), (?:\(int\))?(\d+), (?:\(int\))?(\d+), (?:\(int\))?(\d+)', name)
    if m:
        return f'chain<H={m.group(1)}, slots={m.group(2)}, cluster={m.group(4)}>'
    m = re.search(r'dw_adam_kernel<(?:\(int\))?(\d+)', name)
    if m:
        return f'dw_adam<BN={m.group(1)}>'
    m = re.search(r'tc_gemm_kernel<(?:\(na::tc::Mode\))?(\d)', name)
    if m:
        return 'tc_gemm[' + ['raw', 'fwd_sine', 'fwd_out', 'dx', 'dw', 'fwd_dot'][int(m.group(1))] + ']'
    m = re.search(r'(\w+::)?(\w+_kernel)', name)
    return (m.group(1) or '') + m.group(2) if m else name[:40]


# keep our own kernels from the first chain launch until `epochs` epochs of Adam (5 shape groups) are done
seq = [(label(r['name']), r) for r in launches.values()]
i0 = next(i for i, (k, _) in enumerate(seq) if k.startswith('chain'))
picked, adam = [], 0
for k, r in seq[i0:]:
    if 'at::' in r['name'] or 'sgemm' in k:
        break
    if 'mirror' in k or 'scale_params' in k:      # one-off plan set-up, not part of an epoch
        continue
    if 'xop' in k:                                # one-off plan set-up
        continue
    picked.append((k, r))
    if 'adam_kernel' in k and 'dw_adam' not in k:
        adam += 1
        if adam == 5 * epochs:
            break

agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for k, r in picked:
    a = agg[k]
    a[0] += 1
    a[1] += r['us']
    a[2] += r.get('rd', 0.0)
    a[3] += r.get('wr', 0.0)
tot = sum(a[1] for a in agg.values())
print(f'{"us/epoch":>10} {"share":>6} {"launches/ep":>11} {"rd MB/ep":>10} {"wr MB/ep":>10}  kernel')
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f'{a[1] / epochs:10.1f} {100 * a[1] / tot:5.1f}% {a[0] / epochs:11.1f} {a[2] / epochs / 1e6:10.1f} '
          f'{a[3] / epochs / 1e6:10.1f}  {k}')
print(f'{tot / epochs:10.1f} total per epoch (serialised, cold-cache ncu times)')

chain = [a for k, a in agg.items() if k.startswith('chain')]
dwa = [a for k, a in agg.items() if k.startswith('dw_adam')]
out = {
    'source': f'{path}: ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum '
              f'--clock-control none, NERFATTN_NO_GRAPH=1 python bench.py --steps 1 --warmup 0 --epochs 3 --no-e2e; '
              f'first {epochs} epochs',
    'chain_dram_bytes_per_epoch': int(sum(a[2] + a[3] for a in chain) / epochs),
    'dw_adam_dram_bytes_per_epoch': int(sum(a[2] + a[3] for a in dwa) / epochs),
    'dw_adam_us_per_epoch_ncu': round(sum(a[1] for a in dwa) / epochs, 1),
    'step_dram_bytes_per_epoch': int(sum(a[2] + a[3] for a in agg.values()) / epochs),
    'chain_us_per_epoch_ncu': round(sum(a[1] for a in chain) / epochs, 1),
    'step_us_per_epoch_ncu': round(tot / epochs, 1),
}
with open(os.path.join(os.path.dirname(path) or '.', 'ncu_traffic.json'), 'w') as f:
    json.dump(out, f, indent=1)
print(json.dumps(out))
