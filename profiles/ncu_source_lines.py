"""Warp-stall samples per CUDA source line of one `ncu --set full --import-source on` report.
usage: python profiles/ncu_source_lines.py report.ncu-rep [top=40]"""
import csv
import subprocess
import sys

rep, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'],
                     capture_output=True, text=True).stdout
rows, fname, hdr = [], '', None
for r in csv.reader(raw.splitlines()):
    if len(r) >= 2 and r[0] == 'File Name':
        fname = r[1].split('/')[-1]
    elif r and r[0] == 'Line No':
        hdr = r
    elif hdr and len(r) == len(hdr) and r[0].isdigit():
        rows.append((fname, r))
ix = {h: i for i, h in enumerate(hdr)}


def f(r, k):
    try:
        return float(r[ix[k]].replace(',', ''))
    except ValueError:
        return 0.0


tot = sum(f(r, '# Samples') for _, r in rows)
print(f'{rep}: {tot:.0f} samples')
print(f'{"samples":>8} {"share":>6} {"long_sb":>8} {"short_sb":>8} {"wait":>6} {"barrier":>7} {"mio":>5} {"instr":>10}  line')
for fn, r in sorted(rows, key=lambda fr: -f(fr[1], '# Samples'))[:top]:
    print(f"{f(r, '# Samples'):8.0f} {100 * f(r, '# Samples') / tot:5.1f}% {f(r, 'stall_long_sb'):8.0f} {f(r, 'stall_short_sb'):8.0f} "
          f"{f(r, 'stall_wait'):6.0f} {f(r, 'stall_barrier'):7.0f} {f(r, 'stall_mio'):5.0f} {f(r, 'Instructions Executed'):10.0f}  "
          f"{fn}:{r[0]} {r[1].strip()[:100]}")
