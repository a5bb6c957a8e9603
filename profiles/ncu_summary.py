"""Key metrics of one `ncu --set full` report as text.  usage: python profiles/ncu_summary.py report.ncu-rep 'header line'"""
import csv
import subprocess
import sys

rep, header = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else '')
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ['Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration.sum', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_registers', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'lts__t_sector_hit_rate.pct',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__cycles_elapsed.avg']
if header:
    print(header)
for w in want:
    if w in hdr:
        i = hdr.index(w)
        print(f'{w:75s} {vals[i]} {units[i]}')
for i, h in enumerate(hdr):
    if 'pcsamp_warps_issue_stalled' in h and 'not_issued' not in h:
        try:
            if float(vals[i].replace(',', '')) > 0.03 * float(vals[hdr.index('smsp__pcsamp_sample_count')].replace(',', '')):
                print(f'{h:75s} {vals[i]} {units[i]}')
        except ValueError:
            pass
